"""Drop-in nn.Module facades for the GSN hot path.

Same constructor arguments, forward() return tuples, parameter / buffer names and shapes as the
reference modules, so `audiozen/trainer.py` and the model-zoo `pytorch_model.bin` files work
unchanged; the arithmetic between "magnitude in" and "coefficients out" runs in libgsn_b200.so.

Reference modules mirrored (paths relative to the reference root):
  ESN = audiozen/models/spiking_fullsubnet/efficient_spiking_neuron.py  (GSUCell :104, GSULayer :70,
        StackedGSU :43, efficient_spiking_neuron :12)
  MSF = audiozen/models/spiking_fullsubnet/modeling_spiking_fullsubnet.py (SequenceModel :12,
        SubBandSequenceModel :128, SubbandModel :172, SpikingFullSubNet :349)
  CGN = audiozen/models/cirm_gsn/modeling_cirm_gsn.py (Model :162)

The modules hold parameters only; there is no per-frame Python loop and no CPU / eager fallback:
CPU tensors raise.  Inference (no grad, eval mode) takes the fused path below; when gradients are required
or BatchNorm is in training mode, `forward` takes the autograd path of `training.py` (CUDA forward/BPTT
kernels for the recurrence, differentiable torch ops around it).
"""
from __future__ import annotations

import math
import os
from collections import namedtuple


import torch
import torch.nn as nn

from . import ops

MemoryState = namedtuple("MemoryState", ["hx", "cx"])

__all__ = ["MemoryState", "efficient_spiking_neuron", "GSUCell", "GSULayer", "StackedGSU",
           "SequenceModel", "SubBandSequenceModel", "SubbandModel", "SpikingFullSubNet", "CirmGSN", "Separator"]


# ------------------------------------------------------------------------------------------------
# neuron stack (ESN)
# ------------------------------------------------------------------------------------------------
class GSUCell(nn.Module):
    """Parameter holder of one gated spiking layer: weight_ih [gH,K], weight_hh [gH,H], bias_ih [2H],
    optional batchnorm (ESN:105-130).  All three parameters start U(+-1/sqrt(H)) like the reference
    (whose reset_parameters also overwrites the zero bias)."""

    def __init__(self, input_size, hidden_size, shared_weights=False, bn=False):
        super().__init__()
        self.input_size, self.hidden_size = input_size, hidden_size
        self.shared_weights, self.use_bn = shared_weights, bn
        g = 1 if shared_weights else 2
        self.weight_ih = nn.Parameter(torch.empty(g * hidden_size, input_size))
        self.weight_hh = nn.Parameter(torch.empty(g * hidden_size, hidden_size))
        self.bias_ih = nn.Parameter(torch.empty(2 * hidden_size))
        bound = 1.0 / math.sqrt(hidden_size) if hidden_size > 0 else 0.0
        for p in (self.weight_ih, self.weight_hh, self.bias_ih):
            nn.init.uniform_(p, -bound, bound)
        if bn:
            self.batchnorm = nn.BatchNorm1d(hidden_size)
        self._bn_cache = None

    def folded_bn(self):
        """Eval-mode BatchNorm as the per-channel affine torch's own CPU kernel applies:
        alpha = weight / sqrt(running_var + eps), beta = bias - running_mean * alpha."""
        if not self.use_bn:
            return None, None
        bn = self.batchnorm
        # the training kernels update the running statistics through raw pointers (no _version bump) but always advance
        # num_batches_tracked (training.GSNLayerFn), so its version is part of the key
        nbt = bn.num_batches_tracked
        key = (bn.weight._version, bn.bias._version, bn.running_mean._version, bn.running_var._version,
               nbt._version if nbt is not None else 0, bn.weight.data_ptr(), bn.running_var.data_ptr())
        if self._bn_cache is None or self._bn_cache[0] != key:
            with torch.no_grad():
                invstd = 1.0 / torch.sqrt(bn.running_var + bn.eps)
                alpha = (invstd * bn.weight).contiguous()
                beta = (bn.bias - bn.running_mean * alpha).contiguous()
            old = self._bn_cache
            if old is not None and old[1].shape == alpha.shape and old[1].device == alpha.device:
                # refresh IN PLACE: captured CUDA graphs hold these two tensors (network() calls refresh_folded_bn
                # before every replay, so replays see updated parameters / statistics without re-capturing)
                old[1].copy_(alpha)
                old[2].copy_(beta)
                self._bn_cache = (key, old[1], old[2])
            else:
                self._bn_cache = (key, alpha, beta)
        return self._bn_cache[1], self._bn_cache[2]

    def forward(self, input, state):
        """One frame (ESN:132-153): input [R,K], state (h [R,H], c [R,H]) -> (h, (h, c))."""
        if _needs_autograd(self):
            _autograd_zero_state_only(state, "GSUCell.forward")
            from . import training
            h = training.run_cell_layer(self, input.unsqueeze(0).contiguous())
            return h[0], MemoryState(h[0], None)
        h, c, (hT, cT) = _run_layer(self, input.unsqueeze(0).contiguous(), state, want_c=False,
                                    backend="auto")
        return hT, MemoryState(hT, cT)


class GSULayer(nn.Module):
    def __init__(self, cell, *cell_args):
        super().__init__()
        self.cell = cell(*cell_args)

    def forward(self, input, state):
        """input [T,R,K] -> (h [T,R,H], final state) (ESN:75-81) -- one kernel, no Python time loop."""
        if _needs_autograd(self):
            _autograd_zero_state_only(state, "GSULayer.forward")
            from . import training
            h = training.run_cell_layer(self.cell, input.contiguous())
            return h, MemoryState(h[-1], None)
        h, _, (hT, cT) = _run_layer(self.cell, input, state, want_c=False, backend="auto")
        return h, MemoryState(hT, cT)


def _lib_supported_pre(K, H):
    from . import _lib
    return bool(_lib.load().gsn_pre_stream_supported(int(K), int(H)))


def _cluster_ctas(rows, H, shared, nt=16):
    """CTAs of one recurrence at row tiling `nt` (16 = finest, 64 = coarsest): its SM demand."""
    return ((rows + nt - 1) // nt) * ((H + 127) // 128 if shared else (H + 63) // 64)


def _chunk_bounds(T, nchunks):
    """Frame ranges of the wavefront schedule: `nchunks` equal chunks.  (GSN_WF_WEIGHTS="1,2,4,..." sets relative
    chunk lengths instead -- a development knob: tapered schedules with short first / last chunks measured 4-9 %
    SLOWER than equal chunks on B200, see profiles/r01_schedule_experiments.md.)"""
    nchunks = max(1, min(nchunks, T))
    env = os.environ.get("GSN_WF_WEIGHTS")
    w = [float(v) for v in env.split(",")] if env else [1.0] * nchunks
    cum, acc = [0], 0.0
    for v in w:
        acc += v
        cum.append(int(round(T * acc / sum(w))))
    cum[-1] = T
    return [(a, b) for a, b in zip(cum[:-1], cum[1:]) if b > a]


def _sm_budgets(demands, total=148, floor=4):
    """Split the SMs between concurrently running recurrences in proportion to their demand (0 = no cap when
    everything fits at the finest tiling)."""
    if sum(demands) <= total:
        return [0] * len(demands)
    return [max(floor, int(total * d / sum(demands))) for d in demands]


def _autograd_zero_state_only(state, who):
    """The autograd kernels (training.GSNLayerFn) start every sequence from a zero state, which is what every caller
    of the reference passes (MSF:100-106).  Anything else must not be silently ignored."""
    if state is None:
        return
    h, c = state[0], state[1]
    if (h is not None and bool((h != 0).any())) or (c is not None and bool((c != 0).any())):
        raise NotImplementedError(f"{who}: a non-zero initial state is not supported on the autograd (training) path; "
                                  f"call it under torch.no_grad() with eval-mode BatchNorm, or start from zeros")


def _needs_autograd(module):
    """True when the call must go through the autograd path: gradients are being recorded for some parameter,
    or a BatchNorm of the module is in training mode (batch statistics, running-stat updates)."""
    if torch.is_grad_enabled() and any(p.requires_grad for p in module.parameters()):
        return True
    return any(isinstance(m, nn.BatchNorm1d) and m.training for m in module.modules())


def _run_layer(cell, x, state, want_c, backend, sm_budget=0, spikes_in=False, bits_in=None, bits_out=None):
    """bits_in: bit-packed copy of a spike-trace input `x`; bits_out: buffer for the bit-packed copy of the
    returned trace (ops.spike_bits_buffer) -- what the next spike-input linear reads instead of fp32."""
    if cell.use_bn and cell.batchnorm.training:
        raise RuntimeError("internal error: training-mode BatchNorm reached the inference kernels")
    if not x.is_cuda:
        raise RuntimeError("spiking_fullsubnet_b200 has no CPU path: move the model and input to CUDA")
    x = x.contiguous()
    # (bias joins in the recurrence, in the reference's order; layers >= 1 see the previous layer's spikes)
    xproj = ops.linear(x, cell.weight_ih.detach(), spikes=spikes_in, sm_budget=sm_budget, bits=bits_in)
    a, b = cell.folded_bn()
    h0 = c0 = None
    if state is not None:
        h0, c0 = state[0].contiguous(), state[1].contiguous()
    return ops.layer_recurrence(xproj, cell.weight_hh.detach(), cell.bias_ih.detach(), a, b,
                                shared=cell.shared_weights, want_c=want_c, h0=h0, c0=c0,
                                want_state=True, backend=backend, sm_budget=sm_budget, out_bits=bits_out)


class StackedGSU(nn.Module):
    def __init__(self, num_layers, layer, first_layer_args, other_layer_args):
        super().__init__()
        self.layers = nn.ModuleList([layer(*first_layer_args)] +
                                    [layer(*other_layer_args) for _ in range(num_layers - 1)])
        self.backend = "auto"
        self.sm_budget = 0  # SMs to plan for (0 = whole device); set by callers that run stacks concurrently

    def forward(self, input, states=None, want_c=False):
        """(ESN:50-62) input [T,R,K], states list of (h,c) or None (zeros) ->
        (output [T,R,H], output_states, all_layer_output = [input, h1..hL])."""
        if _needs_autograd(self):
            # training path: zero initial state (what every caller of the reference passes, MSF:100-106)
            from . import training
            for st in (states or []):
                _autograd_zero_state_only(st, "StackedGSU.forward")
            out, trace = training.run_stack(self, input.contiguous())
            self.last_c = [None] * len(self.layers)
            self.last_bits = None
            return out, [MemoryState(t[-1], None) for t in trace[1:]], trace
        out = input
        out_states, trace = [], [input]
        self.last_c = []
        bits = None
        for i, layer in enumerate(self.layers):
            st = None if states is None else states[i]
            H = layer.cell.hidden_size
            nbits = ops.spike_bits_buffer(out.shape[:-1], H, out.device) if ops.SPIKE_BITS[0] and H <= 320 else None
            h, c, (hT, cT) = _run_layer(layer.cell, out, st, want_c, self.backend, self.sm_budget,
                                        spikes_in=i > 0, bits_in=bits, bits_out=nbits)
            out_states.append(MemoryState(hT, cT))
            trace.append(h)
            self.last_c.append(c)
            out, bits = h, nbits
        self.last_bits = bits  # bit-packed copy of `out` for the caller's proj (None when disabled)
        return out, out_states, trace


def efficient_spiking_neuron(input_size, hidden_size, num_layers, shared_weights=False, bn=False,
                             batch_first=False):
    """Factory with the reference's signature (ESN:12-40)."""
    assert not batch_first
    return StackedGSU(num_layers, GSULayer,
                      first_layer_args=[GSUCell, input_size, hidden_size, shared_weights, bn],
                      other_layer_args=[GSUCell, hidden_size, hidden_size, shared_weights, bn])


# ------------------------------------------------------------------------------------------------
# sequence models (MSF)
# ------------------------------------------------------------------------------------------------
_ACTS = ("tanh", "sigmoid", "relu")


class SequenceModel(nn.Module):
    def __init__(self, input_size, hidden_size, num_layers, sequence_model="GSN", proj_size=0,
                 shared_weights=False, output_activate_function=None, bn=False, use_pre_layer_norm=True):
        super().__init__()
        if use_pre_layer_norm:
            self.pre_layer_norm = nn.LayerNorm(input_size)
        if sequence_model == "GSN":
            self.sequence_model = efficient_spiking_neuron(input_size, hidden_size, num_layers,
                                                           shared_weights=shared_weights, bn=bn)
        elif sequence_model == "LSTM":
            raise NotImplementedError("sequence_model='LSTM' is outside the GSN hot path (SURVEY 8a, a6); "
                                      "use the reference implementation for LSTM models")
        else:
            raise NotImplementedError(f"Sequence model {sequence_model} not implemented.")
        self.proj = nn.Linear(hidden_size, proj_size) if proj_size > 0 else nn.Identity()
        self.output_activate_function = {"tanh": nn.Tanh, "sigmoid": nn.Sigmoid, "relu": nn.ReLU}.get(
            output_activate_function, nn.Identity)()
        self._act = output_activate_function if output_activate_function in _ACTS else None
        self.input_size, self.hidden_size, self.num_layers = input_size, hidden_size, num_layers
        self.proj_size = proj_size
        self.use_pre_layer_norm = use_pre_layer_norm
        self.sequence_model_name = sequence_model

    # --- time-major core used by every facade -----------------------------------------------------
    def run_time_major(self, x):
        """x [T,R,K] already normalised -> (proj_out [T,R,P], activated [T,R,P], all_layer_outputs)."""
        out, _, trace = self.sequence_model(x, None)
        if isinstance(self.proj, nn.Linear):
            res = ops.linear(out, self.proj.weight.detach(), self.proj.bias.detach(), act=self._act,
                             spikes=True, sm_budget=self.sequence_model.sm_budget,
                             bits=getattr(self.sequence_model, "last_bits", None))
            proj, act = res if self._act else (res, res)
        else:
            proj = out
            act = self.output_activate_function(out)
        return proj, act, trace + [proj]

    def forward(self, input):
        """input [R,K,T] -> (output [R,P,T], all_layer_outputs) (MSF:81-125)."""
        assert input.ndim == 3, f"Input tensor must be 3D, but got {input.ndim}D."
        if not input.is_cuda:
            raise RuntimeError("spiking_fullsubnet_b200 has no CPU path: move the model and input to CUDA")
        R, K, T = input.shape
        cm = input.permute(2, 0, 1).contiguous()  # 'b f t -> t b f'
        if _needs_autograd(self):  # differentiable LayerNorm / projections, autograd recurrence kernels
            from . import training
            x = self.pre_layer_norm(cm) if self.use_pre_layer_norm else cm
            act, all_out = training.run_sequence_model(self, x.contiguous())
            return act.permute(1, 2, 0), all_out
        lnw = self.pre_layer_norm.weight.detach() if self.use_pre_layer_norm else None
        lnb = self.pre_layer_norm.bias.detach() if self.use_pre_layer_norm else None
        eps = self.pre_layer_norm.eps if self.use_pre_layer_norm else 1e-5
        x = ops.subband_features(cm, None, 1, 0, K, 0, lnw, lnb, eps)
        _, act, all_out = self.run_time_major(x)
        return act.permute(1, 2, 0), all_out


class LazyOutputs(list):
    """all_layer_outputs of a streamed sequence model: same positions as the reference's list ([x_norm, h1..hL,
    proj_out], MSF:115-125), but the fp32 [T,R,.] tensors nobody may ever read (SynOps accounting is their only
    consumer, audiozen/metric.py:303-327) are produced on first access from what the kernels actually wrote
    (bit-packed spike traces; the gather is re-run for x_norm)."""

    def __init__(self, thunks, widths=None, spike_counts=None, trace_numel=None):
        super().__init__(thunks)
        # for the SynOps / NeuronOps accounting (metrics.py) without materialising anything: last-dim width of every
        # entry, the number of spikes each recurrence emitted (int64 device tensor, counted inside the kernels), and
        # the number of elements of every spike trace
        self.widths, self.spike_counts, self.trace_numel = widths, spike_counts, trace_numel

    def _resolve(self, i):
        v = list.__getitem__(self, i)
        if callable(v):
            v = v()
            list.__setitem__(self, i, v)
        return v

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self._resolve(k) for k in range(*i.indices(len(self)))]
        return self._resolve(i if i >= 0 else len(self) + i)

    def __iter__(self):
        return (self._resolve(k) for k in range(len(self)))

    def __add__(self, other):
        return list(self) + list(other)


class _SeqPlan:
    """Pre-allocated buffers of one sequence model for the frame-chunked wavefront schedule: the full
    [T,R,.] tensors every chunk writes its slice of, the carried (h,c) per layer and chunk boundary, the
    folded BatchNorm affines and one recurrence workspace per layer.  Everything is allocated up front on
    the caller's stream and stays alive until the streams have joined."""

    def __init__(self, model, T, R, nchunks, device, backend):
        self.m = model
        self.sm_budget = 0       # SMs the recurrences of this model plan for (0 = whole device)
        self.lin_budget = 0      # SMs the tcgen05 linears of this model plan for
        stack = model.sequence_model
        self.L = len(stack.layers)
        f32 = dict(device=device, dtype=torch.float32)
        self.x = torch.empty((T, R, model.input_size), **f32)
        self.xproj, self.h, self.hs, self.cs, self.bn, self.ws, self.hb = [], [], [], [], [], [], []
        for layer in stack.layers:
            cell = layer.cell
            if cell.use_bn and cell.batchnorm.training:
                raise RuntimeError("the inference schedule needs model.eval(); training goes through forward()")
            H = cell.hidden_size
            g = 1 if cell.shared_weights else 2
            self.xproj.append(torch.empty((T, R, g * H), **f32))
            self.h.append(torch.empty((T, R, H), **f32))
            self.hb.append(ops.spike_bits_buffer((T, R), H, device) if ops.SPIKE_BITS[0] and H <= 320 else None)
            self.hs.append(torch.zeros((nchunks + 1, R, H), **f32))
            self.cs.append(torch.zeros((nchunks + 1, R, H), **f32))
            self.bn.append(cell.folded_bn())
            self.ws.append(ops.recurrence_workspace(R, H, cell.shared_weights, backend, device))
        self.backend = backend
        if isinstance(model.proj, nn.Linear):
            self.proj = torch.empty((T, R, model.proj_size), **f32)
            self.act = torch.empty_like(self.proj) if model._act else self.proj
        else:
            self.proj = self.act = None

    def run_pre(self, l, k, t0, t1):
        """Input-to-hidden product of frames [t0,t1) of layer l (no dependence on the layer's own state)."""
        cell = self.m.sequence_model.layers[l].cell
        inp = self.x[t0:t1] if l == 0 else self.h[l - 1][t0:t1]
        bits = self.hb[l - 1][t0:t1] if l > 0 and self.hb[l - 1] is not None else None
        ops.linear(inp, cell.weight_ih.detach(), out=self.xproj[l][t0:t1], spikes=l > 0,
                   sm_budget=self.lin_budget, bits=bits)

    def run_rec(self, l, k, t0, t1, pdl=False):
        """Recurrence of frames [t0,t1) of layer l from the state carried out of chunk k-1.  pdl: enqueue it as a
        programmatic dependent of the previous kernel of the stream (= chunk k-1 of this layer)."""
        cell = self.m.sequence_model.layers[l].cell
        a, b = self.bn[l]
        if pdl:
            ops.set_option(ops.OPT_PDL, 1)
        try:
            self._run_rec(cell, a, b, l, k, t0, t1)
        finally:
            if pdl:
                ops.set_option(ops.OPT_PDL, 0)

    def _run_rec(self, cell, a, b, l, k, t0, t1):
        ops.layer_recurrence(self.xproj[l][t0:t1], cell.weight_hh.detach(), cell.bias_ih.detach(), a, b,
                             shared=cell.shared_weights, h0=self.hs[l][k], c0=self.cs[l][k],
                             out_h=self.h[l][t0:t1], out_hT=self.hs[l][k + 1], out_cT=self.cs[l][k + 1],
                             backend=self.backend, workspace=self.ws[l], sm_budget=self.sm_budget,
                             out_bits=self.hb[l][t0:t1] if self.hb[l] is not None else None)

    def run_post(self, k, t0, t1):
        """Output projection (+ activation) of frames [t0,t1)."""
        m = self.m
        if self.proj is not None:
            ops.linear(self.h[-1][t0:t1], m.proj.weight.detach(), m.proj.bias.detach(), act=m._act,
                       out=self.proj[t0:t1], out_act=self.act[t0:t1] if m._act else None, spikes=True,
                       sm_budget=self.lin_budget, bits=self.hb[-1][t0:t1] if self.hb[-1] is not None else None)

    def outputs(self):
        """(proj_out, activated, all_layer_outputs) exactly as SequenceModel.run_time_major returns them."""
        last = self.h[-1]
        if self.proj is None:
            return last, self.m.output_activate_function(last), [self.x] + self.h + [last]
        return self.proj, self.act, [self.x] + self.h + [self.proj]


class SubBandSequenceModel(SequenceModel):
    def __init__(self, df_order, num_spks, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.df_order, self.num_spks = df_order, num_spks

    def forward(self, input_features):
        """[B,N,1,fs,T] -> ([B,df,S,N*fc,T,2], all_layer_outputs) (MSF:134-169)."""
        B, N, C, fs, T = input_features.shape
        assert C == 1, "Only mono audio is supported."
        out, all_out = super().forward(input_features.reshape(B * N, fs, T))
        return coef_layout(out.permute(2, 0, 1), B, N, self.df_order, self.num_spks), all_out


def coef_layout(proj, B, N, df, S):
    """proj [T, B*N, P] -> [B, df, S, N*fc, T, 2]: '(b n) (c fc df s) t -> b df s (n fc) t c'
    (MSF:160-167); pure index arithmetic (a strided view + one copy)."""
    T, _, P = proj.shape
    fc = P // (2 * df * S)
    v = proj.reshape(T, B, N, 2, fc, df, S)
    return v.permute(1, 5, 6, 2, 4, 0, 3).reshape(B, df, S, N * fc, T, 2)


_BAND_STREAMS = {}


def _band_streams(device, n, priority=0, tag=None):
    key = (device.index, n, priority, tag)
    if key not in _BAND_STREAMS:
        _BAND_STREAMS[key] = [torch.cuda.Stream(device=device, priority=priority) for _ in range(n)]
    return _BAND_STREAMS[key]


class SubbandModel(nn.Module):
    def __init__(self, freq_cutoffs, center_freq_sizes, neighbor_freq_sizes, df_orders, num_spks, **kwargs):
        super().__init__()
        assert len(freq_cutoffs) - 1 == len(center_freq_sizes), "Number of subbands must be equal to len(cutoffs)."
        self.sb_models = nn.ModuleList([
            SubBandSequenceModel(input_size=(c + n * 2) + c, proj_size=2 * c * d * num_spks, df_order=d,
                                 num_spks=num_spks, **kwargs)
            for c, n, d in zip(center_freq_sizes, neighbor_freq_sizes, df_orders)])
        self.freq_cutoffs = freq_cutoffs
        self.center_freq_sizes = center_freq_sizes
        self.neighbor_freq_sizes = neighbor_freq_sizes
        self.df_orders = df_orders
        self.num_spks = num_spks
        self.concurrent_bands = True

    def run_time_major(self, cm, fb):
        """cm [T,B,F] compressed magnitude, fb [T,B,f_fb] full-band output (tiled by index) ->
        list of proj outputs [T, B*N_i, P_i] and the per-band all_layer_outputs (MSF:216-263)."""
        if _needs_autograd(self):  # differentiable gather / LayerNorm / projections, autograd recurrence kernels
            from . import training
            return training.run_subband_model(self, cm, fb)
        T, B, F = cm.shape
        for i in range(len(self.sb_models)):
            lo, hi, ctr = self.freq_cutoffs[i], self.freq_cutoffs[i + 1], self.center_freq_sizes[i]
            if (hi - lo) % ctr != 0:
                raise ValueError(f"Number of frequency bins must be divisible by the center frequency."
                                 f"GOT: ctr_freq={ctr}, upper_cutoff_freq={hi}, lower_cutoff_freq={lo}")

        def run_band(i):
            m = self.sb_models[i]
            lo, hi = self.freq_cutoffs[i], self.freq_cutoffs[i + 1]
            ctr, nbr = self.center_freq_sizes[i], self.neighbor_freq_sizes[i]
            lnw = m.pre_layer_norm.weight.detach() if m.use_pre_layer_norm else None
            lnb = m.pre_layer_norm.bias.detach() if m.use_pre_layer_norm else None
            eps = m.pre_layer_norm.eps if m.use_pre_layer_norm else 1e-5
            x = ops.subband_features(cm, fb, (hi - lo) // ctr, lo, ctr, nbr, lnw, lnb, eps)
            proj, _, all_out = m.run_time_major(x)
            return proj, all_out

        n = len(self.sb_models)
        demands = [_cluster_ctas(B * ((self.freq_cutoffs[i + 1] - self.freq_cutoffs[i]) // self.center_freq_sizes[i]),
                                 m.hidden_size, m.sequence_model.layers[0].cell.shared_weights)
                   for i, m in enumerate(self.sb_models)]
        budgets = _sm_budgets(demands) if self.concurrent_bands and n > 1 else [0] * n
        for m, b in zip(self.sb_models, budgets):
            m.sequence_model.sm_budget = b
        if not self.concurrent_bands or n == 1:
            res = [run_band(i) for i in range(n)]
        else:
            # the bands are independent recurrences on different weights: fork one stream per band so
            # their (latency-bound, few-CTA) kernels share the 148 SMs, then join on the caller's stream
            main = torch.cuda.current_stream(cm.device)
            streams = _band_streams(cm.device, n)
            fork = torch.cuda.Event()
            fork.record(main)
            res = []
            for i in range(n):
                streams[i].wait_event(fork)
                with torch.cuda.stream(streams[i]):
                    res.append(run_band(i))
                    done = torch.cuda.Event()
                    done.record(streams[i])
                main.wait_event(done)
                if not torch.cuda.is_current_stream_capturing():
                    for t in [res[-1][0]] + res[-1][1]:
                        t.record_stream(main)
        return [r[0] for r in res], [r[1] for r in res]

    def forward(self, noisy_input, fb_output):
        """noisy_input [B,1,F,T], fb_output [B,1,F,T] (already tiled) -> (coef list, all_layer_outputs)."""
        B, C, F, T = noisy_input.shape
        assert C == 1, "Only mono audio is supported."
        cm = noisy_input[:, 0].permute(2, 0, 1).contiguous()
        fb = fb_output[:, 0].permute(2, 0, 1).contiguous()
        projs, all_outs = self.run_time_major(cm, fb)
        coefs = [coef_layout(p, B, p.shape[1] // B, d, self.num_spks) for p, d in zip(projs, self.df_orders)]
        return coefs, all_outs


def _stft(y, n_fft, hop, win):
    # audiozen/acoustics/audio_feature.py:236-294 (hann window, center=True, pad_mode="constant")
    window = torch.hann_window(n_fft, device=y.device)
    return torch.stft(y, n_fft, hop, win, window=window, return_complex=True, pad_mode="constant")


def _stft_fused(y, n_fft, hop, win, f_keep=None, fdrc=None):
    """_stft for the inference path: ONE framing kernel (gsn_frame_signal: zero padding, framing, window) + cuFFT's
    batched real FFT instead of torch.stft's pad / as_strided / multiply / FFT sequence.  Same [B,F,T] view of cuFFT's
    [B,T,F] output.  (torch.stft stays on the training path: it is differentiable.)"""
    if win != n_fft or n_fft % 8 != 0 or y.dim() != 2 or y.requires_grad or not y.is_contiguous() or y.data_ptr() % 16:
        return _stft(y, n_fft, hop, win)
    window = _hann(n_fft, y.device)
    if _fft_kernels(n_fft, win):
        # the recipes' transform length: framing + window + real FFT + |X|**fdrc in ONE kernel (gsn_stft_compress); the
        # compressed magnitude rides along on the tensor and ops.compress_mag hands it to the network as is
        spec, cm = ops.stft_compress(y, window, hop, f_keep, 0.5 if fdrc is None else fdrc)
        if cm is not None:
            spec._gsn_cm = (int(f_keep), float(fdrc), cm)
        return spec
    return torch.fft.rfft(ops.frame_signal(y, window, hop), dim=-1).transpose(1, 2)


def _fft_kernels(n_fft, win):
    """True when the library's own FFT kernels apply (n_fft = win_length = 512; GSN_FFT_FUSED=0: cuFFT path)."""
    return n_fft == ops.FFT_FUSED_N and win == n_fft and os.environ.get("GSN_FFT_FUSED", "1") != "0"


def _hann(n_fft, device):
    key = (n_fft, device.index)
    window = _HANN.get(key)
    if window is None:
        window = _HANN[key] = torch.hann_window(n_fft, device=device)
    return window


def _istft(spec, n_fft, hop, win, length):
    # audiozen/acoustics/audio_feature.py:297-347
    window = torch.hann_window(n_fft, device=spec.device)
    return torch.istft(spec, n_fft, hop, win, window=window, length=length)


_ISTFT_ENV = {}


def _istft_nosync(spec, n_fft, hop, win, length):
    """torch.istft(center=True, hann window, normalized=False) without its host-synchronising window-envelope
    check, so it can be enqueued asynchronously / captured in a CUDA graph: inverse real FFT of every frame,
    synthesis window, overlap-add (F.fold), division by the overlap-added squared window (cached per shape),
    removal of the n_fft//2 centre padding.  spec complex [B,F,T] -> [B,length]."""
    assert win == n_fft, "win_length != n_fft is not used by any recipe"
    B, F, T = spec.shape
    window = torch.hann_window(n_fft, device=spec.device)
    frames = torch.fft.irfft(spec.transpose(1, 2), n=n_fft, dim=-1) * window          # [B,T,n_fft]
    full = n_fft + hop * (T - 1)
    y = torch.nn.functional.fold(frames.transpose(1, 2), output_size=(1, full), kernel_size=(1, n_fft),
                                 stride=(1, hop)).reshape(B, full)
    key = (n_fft, hop, T, length, spec.device.index)
    capturing = spec.is_cuda and torch.cuda.is_current_stream_capturing()
    env = _ISTFT_ENV.get(key)
    if env is None or capturing:
        wsq = (window * window).reshape(1, n_fft, 1).expand(1, n_fft, T)
        env = torch.nn.functional.fold(wsq, output_size=(1, full), kernel_size=(1, n_fft),
                                       stride=(1, hop)).reshape(1, full)
        if not capturing:
            _ISTFT_ENV[key] = env
    start = n_fft // 2
    out = y[:, start:start + length] / env[:, start:start + length]
    if out.shape[1] < length:
        out = torch.nn.functional.pad(out, (0, length - out.shape[1]))
    return out


_HANN = {}


def _empty_spec_like(cmp, S):
    """Uninitialised enhanced spectrum [B,S,F,T] in the layout of `cmp` [B,F,T] (torch.stft's result is a transposed
    view of cuFFT's [B,T,F]: keeping that layout end to end avoids every transpose copy up to the inverse FFT)."""
    B, F, T = cmp.shape
    if cmp.is_contiguous():
        return torch.empty((B, S, F, T), dtype=cmp.dtype, device=cmp.device)
    return torch.empty((B, S, T, F), dtype=cmp.dtype, device=cmp.device).transpose(2, 3)


def _merge_speakers(enh):
    """[B,S,F,T] -> [B*S,F,T] without a copy in either layout."""
    B, S, F, T = enh.shape
    if enh.is_contiguous():
        return enh.reshape(B * S, F, T)
    return enh.transpose(2, 3).reshape(B * S, T, F).transpose(1, 2)


def _istft_fused(spec, n_fft, hop, win, length):
    """torch.istft(center=True, hann window, normalized=False), asynchronous and graph-capturable: cuFFT inverse real
    FFT of every frame, then ONE kernel (gsn_overlap_add) for synthesis window, overlap-add, division by the
    overlap-added squared window and removal of the centre padding (audio_feature.py:297-347).
    spec complex [B,F,T] -> [B,length]."""
    assert win == n_fft, "win_length != n_fft is not used by any recipe"
    window = _hann(n_fft, spec.device)
    if _fft_kernels(n_fft, win) and spec.shape[1] == n_fft // 2 + 1 and spec.transpose(1, 2).is_contiguous():
        frames = ops.irfft_frames(spec)  # gsn_irfft_frames: no defensive clone, no separate scaling pass
    else:
        frames = torch.fft.irfft(spec.transpose(1, 2), n=n_fft, dim=-1).contiguous()  # [B,T,n_fft]; no copy when time-major
    return ops.overlap_add(frames, window, hop, length)


def _fused_back_end(projs, cmp, Ns, ctrs, dfs, layout, n_fft, hop, win, length, want_mag=True):
    """Deep filter + pass-through + inverse FFT in one kernel (gsn_deepfilter_irfft), then the overlap-add: the back end
    of forward() for ONE speaker at the recipes' transform length.  Returns (y [B,length], |enh| [B,F,T]) or None when
    the fused kernel does not apply (the callers then take the per-band kernels + cuFFT)."""
    if not _fft_kernels(n_fft, win) or len(projs) > 4 or cmp.shape[1] != n_fft // 2 + 1:
        return None
    if cmp.is_contiguous() or not cmp.transpose(1, 2).is_contiguous():
        return None
    frames, mag, _ = ops.deepfilter_irfft([p.contiguous() for p in projs], cmp, Ns, ctrs, dfs, layout=layout,
                                          want_mag=want_mag)
    y = ops.overlap_add(frames, _hann(n_fft, cmp.device), hop, length)
    return y, (mag[:, 0] if mag is not None else None)


class _GraphedNetwork:
    """CUDA-graph replay of `network()` for the drop-in models: `_network(mag)` is the eager launch sequence,
    `_network_sched(mag)` what gets captured (subclasses may substitute a different schedule)."""

    use_cuda_graph = False
    frame_chunks = 1

    def enable_cuda_graph(self, flag=True, frame_chunks=8):
        """Replay the hot path from a CUDA graph captured per input shape (the returned tensors are the graph's
        static output buffers, OVERWRITTEN by the next call with the same shape).  `frame_chunks` > 1 selects the
        frame-chunked wavefront schedule where the model supports it (SpikingFullSubNet)."""
        self.use_cuda_graph = bool(flag)
        self.frame_chunks = max(1, int(frame_chunks))
        self._graphs = {}
        return self

    def _network_sched(self, mag):
        return self._network(mag)

    def refresh_folded_bn(self):
        """Bring the cached eval-BatchNorm affines of every cell up to date (in place: captured graphs read them)."""
        cells = self.__dict__.get("_gsu_cells")
        if cells is None:
            cells = self.__dict__["_gsu_cells"] = [m for m in self.modules() if isinstance(m, GSUCell) and m.use_bn]
        for c in cells:
            c.folded_bn()

    def _warn_cell_hooks(self):
        """The reference's debug mode (audiozen/trainer.py:354-356, DebugUnderflowOverflow) hooks every sub-module; the
        frame loop lives inside the kernels here, so `GSUCell.forward` / `GSULayer.forward` are never called per frame
        and their hooks cannot fire.  Say so once instead of letting a debugging session believe it saw every frame; the
        per-layer spike traces are in the returned `all_layer_outputs`, hooks on the top-level module do fire."""
        if self.__dict__.get("_cell_hooks_checked"):
            return
        self.__dict__["_cell_hooks_checked"] = True
        hooked = [n for n, m in self.named_modules()
                  if isinstance(m, (GSUCell, GSULayer, StackedGSU)) and (m._forward_hooks or m._forward_pre_hooks)]
        if hooked:
            import warnings
            warnings.warn(f"spiking_fullsubnet_b200: forward hooks on {len(hooked)} GSN sub-modules (e.g. {hooked[0]!r}) are "
                          "not called: the per-frame loop runs inside the CUDA kernels.  Hook the top-level module, or "
                          "read the per-layer traces the model returns.", stacklevel=3)

    def network(self, mag):
        if not mag.is_cuda:
            raise RuntimeError("spiking_fullsubnet_b200 has no CPU path: move the model and input to CUDA")
        self._warn_cell_hooks()
        if torch.cuda.is_current_stream_capturing():
            return self._network_sched(mag) if self.use_cuda_graph else self._network(mag)
        if not self.use_cuda_graph:
            if getattr(self, "streaming", False):
                res = self._network_stream(mag)
                if res is not None:
                    return res
            return self._network(mag)
        graphs = self.__dict__.setdefault("_graphs", {})
        key = ("network", tuple(mag.shape), mag.dtype, mag.device.index)
        entry = graphs.get(key)
        if entry is None:
            # keeps the strides of a complex STFT (it may be time-major, ops._spec_layout); a real magnitude is made
            # contiguous by every schedule before it is compressed
            static_in = torch.empty_like(mag) if mag.is_complex() else torch.empty(mag.shape, dtype=mag.dtype, device=mag.device)
            static_in.copy_(mag)
            with torch.no_grad():
                self._network(static_in)  # warm-up outside the capture (lazy CUDA initialisation)
                torch.cuda.synchronize(mag.device)
                # The compression of the input (the first kernel of every schedule) stays outside the graph and runs on
                # the caller's tensor in front of each replay: no copy of the input into a static buffer.  `hook`
                # collects what ops.compress_mag was asked for during the capture.
                hook = {"args": None, "cm": None, "bad": False}
                if os.environ.get("GSN_GRAPH_INPUT_COPY", "0") != "1":
                    static_in._gsn_cm_capture = hook
                graph = torch.cuda.CUDAGraph()
                n0 = ops.LAUNCHES[0]
                with torch.cuda.graph(graph):
                    static_out = self._network_sched(static_in)
                if hasattr(static_in, "_gsn_cm_capture"):
                    del static_in._gsn_cm_capture
                if hook["bad"]:  # not a schedule the shortcut knows: capture again with the compression inside
                    hook = {"args": None, "cm": None, "bad": False}
                    graph = torch.cuda.CUDAGraph()
                    n0 = ops.LAUNCHES[0]
                    with torch.cuda.graph(graph):
                        static_out = self._network_sched(static_in)
                eager_cm = hook["args"] is not None
                self.graph_launches = ops.LAUNCHES[0] - n0 + int(eager_cm)  # kernels of this library per step
            entry = graphs[key] = (graph, static_in, static_out, hook if eager_cm else None)
        graph, static_in, static_out, hook = entry
        self.refresh_folded_bn()
        if hook is not None:
            ops.compress_mag(mag if mag.is_complex() else mag.contiguous(), hook["args"][0], hook["args"][1], out=hook["cm"])
        else:
            static_in.copy_(mag)
        graph.replay()
        return static_out


class _StreamingPipeline:
    """The streaming schedule shared by surface A and surface B: plan + launch of the persistent stages."""

    # ---- streaming schedule: every stage of every sequence model is ONE persistent kernel for all T frames, chained
    #      to its producers through per-frame counters (gsn_recurrence_stream / gsn_pre_stream /
    #      gsn_linear_spike_bits_stream); ~20 launches per step, no xproj of layers >= 1, bit-packed traces only.
    streaming = False
    strict_outputs = False

    def enable_streaming(self, flag=True, strict_outputs=False):
        """Run the hot path as the frame-granular streaming pipeline where it fits on the device (else the previous
        schedules).  strict_outputs=True also materialises the reference-shaped fp32 traces inside the kernels."""
        self.streaming = bool(flag)
        self.strict_outputs = bool(strict_outputs)
        self._graphs = {}
        return self

    def _stream_models(self, B):
        sb = self.sb_model
        out = [dict(m=self.fb_model, R=B, N=1, lo=0, ctr=self.fb_input_size, nbr=0, fb=False)]
        for i, m in enumerate(sb.sb_models):
            lo, hi, ctr = sb.freq_cutoffs[i], sb.freq_cutoffs[i + 1], sb.center_freq_sizes[i]
            if (hi - lo) % ctr != 0:
                raise ValueError(f"Number of frequency bins must be divisible by the center frequency."
                                 f"GOT: ctr_freq={ctr}, upper_cutoff_freq={hi}, lower_cutoff_freq={lo}")
            N = (hi - lo) // ctr
            out.append(dict(m=m, R=B * N, N=N, lo=lo, ctr=ctr, nbr=sb.neighbor_freq_sizes[i], fb=True))
        return out

    # cost model of the helper stages (measured on B200, tools/stream_stage_timing.py / stream_fused0_check.py):
    #   gsn_xplanes_stream: (0.009 + 0.00028 Kmma) us per row per CTA;
    #   tcgen05 stages: ~60 cycles per MMA at 64-row tiles (3 planes for spike inputs, 8 plane pairs for real inputs)
    _STREAM_TARGET_US = 1.0   # helper stages must be faster than the recurrences' frame time (1.2 - 1.3 us)
    _XOP_RING = 64            # frames of layer-0 operand images kept (ring; S: 64 x 202 KB = 13 MB, L2-resident)

    @staticmethod
    def _xplanes_us(R, K):
        kmma = (K + 15) // 16 * 16
        return R * (0.009 + 0.00028 * kmma)

    @staticmethod
    def _stage_us(R, K, passes):
        # per 64-row tile: 60 cycles per 128x64x16 MMA + 0.15 us of hand-overs (tools/stage_k_timing.py: R = 96 / 256 at
        # K = 160 1.71 / 3.90 us per frame, R = 192 / 128 at K = 256 4.4 - 5.0 / 3.3 us)
        kmma = (K + 15) // 16 * 16
        return (R / 64.0) * (passes * (kmma // 16) * 60.0 / 1965.0 + 0.15)

    @classmethod
    def _xplanes_ctas(cls, R, K, target):
        return max(1, int(math.ceil(cls._xplanes_us(R, K) / target)))

    @classmethod
    def _stage_ctas(cls, R, K, passes, target):
        return max(1, int(math.ceil(cls._stage_us(R, K, passes) / target)))

    def _stream_plan(self, B, sm_total=None, scale=1.0):
        """Plan of the whole network at batch B (None: not co-resident).  scale > 1 sizes the helper stages for a
        proportionally longer frame time (fewer helper CTAs: more utterances per wave)."""
        return self._stream_plan_for(self._stream_models(B), sm_total, scale)

    def _stream_plan_for(self, models, sm_total=None, scale=1.0):
        """Stage list with CTA counts, or None when the pipeline cannot be co-resident (all kernels spin on each
        other's counters, so every CTA of every stage must be resident at once: one CTA per SM)."""
        if sm_total is None:
            sms = torch.cuda.get_device_properties(torch.cuda.current_device()).multi_processor_count \
                if torch.cuda.is_available() else 148
            # A thread-block cluster needs all its SMs free inside ONE GPC, the single-CTA helper stages run anywhere;
            # when helper CTAs are placed first they can leave a GPC with an odd SM free, and a cluster still waiting
            # would never start (every resident kernel spins on its producers).  One spare SM per GPC (8) rules that
            # out for 2-CTA clusters whatever the launch order; wider clusters additionally run on higher-priority
            # streams (_stream_run).  (A gate kernel that holds the helpers back until all clusters are resident was
            # tried and dropped: with streams aliased onto few hardware queues it can block a cluster behind itself.)
            sm_total = int(os.environ.get("GSN_STREAM_SMS", sms - 8))
        # helper stages are sized for (a bit less than) the recurrences' frame time, which grows with the hidden size:
        # 1.15 us at H = 160, 1.36 us at H = 256 / 320 (xproj input), so wide models leave more SMs to the recurrences
        h_max = max(d["m"].hidden_size for d in models)
        target = scale * float(os.environ.get("GSN_STREAM_TARGET_US",
                                              self._STREAM_TARGET_US * (1.0 + 0.001 * max(0, h_max - 160))))
        helpers = 0
        for d in models:
            m = d["m"]
            cells = [l.cell for l in m.sequence_model.layers]
            if any(not c.shared_weights for c in cells) or any(c.use_bn and c.batchnorm.training for c in cells):
                return None
            H, K, R = m.hidden_size, m.input_size, d["R"]
            if not isinstance(m.proj, nn.Linear) or m.proj_size > 2048 or H > 320 or K > 256:
                return None
            C = (H + 127) // 128
            d["C"] = C
            d["fused0"] = ops.stream_ctas(R, H, K, True) > 0
            d["layers"] = [dict(fused=(l > 0 and ops.stream_ctas(R, H, H, True) > 0)) for l in range(len(cells))]
            if d["fused0"]:
                d["pre_p"] = self._xplanes_ctas(R, K, target)
                helpers += d["pre_p"]
            else:
                if not _lib_supported_pre(K, H) or d.get("needs_div"):
                    return None  # (the separate front end has no row divisor: surface B needs the fused layer 0)
                d["pre_p"] = self._stage_ctas(R, K, 8, target)
                helpers += C * d["pre_p"]
            d["lin_p"] = self._stage_ctas(R, H, 3, target)
            d["proj_p"] = self._stage_ctas(R, H, 3, target)
            helpers += sum(0 if (ly["fused"] or i == 0) else C * d["lin_p"] for i, ly in enumerate(d["layers"]))
            helpers += ((m.proj_size + 127) // 128) * d["proj_p"]
        # only with the finest row tile: 32- and 64-row frames cost 2.9 / 5.1 us, the chunked wavefront is faster there
        nt = 16
        rec = sum(((d["R"] + nt - 1) // nt) * d["C"] * len(d["layers"]) for d in models)
        if rec + helpers > sm_total:
            return None
        # SMs the plan leaves idle go to the helper stage with the longest frame time, one CTA (or one CTA per
        # slice) at a time: the pipeline runs at the pace of its slowest stage, and the cost model above is only
        # good to ~15 % (S: 0.83 -> 0.78 ms per step)
        spare = sm_total - rec - helpers
        cand = []  # [frame time of one CTA (us), plan entry, key, SMs per step]
        for d in models:
            m = d["m"]
            H, K, R, C = m.hidden_size, m.input_size, d["R"], d["C"]
            cand.append([self._xplanes_us(R, K) if d["fused0"] else self._stage_us(R, K, 8), d, "pre_p",
                         1 if d["fused0"] else C])
            if any(not (ly["fused"] or i == 0) for i, ly in enumerate(d["layers"])):
                cand.append([self._stage_us(R, H, 3), d, "lin_p",
                             C * sum(0 if (ly["fused"] or i == 0) else 1 for i, ly in enumerate(d["layers"]))])
            cand.append([self._stage_us(R, H, 3), d, "proj_p", (m.proj_size + 127) // 128])
        while spare > 0:
            cand.sort(key=lambda c: -c[0] / c[1][c[2]])
            pick = next((c for c in cand if c[3] <= spare), None)
            if pick is None or pick[0] / pick[1][pick[2]] < 0.25:
                break
            pick[1][pick[2]] += 1
            spare -= pick[3]
        for d in models:
            d["nt"] = nt
        return models

    def _stream_frame_us(self, B):
        """Frame time of the recurrences (measured): 1.2 us up to H = 240, 1.46 us for the wide models in the pipeline."""
        return 1.2 if max(d["m"].hidden_size for d in self._stream_models(B)) <= 240 else 1.46

    def _stream_wave_plan(self, B, T=501):
        """(utterances per wave, helper scale, estimated us) of the cheapest wave schedule, or None: every wave costs T
        frame times whatever its size, so fewer waves win even when that means slower helper stages (L: 8 waves with the
        helpers sized for 1.5 us beat 10 waves at 1.16 us)."""
        cache = self.__dict__.setdefault("_wave_cache", {})
        if (B, T) not in cache:
            best = None
            frame_us = self._stream_frame_us(B)
            h_max = max(d["m"].hidden_size for d in self._stream_models(B))
            base = self._STREAM_TARGET_US * (1.0 + 0.001 * max(0, h_max - 160))
            for scale in (1.0, 1.15, 1.3, 1.5):
                b = next((b for b in range(B, 0, -1) if self._stream_plan(b, scale=scale) is not None), None)
                if b is None:
                    continue
                us = -(-B // b) * (T * max(frame_us, base * scale) + 120.0)
                if best is None or us < best[2] - 1e-9:
                    best = (b, scale, us)
            cache[(B, T)] = best
        return cache[(B, T)]

    def _stream_wave_size(self, B):
        """Utterances per wave of the cheapest wave schedule (None: not even one utterance is co-resident)."""
        plan = self._stream_wave_plan(B)
        return None if plan is None else plan[0]

    def _waves_pay(self, B, T):
        """Waves are latency-bound (every wave costs T frame times whatever its size), the wavefront / band-stream
        schedules throughput-bound (measured 0.24 - 0.33 of the 6.3 row-frames per us per SM the recurrence sustains):
        S at batch 64 = 2 waves 1.5 ms vs 4.1 ms, L at 64 x 10 s = 8 waves 15.1 vs 22.6 ms, but M at batch 32 = 3 waves
        2.1 vs 1.7 ms.  Estimate both and take the smaller."""
        plan = self._stream_wave_plan(B, T)
        if plan is None:
            return False
        mode = os.environ.get("GSN_STREAM_WAVES", "auto")
        if mode in ("0", "1"):
            return mode == "1"
        models = self._stream_models(B)
        waves_us = plan[2]
        row_frames = sum(d["R"] * len(d["m"].sequence_model.layers) for d in models) * T
        return waves_us < 0.8 * row_frames / (6.3 * 148 * 0.28)  # (the estimate is only good to ~20 %: M is a tie)

    def _network_stream_waves(self, mag):
        """Batches whose pipeline does not fit on the device at once (L at batch 64: 688 recurrence CTAs) run it in
        WAVES of as many utterances as do fit (utterances are independent, MSF:155): the same persistent kernels, one
        co-resident pipeline after the other, results concatenated along the rows."""
        dev = mag.device
        B, F, T = mag.shape
        if not self._waves_pay(B, T):
            return None
        b, scale, _ = self._stream_wave_plan(B, T)
        fbm = self.fb_model
        rep = (self.n_fft // 2 + 1) // self.fb_input_size
        if rep * fbm.proj_size < F - 1:
            raise ValueError(f"full-band output ({fbm.proj_size} bins x {rep}) does not cover {F - 1} bins")
        ops.stream_preload(dev)
        cm = ops.compress_mag(mag if mag.is_complex() else mag.contiguous(), F - 1, self.fdrc)
        waves, launches = [], []
        for w, lo in enumerate(range(0, B, b)):
            hi = min(B, lo + b)
            waves.append(self._stream_run(self._stream_plan(hi - lo, scale=scale), cm[:, lo:hi].contiguous(), tag=f"a{w}"))
            if self.__dict__.get("record_stream_launches"):
                launches += self.stream_launches
        if launches:
            self.stream_launches = launches
        nmod = len(waves[0])

        def merged(k):  # model k over all waves: (proj, all_layer_outputs)
            parts = [wv[k] for wv in waves]
            proj = torch.cat([p[0] for p in parts], dim=1)
            louts = [p[1] for p in parts]
            n = len(louts[0])
            thunks = [(lambda i=i: torch.cat([lo_[i] for lo_ in louts], dim=1)) for i in range(n - 1)] + [proj]
            counts = louts[0].spike_counts
            for lo_ in louts[1:]:
                counts = counts + lo_.spike_counts
            return proj, LazyOutputs(thunks, widths=louts[0].widths, spike_counts=counts,
                                     trace_numel=sum(lo_.trace_numel for lo_ in louts))

        out = [merged(k) for k in range(nmod)]
        # bit-packed traces of the whole batch, per model and layer (rows are batch-major: waves concatenate)
        self.last_spike_bits = [[torch.cat([wv[k][2][l] for wv in waves], dim=1) for l in range(len(waves[0][k][2]))]
                                for k in range(nmod)]
        self.stream_waves = (b, len(waves))
        return [o[0] for o in out[1:]], out[0][1], [o[1] for o in out[1:]]

    def _network_stream(self, mag):
        dev = mag.device
        B, F, T = mag.shape
        models = self._stream_plan(B)
        if models is None:
            return self._network_stream_waves(mag)
        self.stream_waves = (B, 1)
        fbm = self.fb_model
        rep = (self.n_fft // 2 + 1) // self.fb_input_size
        if rep * fbm.proj_size < F - 1:
            raise ValueError(f"full-band output ({fbm.proj_size} bins x {rep}) does not cover {F - 1} bins")
        ops.stream_preload(dev)
        cm = ops.compress_mag(mag if mag.is_complex() else mag.contiguous(), F - 1, self.fdrc)
        results = self._stream_run(models, cm, tag="a")
        return [r[0] for r in results[1:]], results[0][1], [r[1] for r in results[1:]]

    def _stream_run(self, models, cm, fb_act=None, tag=""):
        """Launch the streaming pipeline of `models` (a plan of _stream_plan_for) on the compressed magnitude cm
        [T,B,F]: every stage on its own stream, forked from and joined to the current one.  A model with d["fb"] reads
        the full-band output: of the plan's own full-band model (frame by frame, through its proj stage's counters) or
        `fb_act` [T,B,f_fb] when that is already complete.  d["div"]: per-row divisor of the layer-0 input (surface
        B's norms).  Returns [(proj, all_layer_outputs, bits)] per model."""
        dev = cm.device
        T, B, _ = cm.shape
        strict = self.strict_outputs
        main = torch.cuda.current_stream(dev)
        f32 = dict(device=dev, dtype=torch.float32)
        ncnt = sum(2 + 2 * len(d["layers"]) for d in models)
        counters = ops.frame_counters(T, dev, ncnt)
        nxt = iter(range(ncnt))
        nrec = sum(len(d["layers"]) for d in models)
        spike_counts = torch.zeros(nrec, device=dev, dtype=torch.int64)  # spikes emitted per (model, layer) launch
        rec_idx = 0
        streams = _band_streams(dev, ncnt, priority=0, tag=("stream_pipeline", tag))
        st_it = iter(streams)
        # clusters of three or more CTAs need that many free SMs inside ONE GPC: when the device is (nearly) full they must
        # be placed before the 1- and 2-CTA kernels have scattered over the GPCs (L, H = 320: the full-band recurrences
        # started 330 us late, after a helper stage had finished and freed its SMs) -> a higher stream priority
        st_hi = iter(_band_streams(dev, nrec, priority=-1, tag=("stream_pipeline_hi", tag)))
        # operand-image buffers of the fused layer-0 path: zero-filled ONCE (padding rows), on the main stream and
        # before the fork, then reused by every call of this shape
        keep = self.__dict__.setdefault("_xop_cache", {})
        for mi, d in enumerate(models):
            if d["fused0"]:
                m = d["m"]
                budget = ((d["R"] + d["nt"] - 1) // d["nt"]) * d["C"]
                nt0 = ops.stream_tile(d["R"], m.hidden_size, m.input_size, True, budget)
                # a RING of frames (the layer-0 recurrence's out counters are the producer's back-pressure): the images
                # are consumed microseconds after they are written and never need to reach DRAM
                ring = min(T, int(os.environ.get("GSN_XOP_RING", self._XOP_RING)))
                key = (tag, mi, ring, d["R"], m.input_size, nt0, dev.index)
                if key not in keep:
                    keep[key] = ops.xplanes_buffer(ring, d["R"], m.input_size, nt0, dev)
                d["xop"], d["nt0"], d["ring"] = keep[key], nt0, ring
            # spike operand images between fused layers (layer l-1 writes, layer l fetches with one bulk copy per frame
            # instead of expanding bits in its loader warp): rings like the layer-0 images, L2-resident
            d["img"] = {}
            if d["nt"] == 16 and os.environ.get("GSN_STREAM_IMAGES", "1") != "0":
                iring = min(T, int(os.environ.get("GSN_XOP_RING", self._XOP_RING)))
                for l, ly in enumerate(d["layers"]):
                    if l > 0 and ly["fused"]:
                        key = (tag, mi, "img", l, iring, d["R"], d["m"].hidden_size, dev.index)
                        if key not in keep:
                            keep[key] = ops.spike_image_buffer(iring, d["R"], d["m"].hidden_size, dev)
                        d["img"][l] = (keep[key], iring)
        # folded BatchNorm affines are (re)computed by torch kernels on the main stream while a graph is captured: they
        # must precede the fork too.  `hold` keeps every buffer of this call alive until the next one, so that the
        # caching allocator cannot hand a block to a later allocation while a concurrently running stage still uses it
        hold = []
        for d in models:
            d["bn"] = [l.cell.folded_bn() for l in d["m"].sequence_model.layers]
            hold.append(d["bn"])
        fork = torch.cuda.Event()
        fork.record(main)
        used = []

        def on_stream(fn, wide_cluster=False):
            stq = next(st_hi) if wide_cluster else next(st_it)
            stq.wait_event(fork)
            with torch.cuda.stream(stq):
                fn()
            used.append(stq)

        def ln(m):
            if not m.use_pre_layer_norm:
                return None, None, 1e-5
            return m.pre_layer_norm.weight.detach(), m.pre_layer_norm.bias.detach(), m.pre_layer_norm.eps

        fb_cnt = None
        fb_target = 0
        all_helpers = []  # enqueued after the recurrences of ALL models: clusters are placed while whole GPCs are free
        results = []
        record = [] if self.__dict__.get("record_stream_launches") else None
        # The recurrences (thread-block clusters) are enqueued before the helper stages of their model so that cluster
        # placement is not fragmented by single-CTA kernels; consumers spin on their producers' frame counters.
        for mi, d in enumerate(models):
            m, R = d["m"], d["R"]
            H, K, C = m.hidden_size, m.input_size, d["C"]
            cells = [l.cell for l in m.sequence_model.layers]
            nt = d["nt"]
            budget = ((R + nt - 1) // nt) * C  # makes the tile picker choose nt
            w_ln, b_ln, e_ln = ln(m)
            x_out = torch.empty((T, R, K), **f32) if strict else None
            geo = (d["N"], d["lo"], d["ctr"], d["nbr"])
            fbt = fb_act if d["fb"] else None
            c_pre = counters[next(nxt)]
            c_l0 = counters[next(nxt)]  # out counters of the layer-0 recurrence
            pre_in = dict(in_cnt=fb_cnt if d["fb"] else None, in_target=fb_target)
            w_ih0 = cells[0].weight_ih.detach()
            if d["fused0"]:
                xop, nt0 = d["xop"], d["nt0"]
                xproj = None
                bp = dict(ring=d["ring"], bp_cnt=c_l0, bp_target=ops.stream_ctas(R, H, K, True, budget),
                          row_div=d.get("div"))
                pre = (lambda geo=geo, fbt=fbt, nt0=nt0, xop=xop, w_ln=w_ln, b_ln=b_ln, e_ln=e_ln, x_out=x_out, c_pre=c_pre,
                       pre_in=pre_in, d=d, bp=bp: ops.xplanes_stream(cm, fbt, *geo, nt0, xop, w_ln, b_ln, e_ln, out_x=x_out,
                                                                      out_cnt=c_pre, ctas=d["pre_p"], **pre_in, **bp))
                pre_target = R
            else:
                xop = None
                xproj = torch.empty((T, R, H), **f32)
                hold.append(xproj)
                pre = (lambda geo=geo, fbt=fbt, w_ih0=w_ih0, xproj=xproj, w_ln=w_ln, b_ln=b_ln, e_ln=e_ln, x_out=x_out,
                       c_pre=c_pre, pre_in=pre_in, d=d: ops.pre_stream(cm, fbt, *geo, w_ih0, w_ln, b_ln, e_ln, out_x=x_out,
                                                                       out_xproj=xproj, out_cnt=c_pre,
                                                                       ctas_per_slice=d["pre_p"], **pre_in))
                pre_target = R * C
            in_cnt, in_target = c_pre, pre_target
            bits_prev = None
            c_pending = None  # out counters of the next layer when the current one already needs them (back-pressure)
            bits_all, h_all = [], []
            helpers = [pre]
            for l, (cell, ly) in enumerate(zip(cells, d["layers"])):
                a, b = d["bn"][l]
                bits = ops.spike_bits_buffer((T, R), H, dev)
                h_out = torch.empty((T, R, H), **f32) if strict else None
                c_out = c_l0 if l == 0 else (c_pending if c_pending is not None else counters[next(nxt)])
                kw = dict(out_bits=bits, out_h=h_out, out_cnt=c_out, sm_budget=budget, in_cnt=in_cnt, in_target=in_target,
                          spike_count=spike_counts[rec_idx + l:rec_idx + l + 1])
                if l + 1 in d["img"]:  # my spikes also leave as the operand image of layer l + 1 (its counters = my back-pressure)
                    c_next = counters[next(nxt)]
                    kw.update(img_out=d["img"][l + 1][0], img_ring=d["img"][l + 1][1], bp_cnt=c_next,
                              bp_target=ops.stream_ctas(R, H, H, True, budget))
                else:
                    c_next = None
                w_hh, bias = cell.weight_hh.detach(), cell.bias_ih.detach()
                fused = ly["fused"] or (l == 0 and d["fused0"])
                if record is not None:
                    # the same launch without counters (its inputs are complete once this step has run): timed alone
                    kin = (K if l == 0 else H) if fused else 0
                    ins = (dict(in_planes=xop, w_ih=w_ih0, frames_rows=(T, R), planes_ring=d["ring"]) if (l == 0 and d["fused0"]) else
                           dict(in_bits=bits_prev, w_ih=cell.weight_ih.detach()) if ly["fused"] else None)
                    record.append(dict(model=mi, layer=l, T=T, R=R, H=H, K_in=kin, fused=fused,
                                       flops=2.0 * T * R * H * (H + kin), w_hh=w_hh, bias=bias, a=a, b=b, ins=ins,
                                       out_bits=bits, budget=budget, out_cnt=c_out,
                                       img={k: kw[k] for k in ("img_out", "img_ring", "bp_cnt", "bp_target") if k in kw}))
                if l == 0 and d["fused0"]:
                    on_stream(lambda w_hh=w_hh, bias=bias, a=a, b=b, kw=kw, xop=xop, w_ih0=w_ih0, R=R, ring=d["ring"]:
                              ops.recurrence_stream(w_hh, bias, a, b, in_planes=xop, w_ih=w_ih0, frames_rows=(T, R),
                                                    planes_ring=ring, **kw), wide_cluster=C >= 3)
                elif l == 0:
                    if record is not None:
                        record[-1]["ins"] = dict(xproj=xproj)
                    on_stream(lambda w_hh=w_hh, bias=bias, a=a, b=b, kw=kw, xp=xproj:
                              ops.recurrence_stream(w_hh, bias, a, b, xproj=xp, **kw), wide_cluster=C >= 3)
                elif ly["fused"] and l in d["img"]:
                    if record is not None:
                        record[-1]["ins"] = dict(in_image=d["img"][l][0], planes_ring=d["img"][l][1], frames_rows=(T, R),
                                                 w_ih=cell.weight_ih.detach())
                        record[-1]["scratch_out"] = True  # relaunched alone the ring holds the last frames only
                    on_stream(lambda w_hh=w_hh, bias=bias, a=a, b=b, kw=kw, im=d["img"][l], w=cell.weight_ih.detach():
                              ops.recurrence_stream(w_hh, bias, a, b, in_image=im[0], planes_ring=im[1],
                                                    frames_rows=(T, R), w_ih=w, **kw), wide_cluster=C >= 3)
                elif ly["fused"]:
                    on_stream(lambda w_hh=w_hh, bias=bias, a=a, b=b, kw=kw, bp=bits_prev, w=cell.weight_ih.detach():
                              ops.recurrence_stream(w_hh, bias, a, b, in_bits=bp, w_ih=w, **kw), wide_cluster=C >= 3)
                else:
                    xp = torch.empty((T, R, H), **f32)
                    hold.append(xp)
                    c_lin = counters[next(nxt)]
                    helpers.append(lambda w=cell.weight_ih.detach(), bp=bits_prev, xp=xp, ic=in_cnt, it=in_target,
                                   c_lin=c_lin, d=d: ops.linear_bits_stream(bp, w, out=xp, ctas=C * d["lin_p"], in_cnt=ic,
                                                                            in_target=it, out_cnt=c_lin))
                    kw.update(in_cnt=c_lin, in_target=R * C)
                    if record is not None:
                        record[-1]["ins"] = dict(xproj=xp)
                    on_stream(lambda w_hh=w_hh, bias=bias, a=a, b=b, kw=kw, xp=xp:
                              ops.recurrence_stream(w_hh, bias, a, b, xproj=xp, **kw), wide_cluster=C >= 3)
                in_cnt, in_target = c_out, ops.stream_ctas(R, H, (K if l == 0 else H) if fused else 0, fused, budget)
                bits_prev = bits
                c_pending = c_next
                bits_all.append(bits)
                h_all.append(h_out)
            P = m.proj_size
            proj = torch.empty((T, R, P), **f32)
            act = torch.empty_like(proj) if m._act else proj
            hold.append(act)
            c_proj = counters[next(nxt)]
            pslices = (P + 127) // 128
            helpers.append(lambda m=m, bp=bits_prev, proj=proj, act=act, ic=in_cnt, it=in_target, c_proj=c_proj, d=d:
                           ops.linear_bits_stream(bp, m.proj.weight.detach(), m.proj.bias.detach(), act=m._act, out=proj,
                                                  out_act=act if m._act else None, ctas=pslices * d["proj_p"], in_cnt=ic,
                                                  in_target=it, out_cnt=c_proj))
            all_helpers += helpers
            if not d["fb"]:
                fb_cnt, fb_target, fb_act = c_proj, R * pslices, act

            # all_layer_outputs in the reference's positions; fp32 traces on demand unless strict
            def lazy_x(geo=geo, fbt=fbt, w_ln=w_ln, b_ln=b_ln, e_ln=e_ln, div=d.get("div")):
                x = ops.subband_features(cm, fbt, geo[0], geo[1], geo[2], geo[3], w_ln, b_ln, e_ln)
                if div is None:
                    return x
                if div.dim() == 1:  # per utterance
                    return (x.view(T, B, -1) / div.view(1, B, 1)).view_as(x)
                return x / div.unsqueeze(-1)

            entries = [x_out if strict else lazy_x]
            for bits, h_out in zip(bits_all, h_all):
                entries.append(h_out if strict else (lambda bits=bits, H=H: ops.unpack_spikes(bits, H)))
            entries.append(proj)
            nl = len(cells)
            results.append((proj, LazyOutputs(entries, widths=[K] + [H] * nl + [P],
                                              spike_counts=spike_counts[rec_idx:rec_idx + nl], trace_numel=T * R * H),
                            bits_all, act))
            rec_idx += nl
        for fn in all_helpers:
            on_stream(fn)
        for stq in used:
            done = torch.cuda.Event()
            done.record(stq)
            main.wait_event(done)
        if record is not None:
            self.stream_launches = record
        self.__dict__.setdefault("_stream_keepalive", {})[tag] = (cm, counters, results, hold, spike_counts, fb_act)
        bits = [r[2] for r in results]
        # spike bits of the whole network in model order; a two-phase caller (surface B) runs "b1" (full band) then "b2"
        self.last_spike_bits = self.__dict__.get("last_spike_bits", []) + bits if tag == "b2" else bits
        return results


class SpikingFullSubNet(_StreamingPipeline, _GraphedNetwork, nn.Module):
    """Surface A (MSF:349-474).  forward(wave [B,L]) ->
    (enh_y [B,L], enh_mag [B,F,T], fb_all_layer_outputs, sb_all_layer_outputs), or for num_spks > 1
    (enh_y [B,S,L], fb_all_layer_outputs, sb_all_layer_outputs)."""

    def __init__(self, n_fft, hop_length, win_length, fdrc, fb_input_size, fb_hidden_size, fb_num_layers,
                 fb_proj_size, fb_output_activate_function, sb_hidden_size, sb_num_layers, freq_cutoffs,
                 df_orders, center_freq_sizes, neighbor_freq_sizes, use_pre_layer_norm_fb=True,
                 use_pre_layer_norm_sb=True, bn=False, shared_weights=False, sequence_model="GSN", num_spks=1):
        super().__init__()
        self.fb_model = SequenceModel(input_size=fb_input_size, hidden_size=fb_hidden_size,
                                      num_layers=fb_num_layers, shared_weights=shared_weights,
                                      sequence_model=sequence_model, proj_size=fb_proj_size,
                                      output_activate_function=fb_output_activate_function, bn=bn,
                                      use_pre_layer_norm=use_pre_layer_norm_fb)
        self.sb_model = SubbandModel(freq_cutoffs=freq_cutoffs, center_freq_sizes=center_freq_sizes,
                                     neighbor_freq_sizes=neighbor_freq_sizes, df_orders=df_orders,
                                     num_spks=num_spks, hidden_size=sb_hidden_size, num_layers=sb_num_layers,
                                     shared_weights=shared_weights, sequence_model=sequence_model, bn=bn,
                                     use_pre_layer_norm=use_pre_layer_norm_sb)
        self.subband_model = None
        self.fb_input_size, self.n_fft, self.hop_length, self.win_length = fb_input_size, n_fft, hop_length, win_length
        self.fdrc, self.df_orders, self.num_spks = fdrc, df_orders, num_spks
        self.use_cuda_graph = False
        self.frame_chunks = 1
        self._graphs = {}

    def set_backend(self, backend):
        """'auto' | 'simt' | 'tcgen05' for every recurrence of the model."""
        for m in self.modules():
            if isinstance(m, StackedGSU):
                m.backend = backend
        return self

    # the hot path: magnitude in -> sub-band proj outputs (the coefficients) out  (MSF:434-447).
    # network(mag [B, n_fft//2+1, T]) -> (projs: list of [T, B*N_i, P_i], fb_all, sb_all); see _GraphedNetwork.
    def _network(self, mag):
        F = mag.shape[1]
        cm = ops.compress_mag(mag if mag.is_complex() else mag.contiguous(), F - 1, self.fdrc)  # drops the last bin, MSF:436
        fbm = self.fb_model
        lnw = fbm.pre_layer_norm.weight.detach() if fbm.use_pre_layer_norm else None
        lnb = fbm.pre_layer_norm.bias.detach() if fbm.use_pre_layer_norm else None
        eps = fbm.pre_layer_norm.eps if fbm.use_pre_layer_norm else 1e-5
        x = ops.subband_features(cm, None, 1, 0, self.fb_input_size, 0, lnw, lnb, eps)
        _, fb_act, fb_all = fbm.run_time_major(x)
        # the reference tiles the full-band output (n_fft//2+1)//fb_input_size times (MSF:443); the
        # gather indexes it modulo its width instead, which needs the tiling to cover all bins
        rep = (self.n_fft // 2 + 1) // self.fb_input_size
        if rep * fb_act.shape[2] < F - 1:
            raise ValueError(f"full-band output ({fb_act.shape[2]} bins x {rep}) does not cover {F - 1} bins")
        projs, sb_all = self.sb_model.run_time_major(cm, fb_act)
        return projs, fb_all, sb_all

    def _network_sched(self, mag):
        """The schedule a captured graph replays: the streaming pipeline when enabled and co-resident, else the
        frame-chunked wavefront when it applies, else band streams."""
        if self.streaming:
            res = self._network_stream(mag)
            if res is not None:
                return res
        if self.frame_chunks > 1 and self._fits_wavefront(mag.shape[0]):
            return self._network_wavefront(mag, self.frame_chunks)
        return self._network(mag)

    def _fits_wavefront(self, B):
        """The wavefront schedule keeps every (model, layer) recurrence resident at once; it only pays off
        when all of them fit on the 148 SMs at the finest row tiling (measured: with coarser tiles, L and XL at
        batch 32 are slower in the wavefront than with band streams)."""
        sb = self.sb_model
        total = 0
        for m, rows in [(self.fb_model, B)] + [
                (m, B * ((sb.freq_cutoffs[i + 1] - sb.freq_cutoffs[i]) // sb.center_freq_sizes[i]))
                for i, m in enumerate(sb.sb_models)]:
            cell = m.sequence_model.layers[0].cell
            total += _cluster_ctas(rows, cell.hidden_size, cell.shared_weights) * m.num_layers
        return total <= 148

    def _network_wavefront(self, mag, nchunks):
        """Same results as `_network`, scheduled as a frame-chunked wavefront over one stream per
        (sequence model, layer); meant to be captured into a CUDA graph (see enable_cuda_graph)."""
        dev = mag.device
        B, F, T = mag.shape
        bounds = _chunk_bounds(T, nchunks)
        nchunks = len(bounds)
        main = torch.cuda.current_stream(dev)
        fbm, sbm = self.fb_model, self.sb_model
        rep = (self.n_fft // 2 + 1) // self.fb_input_size
        if rep * fbm.proj_size < F - 1 and isinstance(fbm.proj, nn.Linear):
            raise ValueError(f"full-band output ({fbm.proj_size} bins x {rep}) does not cover {F - 1} bins")
        cm = ops.compress_mag(mag if mag.is_complex() else mag.contiguous(), F - 1, self.fdrc)
        backend = fbm.sequence_model.backend
        fbp = _SeqPlan(fbm, T, B, nchunks, dev, backend)
        geo, sbps = [], []
        for i, m in enumerate(sbm.sb_models):
            lo, hi, ctr = sbm.freq_cutoffs[i], sbm.freq_cutoffs[i + 1], sbm.center_freq_sizes[i]
            if (hi - lo) % ctr != 0:
                raise ValueError(f"Number of frequency bins must be divisible by the center frequency."
                                 f"GOT: ctr_freq={ctr}, upper_cutoff_freq={hi}, lower_cutoff_freq={lo}")
            geo.append(((hi - lo) // ctr, lo, ctr, sbm.neighbor_freq_sizes[i]))
            sbps.append(_SeqPlan(m, T, B * geo[-1][0], nchunks, dev, m.sequence_model.backend))

        # every (model, layer) recurrence runs concurrently in the wavefront: split the SMs between them
        plans_all = [fbp] + sbps
        demands = []
        for p_ in plans_all:
            cell0 = p_.m.sequence_model.layers[0].cell
            demands += [_cluster_ctas(p_.x.shape[1], cell0.hidden_size, cell0.shared_weights)] * p_.L
        bud = _sm_budgets(demands, total=int(os.environ.get("GSN_WF_SM_TOTAL", 148)))
        # SMs left over by the resident recurrence CTAs are what the (persistent) tcgen05 linears can get
        free = max(8, 148 - sum(min(d, b) if b else d for d, b in zip(demands, bud)))
        free = int(os.environ.get("GSN_WF_LIN_BUDGET", free))
        o_ = 0
        for p_ in plans_all:
            p_.sm_budget = bud[o_]
            p_.lin_budget = max(8, free)
            o_ += p_.L

        def ln(m):
            if not m.use_pre_layer_norm:
                return None, None, 1e-5
            return m.pre_layer_norm.weight.detach(), m.pre_layer_norm.bias.detach(), m.pre_layer_norm.eps

        # streams per sequence model: pre[l] (input projections), rec[l] (recurrences), post (proj)
        # (the latency-critical recurrences get high-priority streams so their CTAs are placed first)
        plans = [fbp] + sbps
        # Priorities (lower = placed first): the full-band model is the head of every dependency chain, so its
        # recurrences and the linears between them outrank the sub-band recurrences, which outrank the remaining
        # projections.  GSN_WF_PRIO="fb_rec,fb_lin,sb_rec,sb_lin" overrides (development knob).
        prio = [int(v) for v in os.environ.get("GSN_WF_PRIO", "-3,-2,-1,0").split(",")]
        groups, streams = [], []
        for pi, p in enumerate(plans):
            p_rec, p_lin = (prio[0], prio[1]) if pi == 0 else (prio[2], prio[3])
            lin = _band_streams(dev, p.L + 1, priority=p_lin, tag=("wf_lin", pi))
            rec = _band_streams(dev, p.L, priority=p_rec, tag=("wf_rec", pi))
            groups.append((lin[:p.L], rec, lin[p.L]))
            streams += lin + rec
        fork = torch.cuda.Event()
        fork.record(main)
        for st in streams:
            st.wait_event(fork)

        def staged(stream, waits, fn):
            for w in waits:
                if w is not None:
                    stream.wait_event(w)
            with torch.cuda.stream(stream):
                fn()
                ev = torch.cuda.Event()
                ev.record(stream)
            return ev

        pdl = os.environ.get("GSN_WF_PDL", "1") != "0"

        def run_model(p, group, k, t0, t1, feed, feed_ev):
            """chunk k of one sequence model; `feed` fills p.x[t0:t1] (runs on the layer-0 pre stream)."""
            pre, rec, post = group
            ev = None
            for l in range(p.L):
                def do_pre(l=l):
                    if l == 0:
                        feed()
                    p.run_pre(l, k, t0, t1)
                e_pre = staged(pre[l], [feed_ev if l == 0 else ev], do_pre)
                ev = staged(rec[l], [e_pre], lambda l=l: p.run_rec(l, k, t0, t1, pdl=pdl and k > 0))
            return staged(post, [ev], lambda: p.run_post(k, t0, t1))

        w_fb, b_fb, e_fb = ln(fbm)
        # the gather and fp32 projection kernels get a capped grid here (see GSN_OPT_F32_MAX_CTAS in gsn_b200.h)
        f32_cap = int(os.environ.get("GSN_WF_F32_CTAS", "0"))
        ops.set_option(ops.OPT_F32_MAX_CTAS, f32_cap)
        try:
            self._wavefront_enqueue(bounds, run_model, fbp, sbps, geo, groups, cm, w_fb, b_fb, e_fb, ln)
        finally:
            ops.set_option(ops.OPT_F32_MAX_CTAS, 0)
        for st in streams:
            done = torch.cuda.Event()
            done.record(st)
            main.wait_event(done)
        _, _, fb_all = fbp.outputs()
        outs = [p.outputs() for p in sbps]
        self._keepalive = (cm, fbp, sbps)  # buffers referenced by the captured graph
        return [o[0] for o in outs], fb_all, [o[2] for o in outs]

    def _wavefront_enqueue(self, bounds, run_model, fbp, sbps, geo, groups, cm, w_fb, b_fb, e_fb, ln):
        for k, (t0, t1) in enumerate(bounds):
            ev_fb = run_model(fbp, groups[0], k, t0, t1,
                              lambda: ops.subband_features(cm[t0:t1], None, 1, 0, self.fb_input_size, 0, w_fb,
                                                           b_fb, e_fb, out=fbp.x[t0:t1]), None)
            for bi, (p, (N, lo, ctr, nbr)) in enumerate(zip(sbps, geo)):
                w_sb, b_sb, e_sb = ln(p.m)
                run_model(p, groups[1 + bi], k, t0, t1,
                          lambda p=p, N=N, lo=lo, ctr=ctr, nbr=nbr, w_sb=w_sb, b_sb=b_sb, e_sb=e_sb:
                          ops.subband_features(cm[t0:t1], fbp.act[t0:t1], N, lo, ctr, nbr, w_sb, b_sb, e_sb,
                                               out=p.x[t0:t1]), ev_fb)

    def coefficients(self, mag):
        """Deep-filter coefficient tensors [B, df_i, S, F_i, T, 2] in the reference layout."""
        projs, fb_all, sb_all = self.network(mag)
        B = mag.shape[0]
        return [coef_layout(p, B, p.shape[1] // B, d, self.num_spks)
                for p, d in zip(projs, self.df_orders)], fb_all, sb_all

    def forward(self, input):
        assert input.ndim == 2, f"Input tensor must be 2D, but got {input.ndim}D."
        if not input.is_cuda:
            raise RuntimeError("spiking_fullsubnet_b200 has no CPU path: move the model and input to CUDA")
        if _needs_autograd(self):
            self._warn_cell_hooks()
            from . import training
            return training.spiking_fullsubnet_forward(self, input)
        if not self.use_cuda_graph or torch.cuda.is_current_stream_capturing():
            return self._forward_infer(input)
        # the whole forward() (STFT -> network -> deep filter -> iSTFT) replayed from one CUDA graph per shape
        key = ("forward", tuple(input.shape), input.device.index)
        entry = self._graphs.get(key)
        if entry is None:
            static_in = input.clone()
            with torch.no_grad():
                self._forward_infer(static_in)  # warm-up outside the capture (cuFFT plans, lazy CUDA init)
                torch.cuda.synchronize(input.device)
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph):
                    static_out = self._forward_infer(static_in)
            entry = self._graphs[key] = (graph, static_in, static_out)
        graph, static_in, static_out = entry
        static_in.copy_(input)
        graph.replay()
        return static_out

    def _forward_infer(self, input):
        if self.num_spks == 1 and _fft_kernels(self.n_fft, self.win_length):
            # wave -> (spectrum, compressed magnitude) -> network -> (deep filter + inverse FFT) -> overlap-add: the
            # enhanced spectrum never exists in memory
            F = self.n_fft // 2 + 1
            cmp = _stft_fused(input, self.n_fft, self.hop_length, self.win_length, f_keep=F - 1, fdrc=self.fdrc)
            projs, fb_all, sb_all = self.network(cmp)
            cuts, ctrs = self.sb_model.freq_cutoffs, self.sb_model.center_freq_sizes
            Ns = [(cuts[i + 1] - cuts[i]) // ctrs[i] for i in range(len(projs))]
            res = _fused_back_end(projs, cmp, Ns, ctrs[:len(projs)], self.df_orders[:len(projs)], 0, self.n_fft,
                                  self.hop_length, self.win_length, input.shape[1])
            if res is not None:
                return res[0], res[1], fb_all, sb_all
        enh, mag, fb_all, sb_all = self._forward_spec(input)
        return self._finish(enh, fb_all, sb_all, input.shape[1], mag)

    def _finish(self, enh, fb_all, sb_all, L, mag=None):
        """enh complex [B,S,F,T] -> the reference's return tuple (iSTFT, MSF:463-474)."""
        B, S, F, T = enh.shape
        if S > 1:
            y = _istft_fused(_merge_speakers(enh), self.n_fft, self.hop_length, self.win_length, L)
            return y.reshape(B, S, L), fb_all, sb_all
        enh = enh[:, 0]  # (mag: |enh| straight from the deep-filter kernels of _forward_spec)
        return (_istft_fused(enh, self.n_fft, self.hop_length, self.win_length, L),
                mag[:, 0] if mag is not None else enh.abs(), fb_all, sb_all)

    def _forward_spec(self, input):
        """STFT -> network -> deep filter, all on the complex spectrum (no |stft| / real / imag / repeat /
        torch.complex passes: gsn_compress_spec, gsn_deepfilter_spec, gsn_spec_passthrough)."""
        B, L = input.shape
        cmp = _stft_fused(input, self.n_fft, self.hop_length, self.win_length)  # [B,F,T] complex, time-major view
        projs, fb_all, sb_all = self.network(cmp)
        F, T = cmp.shape[1], cmp.shape[2]
        S = self.num_spks
        enh = _empty_spec_like(cmp, S)
        # enh_mag (MSF:472) is written by the same kernels: no separate |.| pass over the enhanced spectrum
        mag = torch.empty_strided(enh.shape, enh.stride(), dtype=torch.float32, device=enh.device) if S == 1 else None
        cuts, ctrs = self.sb_model.freq_cutoffs, self.sb_model.center_freq_sizes
        lo = 0
        for i, p in enumerate(projs):
            n = (cuts[i + 1] - cuts[i]) // ctrs[i]
            ops.deepfilter_spec(p, cmp, enh, n, ctrs[i], self.df_orders[i], S, lo, mag=mag)
            lo += n * ctrs[i]
        ops.spec_passthrough(cmp, enh, lo, mag=mag)  # un-filtered bins (Nyquist) pass through (MSF:461-468)
        return enh, mag, fb_all, sb_all


class CirmGSN(_GraphedNetwork, nn.Module):
    """cirm_gsn `Model` (CGN:162-244): one full-band GSN over all bins emitting deep-filter coefficients."""

    def __init__(self, n_fft, hop_length, win_length, fdrc, input_size, hidden_size, num_layers, proj_size,
                 output_activate_function, df_order, use_pre_layer_norm_fb=True, bn=False,
                 shared_weights=False, sequence_model="LSTM", num_spks=2):
        super().__init__()
        self.fb_model = SequenceModel(input_size=input_size, hidden_size=hidden_size, num_layers=num_layers,
                                      shared_weights=shared_weights, sequence_model=sequence_model,
                                      proj_size=proj_size * num_spks * df_order * 2,
                                      output_activate_function=output_activate_function, bn=bn,
                                      use_pre_layer_norm=use_pre_layer_norm_fb)
        self.fb_input_size, self.n_fft, self.hop_length, self.win_length = input_size, n_fft, hop_length, win_length
        self.fdrc, self.df_order, self.num_spks = fdrc, df_order, num_spks

    def _network(self, mag):
        """mag [B,F,T] -> (activated proj [T,B,P], all_layer_outputs)."""
        fbm = self.fb_model
        cm = ops.compress_mag(mag if mag.is_complex() else mag.contiguous(), mag.shape[1], self.fdrc)
        lnw = fbm.pre_layer_norm.weight.detach() if fbm.use_pre_layer_norm else None
        lnb = fbm.pre_layer_norm.bias.detach() if fbm.use_pre_layer_norm else None
        x = ops.subband_features(cm, None, 1, 0, self.fb_input_size, 0, lnw, lnb,
                                 fbm.pre_layer_norm.eps if fbm.use_pre_layer_norm else 1e-5)
        _, act, all_out = fbm.run_time_major(x)
        return act, all_out

    def coefficients(self, mag):
        """'b (c d s f) t -> b d s f t c' (CGN:230)."""
        act, all_out = self.network(mag)
        T, B, P = act.shape
        d, S = self.df_order, self.num_spks
        v = act.reshape(T, B, 2, d, S, P // (2 * d * S))
        return v.permute(1, 3, 4, 5, 0, 2).contiguous(), all_out

    def forward(self, input):
        assert input.ndim == 2, f"Input tensor must be 2D, but got {input.ndim}D."
        if not input.is_cuda:
            raise RuntimeError("spiking_fullsubnet_b200 has no CPU path: move the model and input to CUDA")
        if _needs_autograd(self):
            self._warn_cell_hooks()
            from . import training
            return training.cirm_gsn_forward(self, input)
        B, L = input.shape
        cmp = _stft_fused(input, self.n_fft, self.hop_length, self.win_length, f_keep=self.n_fft // 2 + 1,
                          fdrc=self.fdrc)
        act, all_out = self.network(cmp)  # activated proj [T,B,P], features (c d s f) (CGN:230)
        F, T = cmp.shape[1], cmp.shape[2]
        S = self.num_spks
        if S == 1:
            res = _fused_back_end([act.reshape(T, B, -1)], cmp, [1], [F], [self.df_order], 1, self.n_fft,
                                  self.hop_length, self.win_length, L)
            if res is not None:
                return res
        enh = _empty_spec_like(cmp, S)
        ops.deepfilter_spec(act.contiguous(), cmp, enh, 1, F, self.df_order, S, 0, layout=1)  # CGN:128, 233
        if S > 1:
            y = _istft_fused(_merge_speakers(enh), self.n_fft, self.hop_length, self.win_length, L)
            return y.reshape(B, S, L), [all_out]  # the reference returns `*_` of fb_model(...): [all_layer_outputs]
        enh = enh[:, 0]
        return _istft_fused(enh, self.n_fft, self.hop_length, self.win_length, L), enh.abs()


# ------------------------------------------------------------------------------------------------
# surface B: the frozen competition model (recipes/intel_ndns/spiking_fullsubnet_freeze_phase/model_low_freq.py)
# ------------------------------------------------------------------------------------------------
EPSILON = 2.220446049250313e-16  # np.finfo(float).eps, audiozen/constant.py:11


class _FreezeSequenceModel(nn.Module):
    """`SequenceModel` of surface B (model_low_freq.py:42-139): no pre-LayerNorm, `fc_output_layer` instead of
    `proj`, activation selected by capitalised name.  Only sequence_model="GSU" is on the GSN path."""

    def __init__(self, input_size, output_size, hidden_size, num_layers, bidirectional, sequence_model="GSU",
                 output_activate_function="Tanh", num_groups=4, mogrify_steps=5, dropout=0.0,
                 shared_weights=False, bn=False):
        super().__init__()
        if sequence_model != "GSU":
            raise NotImplementedError(f"Not implemented {sequence_model}")
        self.sequence_model = efficient_spiking_neuron(input_size, hidden_size, num_layers,
                                                       shared_weights=shared_weights, bn=bn)
        if int(output_size):
            self.fc_output_layer = nn.Linear(hidden_size * (2 if bidirectional else 1), output_size)
        if output_activate_function:
            mods = {"Tanh": nn.Tanh, "ReLU": nn.ReLU, "ReLU6": nn.ReLU6, "LeakyReLU": nn.LeakyReLU,
                    "PReLU": nn.PReLU}
            if output_activate_function not in mods:
                raise NotImplementedError(f"Not implemented activation function {output_activate_function}")
            self.activate_function = mods[output_activate_function]()
        self.output_activate_function_name = output_activate_function
        self.output_size, self.input_size, self.hidden_size, self.num_layers = output_size, input_size, hidden_size, num_layers
        self.sequence_model_name = sequence_model

    # the interface the streaming plan reads (same names as SequenceModel)
    use_pre_layer_norm = False

    @property
    def proj(self):
        return self.fc_output_layer if int(self.output_size) else None

    @property
    def proj_size(self):
        return int(self.output_size)

    @property
    def _act(self):
        return {"Tanh": "tanh", "ReLU": "relu"}.get(self.output_activate_function_name or None)

    def run_time_major(self, x):
        """x [T,R,K] (already normalised) -> (fc_out [T,R,P], activated, all_layer_outputs)."""
        out, _, trace = self.sequence_model(x, None)
        if int(self.output_size):
            out = ops.linear(out, self.fc_output_layer.weight.detach(), self.fc_output_layer.bias.detach(),
                             spikes=True, bits=getattr(self.sequence_model, "last_bits", None))
            trace = trace + [out]
        act = self.activate_function(out) if self.output_activate_function_name else out
        return out, act, trace

    def forward(self, x):
        """x [B,F,T] -> ([B,P,T], all_layer_outputs) (model_low_freq.py:98-139)."""
        assert x.dim() == 3, f"Shape is {x.shape}."
        if not x.is_cuda:
            raise RuntimeError("spiking_fullsubnet_b200 has no CPU path: move the model and input to CUDA")
        _, act, trace = self.run_time_major(x.permute(2, 0, 1).contiguous())
        return act.permute(1, 2, 0).contiguous(), trace


class _FreezeSubBandWrapper(_FreezeSequenceModel):
    def __init__(self, df_order, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.df_order = df_order


def _utterance_norm(x, B, norm_type):
    """model_low_freq.py:146-217 on a time-major tensor x [T, B*N, K]: statistics per utterance over all of
    its (sub-band, feature, frame) entries -- the reference reduces dims 1.. of [B, N, 1, K, T] (:162, :475)."""
    T = x.shape[0]
    v = x.view(T, B, -1)
    if norm_type == "offline_laplace_norm":
        mu = v.mean(dim=(0, 2), keepdim=True)
        return (v / (mu + EPSILON)).view_as(x)
    if norm_type == "offline_gaussian_norm":
        mu = v.mean(dim=(0, 2), keepdim=True)
        std = v.permute(1, 0, 2).reshape(B, -1).std(dim=1).view(1, B, 1)
        return ((v - mu) / (std + EPSILON)).view_as(x)
    if norm_type == "cumulative_laplace_norm":
        # model_low_freq_count_time.py:173-204 (the variant of the recipe directory that accepts the 5-D sub-band
        # input; model_low_freq.py:181 unpacks four dims and fails on it): every row [B*N] is divided by the mean of
        # ITS OWN K features over the frames 0..t -- causal, so the sub-band models can follow the full-band one
        K = x.shape[2]
        cum = torch.cumsum(x.sum(dim=2), dim=0)                                        # [T, R]
        count = torch.arange(K, K * T + 1, K, dtype=x.dtype, device=x.device).view(T, 1)
        return x / ((cum / count).unsqueeze(-1) + EPSILON)
    raise NotImplementedError("You must set up a type of Norm. e.g. offline_laplace_norm, "
                              "cumulative_laplace_norm, forgetting_norm, etc.")


def _utterance_divisor(x, B, norm_type):
    """The divisor `_utterance_norm` applies, for the streaming front end (gsn_xplanes_stream row_div): [B] per
    utterance for offline_laplace_norm, [T, B*N] per row and frame for cumulative_laplace_norm; None when the norm is
    not a pure division (offline_gaussian_norm).  Same torch reductions as `_utterance_norm`: identical values."""
    T = x.shape[0]
    if norm_type == "offline_laplace_norm":
        return (x.view(T, B, -1).mean(dim=(0, 2)) + EPSILON).contiguous()
    if norm_type == "cumulative_laplace_norm":
        K = x.shape[2]
        cum = torch.cumsum(x.sum(dim=2), dim=0)
        count = torch.arange(K, K * T + 1, K, dtype=x.dtype, device=x.device).view(T, 1)
        return (cum / count + EPSILON).contiguous()
    return None


def _stream_divisor(cm, fb_act, N, lo, ctr, nbr, B, norm_type):
    """`_utterance_divisor` of the gathered input of one model WITHOUT materialising the gather: gsn_subband_rowsums
    leaves one sum per (frame, row), the rest is a reduction over T x R numbers.  (Summation order differs from
    `_utterance_divisor`'s: the last bit of the divisor may.)"""
    rs = ops.subband_rowsums(cm, fb_act, N, lo, ctr, nbr)                              # [T, B*N]
    T = rs.shape[0]
    K = ctr + 2 * nbr + (ctr if fb_act is not None else 0)
    if norm_type == "offline_laplace_norm":
        return (rs.sum(dim=0).view(B, N).sum(dim=1) / float(T * N * K) + EPSILON).contiguous()
    if norm_type == "cumulative_laplace_norm":
        cum = torch.cumsum(rs, dim=0)
        count = torch.arange(K, K * T + 1, K, dtype=rs.dtype, device=rs.device).view(T, 1)
        return (cum / count + EPSILON).contiguous()
    return None


class _FreezeSubbandModel(nn.Module):
    def __init__(self, freq_cutoffs, sb_num_center_freqs, sb_num_neighbor_freqs, fb_num_center_freqs,
                 fb_num_neighbor_freqs, sb_df_orders, sequence_model, hidden_size, activate_function=False,
                 norm_type="offline_laplace_norm", shared_weights=False, bn=False):
        super().__init__()
        if any(n != 0 for n in fb_num_neighbor_freqs) or list(fb_num_center_freqs) != list(sb_num_center_freqs):
            raise NotImplementedError("full-band neighbour bins / centre sizes different from the sub-band ones "
                                      "are not used by any shipped config and are not implemented")
        self.sb_models = nn.ModuleList([
            _FreezeSubBandWrapper(df_order=d, input_size=(c + n * 2) + (fc + fn * 2), output_size=c * 2 * d,
                                  hidden_size=hidden_size, num_layers=2, sequence_model=sequence_model,
                                  bidirectional=False, output_activate_function=activate_function,
                                  shared_weights=shared_weights, bn=bn)
            for c, n, fc, fn, d in zip(sb_num_center_freqs, sb_num_neighbor_freqs, fb_num_center_freqs,
                                       fb_num_neighbor_freqs, sb_df_orders)])
        self.freq_cutoffs = freq_cutoffs
        self.sb_num_center_freqs, self.sb_num_neighbor_freqs = sb_num_center_freqs, sb_num_neighbor_freqs
        self.fb_num_center_freqs, self.fb_num_neighbor_freqs = fb_num_center_freqs, fb_num_neighbor_freqs
        self.norm_type = norm_type

    def band_edges(self, num_freqs):
        """[(lo, hi)] per sub-band model from the interior cut-offs (model_low_freq.py:439-447)."""
        cuts = [0] + list(self.freq_cutoffs) + [num_freqs]
        return [(cuts[i], cuts[i + 1]) for i in range(len(self.sb_models))]

    concurrent_bands = True

    def run_time_major(self, cm, fb):
        T, B, F = cm.shape
        edges = self.band_edges(F)
        for i, (lo, hi) in enumerate(edges):
            ctr = self.sb_num_center_freqs[i]
            if (hi - lo) % ctr != 0:
                raise ValueError(f"The number of center frequencies should be divisible by the subband freqency "
                                 f"interval. Got num_center_freqs={ctr}, upper_cutoff_freq={hi}, and "
                                 f"lower_cutoff_freq={lo}.")

        def run_band(i):
            m, (lo, hi) = self.sb_models[i], edges[i]
            ctr, nbr = self.sb_num_center_freqs[i], self.sb_num_neighbor_freqs[i]
            x = ops.subband_features(cm, fb, (hi - lo) // ctr, lo, ctr, nbr)
            x = _utterance_norm(x, B, self.norm_type).contiguous()
            proj, act, trace = m.run_time_major(x)
            return act, trace

        n = len(self.sb_models)
        # the sub-band models are independent latency-bound recurrences: one stream each, the SMs split between them in
        # proportion to their rows (as SubbandModel.run_time_major does for surface A)
        demands = [_cluster_ctas(B * ((hi - lo) // self.sb_num_center_freqs[i]), m.hidden_size,
                                 m.sequence_model.layers[0].cell.shared_weights)
                   for i, (m, (lo, hi)) in enumerate(zip(self.sb_models, edges))]
        budgets = _sm_budgets(demands) if self.concurrent_bands and n > 1 else [0] * n
        for m, b in zip(self.sb_models, budgets):
            m.sequence_model.sm_budget = b
        if not self.concurrent_bands or n == 1:
            res = [run_band(i) for i in range(n)]
        else:
            main = torch.cuda.current_stream(cm.device)
            streams = _band_streams(cm.device, n, tag="surface_b")
            fork = torch.cuda.Event()
            fork.record(main)
            res = []
            for i in range(n):
                streams[i].wait_event(fork)
                with torch.cuda.stream(streams[i]):
                    res.append(run_band(i))
                    done = torch.cuda.Event()
                    done.record(streams[i])
                main.wait_event(done)
                if not torch.cuda.is_current_stream_capturing():
                    for t in [res[-1][0]] + list(res[-1][1]):
                        t.record_stream(main)
        return [r[0] for r in res], [r[1] for r in res]


class Separator(_StreamingPipeline, _GraphedNetwork, nn.Module):
    """Surface B `model_low_freq.Separator` (:485-618): same network as SpikingFullSubNet with utterance-level
    laplace normalisation instead of LayerNorm; the class the model-zoo checkpoints were trained with.
    forward(wave [B,L] or [B,1,L]) -> (enhanced_y, enhanced_mag, fb_all_layer_outputs, sb_all_layer_outputs)."""

    def __init__(self, sr, n_fft, hop_length, win_length, fdrc, num_freqs, fb_freqs, freq_cutoffs,
                 sb_num_center_freqs, sb_num_neighbor_freqs, fb_num_center_freqs, fb_num_neighbor_freqs,
                 fb_hidden_size, sb_hidden_size, sb_df_orders, sequence_model, fb_output_activate_function,
                 sb_output_activate_function, norm_type, shared_weights=False, bn=False):
        super().__init__()
        self.n_fft, self.hop_length, self.win_length, self.fdrc = n_fft, hop_length, win_length, fdrc
        self.freq_cutoffs, self.sb_df_orders = freq_cutoffs, sb_df_orders
        self.num_repeats, self.fb_freqs, self.num_freqs = num_freqs // fb_freqs, fb_freqs, num_freqs
        self.norm_type = norm_type
        _utterance_norm(torch.zeros(1, 1, 1), 1, norm_type)  # unknown / unsupported norm types fail here
        self.fb_model = _FreezeSequenceModel(input_size=fb_freqs, output_size=fb_freqs, hidden_size=fb_hidden_size,
                                             num_layers=2, bidirectional=False, sequence_model=sequence_model,
                                             output_activate_function=fb_output_activate_function,
                                             shared_weights=shared_weights, bn=bn)
        self.sb_model = _FreezeSubbandModel(freq_cutoffs=freq_cutoffs, sb_num_center_freqs=sb_num_center_freqs,
                                            sb_num_neighbor_freqs=sb_num_neighbor_freqs,
                                            fb_num_center_freqs=fb_num_center_freqs,
                                            fb_num_neighbor_freqs=fb_num_neighbor_freqs, sb_df_orders=sb_df_orders,
                                            hidden_size=sb_hidden_size, sequence_model=sequence_model,
                                            activate_function=sb_output_activate_function,
                                            shared_weights=shared_weights, bn=bn, norm_type=norm_type)

    def set_backend(self, backend):
        for m in self.modules():
            if isinstance(m, StackedGSU):
                m.backend = backend
        return self

    def _network(self, mag):
        """mag [B, n_fft//2+1, T] -> (sub-band fc outputs: list of [T, B*N_i, P_i], fb_all, sb_all)
        (model_low_freq.py:574-586)."""
        B, F, T = mag.shape
        cm = ops.compress_mag(mag if mag.is_complex() else mag.contiguous(), F - 1, self.fdrc)
        x = ops.subband_features(cm, None, 1, 0, self.fb_freqs, 0)
        x = _utterance_norm(x, B, self.norm_type).contiguous()
        _, fb_act, fb_all = self.fb_model.run_time_major(x)
        if self.num_repeats * fb_act.shape[2] < F - 1:
            raise ValueError("full-band output does not cover the spectrum")
        projs, sb_all = self.sb_model.run_time_major(cm, fb_act.contiguous())
        return projs, fb_all, sb_all

    def _network_sched(self, mag):
        if self.streaming:
            res = self._network_stream(mag)
            if res is not None:
                return res
        return self._network(mag)

    def _stream_models(self, B):
        """(full-band plan entry, sub-band plan entries): surface B runs them as TWO pipelines, because its norms
        divide the sub-band input by statistics of the full-band output over all frames."""
        sbm = self.sb_model
        fb = dict(m=self.fb_model, R=B, N=1, lo=0, ctr=self.fb_freqs, nbr=0, fb=False, needs_div=True)
        sbs = []
        for i, (m, (lo, hi)) in enumerate(zip(sbm.sb_models, sbm.band_edges(self.num_freqs))):
            ctr, nbr = sbm.sb_num_center_freqs[i], sbm.sb_num_neighbor_freqs[i]
            if (hi - lo) % ctr != 0:
                raise ValueError(f"The number of center frequencies should be divisible by the subband freqency "
                                 f"interval. Got num_center_freqs={ctr}, upper_cutoff_freq={hi}, and "
                                 f"lower_cutoff_freq={lo}.")
            N = (hi - lo) // ctr
            sbs.append(dict(m=m, R=B * N, N=N, lo=lo, ctr=ctr, nbr=nbr, fb=True, needs_div=True))
        return fb, sbs

    def _stream_plan(self, B, sm_total=None):
        if self.norm_type not in ("offline_laplace_norm", "cumulative_laplace_norm"):
            return None
        fb, sbs = self._stream_models(B)
        for d in [fb] + sbs:
            if d["m"].proj is None or (d["m"].output_activate_function_name and d["m"]._act is None):
                return None
        p1, p2 = self._stream_plan_for([fb], sm_total), self._stream_plan_for(sbs, sm_total)
        return None if p1 is None or p2 is None else (p1, p2)

    def _network_stream(self, mag):
        """The streaming schedule for surface B: pipeline 1 = the full-band model, pipeline 2 = all sub-band models (their
        input is divided by utterance-level statistics of the COMPLETE full-band output, model_low_freq.py:146-171,
        475).  The statistics come from the same gather + torch reductions as the eager path (identical divisors); the
        division itself happens in gsn_xplanes_stream."""
        B, F, T = mag.shape
        plan = self._stream_plan(B)
        if plan is None or F - 1 != self.num_freqs:
            return None
        p1, p2 = plan
        dev = mag.device
        ops.stream_preload(dev)
        cm = ops.compress_mag(mag if mag.is_complex() else mag.contiguous(), F - 1, self.fdrc)
        p1[0]["div"] = _stream_divisor(cm, None, 1, 0, self.fb_freqs, 0, B, self.norm_type)
        r1 = self._stream_run(p1, cm, tag="b1")
        fb_act = r1[0][3]
        if self.num_repeats * fb_act.shape[2] < F - 1:
            raise ValueError("full-band output does not cover the spectrum")
        # the bands' divisors on forked streams (row sums of the gathered input + a small reduction each): the sub-band
        # pipeline cannot start before the last of them
        main = torch.cuda.current_stream(dev)
        streams = _band_streams(dev, len(p2), tag="surface_b_div")
        fork = torch.cuda.Event()
        fork.record(main)
        for d, st in zip(p2, streams):
            st.wait_event(fork)
            with torch.cuda.stream(st):
                d["div"] = _stream_divisor(cm, fb_act, d["N"], d["lo"], d["ctr"], d["nbr"], B, self.norm_type)
                done = torch.cuda.Event()
                done.record(st)
            main.wait_event(done)
            if not torch.cuda.is_current_stream_capturing():
                d["div"].record_stream(main)
        r2 = self._stream_run(p2, cm, fb_act=fb_act, tag="b2")
        return [r[3] for r in r2], r1[0][1], [r[1] for r in r2]

    def coefficients(self, mag):
        """[B, df, F_i, T, 2] per band: '(b n) (c fc df) t -> b df (n fc) t c' (model_low_freq.py:257-263)."""
        projs, fb_all, sb_all = self.network(mag)
        B = mag.shape[0]
        return [coef_layout(p, B, p.shape[1] // B, d, 1)[:, :, 0] for p, d in zip(projs, self.sb_df_orders)], \
            fb_all, sb_all

    def forward(self, noisy_y):
        ndim = noisy_y.dim()
        assert ndim in (2, 3), "Input must be 2D (B, T) or 3D tensor (B, 1, T)"
        if ndim == 3:
            assert noisy_y.size(1) == 1, "Input must be 2D (B, T) or 3D tensor (B, 1, T)"
            noisy_y = noisy_y.squeeze(1)
        if not noisy_y.is_cuda:
            raise RuntimeError("spiking_fullsubnet_b200 has no CPU path: move the model and input to CUDA")
        if _needs_autograd(self):
            self._warn_cell_hooks()
            from . import training
            return training.separator_forward(self, noisy_y)
        B, L = noisy_y.shape
        F = self.n_fft // 2 + 1
        cmp = _stft_fused(noisy_y, self.n_fft, self.hop_length, self.win_length, f_keep=F - 1, fdrc=self.fdrc)
        projs, fb_all, sb_all = self.network(cmp)
        F, T = cmp.shape[1], cmp.shape[2]
        ctrs = list(self.sb_model.sb_num_center_freqs)[:len(projs)]
        Ns = [(b - a) // c for (a, b), c in zip(self.sb_model.band_edges(F - 1), ctrs)]
        res = _fused_back_end(projs, cmp, Ns, ctrs, list(self.sb_df_orders)[:len(projs)], 0, self.n_fft, self.hop_length,
                              self.win_length, L)
        if res is not None:
            return res[0], res[1], fb_all, sb_all
        enh = _empty_spec_like(cmp, 1)
        lo = 0
        for i, (p, (a, b)) in enumerate(zip(projs, self.sb_model.band_edges(F - 1))):
            ctr = self.sb_model.sb_num_center_freqs[i]
            n = (b - a) // ctr
            ops.deepfilter_spec(p.contiguous(), cmp, enh, n, ctr, self.sb_df_orders[i], 1, lo)
            lo += n * ctr
        ops.spec_passthrough(cmp, enh, lo)
        enh = enh[:, 0]
        y = _istft_fused(enh, self.n_fft, self.hop_length, self.win_length, L)  # asynchronous, graph-capturable
        return y, enh.abs(), fb_all, sb_all
