"""Autograd (training) path of the drop-in modules.

The recurrence -- forward with training-mode BatchNorm statistics and BPTT with the Triangle surrogate
gradient (ESN:84-101, 132-153) -- runs in libgsn_b200 (`gsn_layer_train_forward/backward`, cooperative CUDA
kernels) behind `GSNLayerFn`; everything around it (STFT, compression, sub-band gather, LayerNorm, the
input/output projections, deep filter, iSTFT) is expressed with differentiable torch ops so that PyTorch's
autograd, `accelerator.backward`, DDP gradient all-reduce and `clip_grad_norm_` of the reference trainer
(`audiozen/trainer.py:409-422`) work unchanged.  The gradients of the weights inside the recurrence are
formed from the kernel's dL/d(gate pre-activation) trace with plain GEMMs:
    dW_hh = dz^T h_{t-1},   dW_ih = dz^T x,   dx = dz W_ih   (the last two by autograd through F.linear).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import ops


class GSNLayerFn(torch.autograd.Function):
    """h[T,R,H] = GSULayer(xproj[T,R,gH]) with zero initial state; xproj = x @ W_ih^T (no bias)."""

    @staticmethod
    def forward(ctx, xproj, w_hh, bias, bn_weight, bn_bias, cell):
        bn = cell.batchnorm if cell.use_bn else None
        training = bool(bn is not None and bn.training)
        if training and bn.momentum is None:
            raise NotImplementedError("BatchNorm1d(momentum=None) (cumulative moving average) is not supported by the "
                                      "training kernels; the reference recipes use the default momentum")
        momentum = 0.1 if bn is None else bn.momentum
        if bn is not None and not training and (ctx.needs_input_grad[3] or ctx.needs_input_grad[4]):
            # eval-mode BatchNorm with gradients enabled: the reference's autograd would give bn.weight / bn.bias a
            # gradient through c; the kernels only form them from batch statistics.  Say so instead of returning
            # silent Nones (freeze the affine with requires_grad_(False) to fine-tune with frozen statistics).
            import warnings
            warnings.warn("spiking_fullsubnet_b200: BatchNorm in eval mode gets no weight / bias gradient from the GSN "
                          "training kernels (its parameters are treated as frozen)", stacklevel=2)
        eps = 1e-5 if bn is None else bn.eps
        rm = bn.running_mean if bn is not None else None
        rv = bn.running_var if bn is not None else None
        T = xproj.shape[0]
        h, c, f, g, xhat, invstd = ops.layer_train_forward(
            xproj.contiguous(), w_hh.contiguous(), bias.contiguous(),
            None if bn is None else bn_weight.contiguous(), None if bn is None else bn_bias.contiguous(),
            rm, rv, training, momentum, eps, cell.shared_weights)
        if training:
            bn.num_batches_tracked += T  # BatchNorm is called once per frame (ESN:149-150)
        rv_saved = rv.clone() if (bn is not None and not training) else None
        ctx.save_for_backward(w_hh, h, c, f, g, xhat, invstd, bn_weight if bn is not None else None, rv_saved)
        ctx.meta = (training, eps, cell.shared_weights, bn is not None)
        return h

    @staticmethod
    def backward(ctx, dh):
        w_hh, h, c, f, g, xhat, invstd, bn_weight, rv = ctx.saved_tensors
        training, eps, shared, has_bn = ctx.meta
        dz, dbias, dgamma, dbeta = ops.layer_train_backward(
            dh.contiguous(), w_hh.contiguous(), c, f, g, xhat, invstd,
            bn_weight.contiguous() if has_bn else None, rv, training, eps, shared)
        T, R, gH = dz.shape
        H = h.shape[2]
        # dW_hh = sum_t dz_t^T h_{t-1}, h_{-1} = 0: frames 1..T-1 of dz against frames 0..T-2 of h
        if T > 1:
            dw_hh = dz[1:].reshape(-1, gH).t() @ h[:-1].reshape(-1, H)
        else:
            dw_hh = torch.zeros_like(w_hh)
        if has_bn and not training:  # eval-mode BatchNorm: the affine still gets gradients through c
            dgamma = dbeta = None    # (not needed by any recipe: BN in eval mode is frozen)
        return dz, dw_hh, dbias, dgamma, dbeta, None


def unfold_index(lo, hi, ctr, nbr, num_freqs, device):
    """Bin indices [N, ctr + 2*nbr] of `_freq_unfold` (MSF:265-312): reflect padding only at the spectrum edges."""
    if (hi - lo) % ctr != 0:
        raise ValueError(f"Number of frequency bins must be divisible by the center frequency."
                         f"GOT: ctr_freq={ctr}, upper_cutoff_freq={hi}, lower_cutoff_freq={lo}")
    n = torch.arange((hi - lo) // ctr, device=device)[:, None]
    j = torch.arange(ctr + 2 * nbr, device=device)[None, :]
    q = lo + n * ctr - nbr + j
    q = torch.where(q < 0, -q, q)
    q = torch.where(q > num_freqs - 1, 2 * (num_freqs - 1) - q, q)
    return q


def run_cell_layer(cell, x):
    """One GSULayer (ESN:75-81) with autograd from a zero state: x [T,R,K] -> h [T,R,H]."""
    bn = cell.batchnorm if cell.use_bn else None
    return GSNLayerFn.apply(F.linear(x, cell.weight_ih), cell.weight_hh, cell.bias_ih,
                            bn.weight if bn is not None else None, bn.bias if bn is not None else None, cell)


def run_stack(stack, x):
    """StackedGSU.forward (ESN:50-62) with autograd: x [T,R,K] -> (out, all_layer_output)."""
    trace = [x]
    out = x
    for layer in stack.layers:
        cell = layer.cell
        xproj = F.linear(out, cell.weight_ih)
        bn = cell.batchnorm if cell.use_bn else None
        out = GSNLayerFn.apply(xproj, cell.weight_hh, cell.bias_ih, bn.weight if bn is not None else None,
                               bn.bias if bn is not None else None, cell)
        trace.append(out)
    return out, trace


def run_sequence_model(m, x):
    """x [T,R,K] already normalised -> (activated output [T,R,P], all_layer_outputs) (MSF:115-122; surface B:
    model_low_freq.py:98-139 with `fc_output_layer` / `activate_function`)."""
    out, trace = run_stack(m.sequence_model, x)
    if hasattr(m, "fc_output_layer") or hasattr(m, "output_size"):  # surface B naming
        if int(m.output_size):
            out = m.fc_output_layer(out)
            trace = trace + [out]
        return (m.activate_function(out) if m.output_activate_function_name else out), trace
    out = m.proj(out)
    trace = trace + [out]
    return m.output_activate_function(out), trace


def deepfilter(spec, coef, order):
    """spec complex [B,F,T], coef [B,df,S,F,T,2] -> [B,S,F,T]  (MSF:315-346)."""
    T = spec.shape[-1]
    cc = torch.complex(coef[..., 0], coef[..., 1])
    pad = F.pad(spec, (order - 1, 0))
    out = 0
    for d in range(order):
        out = out + pad[:, None, :, d:d + T] * cc[:, d]
    return out


def _fork_join(device, thunks):
    """Run independent pieces of the forward on forked streams and join them on the caller's stream."""
    from .modeling import _band_streams
    if len(thunks) == 1:
        return [thunks[0]()]
    main = torch.cuda.current_stream(device)
    streams = _band_streams(device, len(thunks))
    fork = torch.cuda.Event()
    fork.record(main)
    results = []
    for st, fn in zip(streams, thunks):
        st.wait_event(fork)
        with torch.cuda.stream(st):
            res = fn()
            done = torch.cuda.Event()
            done.record(st)
        main.wait_event(done)
        for t in _tensors(res):
            t.record_stream(main)
        results.append(res)
    return results


def _tensors(obj):
    if torch.is_tensor(obj):
        yield obj
    elif isinstance(obj, (list, tuple)):
        for o in obj:
            yield from _tensors(o)


def run_subband_model(sbm, cm, fb_act):
    """SubbandModel (MSF:216-263) with autograd on time-major tensors: cm [T,B,F] compressed magnitude, fb_act
    [T,B,f_fb] full-band output (tiled by index, MSF:443) -> (activated outputs [T, B*N_i, P_i], all_layer_outputs)."""
    T, B, Fq = cm.shape

    def band(i, m):
        lo, hi = sbm.freq_cutoffs[i], sbm.freq_cutoffs[i + 1]
        ctr, nbr = sbm.center_freq_sizes[i], sbm.neighbor_freq_sizes[i]
        qi = unfold_index(lo, hi, ctr, nbr, Fq, cm.device)          # [N, ctr+2nbr]
        qf = unfold_index(lo, hi, ctr, 0, Fq, cm.device) % fb_act.shape[2]  # tiled full-band output (MSF:443)
        N = qi.shape[0]
        xb = torch.cat([cm[:, :, qi], fb_act[:, :, qf]], dim=-1).reshape(T, B * N, -1)
        if m.use_pre_layer_norm:
            xb = m.pre_layer_norm(xb)
        return run_sequence_model(m, xb.contiguous())

    # the sub-band models are independent: one stream each, forward AND backward (autograd replays every
    # backward node on the stream of its forward), so their latency-bound recurrence kernels overlap
    res = _fork_join(cm.device, [lambda i=i, m=m: band(i, m) for i, m in enumerate(sbm.sb_models)])
    return [r[0] for r in res], [r[1] for r in res]


def spiking_fullsubnet_forward(model, wave):
    """SpikingFullSubNet.forward (MSF:415-474) on the autograd path."""
    from .modeling import _istft, _stft, coef_layout
    B, L = wave.shape
    cmp = _stft(wave, model.n_fft, model.hop_length, model.win_length)  # [B,F,T]
    mag = cmp.abs()
    Fq = mag.shape[1] - 1
    cm = (mag ** model.fdrc)[:, :-1, :].permute(2, 0, 1)  # [T,B,256] time-major
    fbm = model.fb_model
    x = cm[..., : model.fb_input_size]
    if fbm.use_pre_layer_norm:
        x = fbm.pre_layer_norm(x)
    fb_act, fb_all = run_sequence_model(fbm, x.contiguous())
    S = model.num_spks
    acts, sb_all = run_subband_model(model.sb_model, cm, fb_act)
    coefs = [coef_layout(a, B, a.shape[1] // B, d, S) for a, d in zip(acts, model.sb_model.df_orders)]
    enh, lo = [], 0
    for coef, order in zip(coefs, model.df_orders):
        nf = coef.shape[3]
        enh.append(deepfilter(cmp[:, lo:lo + nf], coef, order))
        lo += nf
    enh = torch.cat(enh + [cmp[:, None, lo:].expand(-1, S, -1, -1)], dim=2)  # un-filtered bins pass through
    Fz, Tz = enh.shape[2], enh.shape[3]
    if S > 1:
        y = _istft(enh.reshape(B * S, Fz, Tz), model.n_fft, model.hop_length, model.win_length, L)
        return y.reshape(B, S, L), fb_all, sb_all
    enh = enh[:, 0]
    return _istft(enh, model.n_fft, model.hop_length, model.win_length, L), enh.abs(), fb_all, sb_all


def _utterance_norm(x, B, norm_type):
    """Differentiable twin of modeling._utterance_norm (model_low_freq.py:146-217) on x [T, B*N, K]."""
    from .modeling import EPSILON
    T = x.shape[0]
    v = x.reshape(T, B, -1)
    mu = v.mean(dim=(0, 2), keepdim=True)
    if norm_type == "offline_laplace_norm":
        return (v / (mu + EPSILON)).reshape(x.shape)
    std = v.permute(1, 0, 2).reshape(B, -1).std(dim=1).view(1, B, 1)
    return ((v - mu) / (std + EPSILON)).reshape(x.shape)


def separator_forward(model, wave):
    """Surface B `Separator.forward` (model_low_freq.py:561-618) on the autograd path."""
    from .modeling import _stft, coef_layout
    B, L = wave.shape
    cmp = _stft(wave, model.n_fft, model.hop_length, model.win_length)
    mag = cmp.abs()
    Fq = mag.shape[1] - 1
    cm = (mag ** model.fdrc)[:, :-1, :].permute(2, 0, 1)  # [T,B,256]
    T = cm.shape[0]
    x = _utterance_norm(cm[..., : model.fb_freqs], B, model.norm_type)
    fb_act, fb_all = run_sequence_model(model.fb_model, x.contiguous())
    sbm = model.sb_model
    coefs, sb_all = [], []

    def band(i, m, lo, hi):
        ctr, nbr = sbm.sb_num_center_freqs[i], sbm.sb_num_neighbor_freqs[i]
        qi = unfold_index(lo, hi, ctr, nbr, Fq, cm.device)
        qf = unfold_index(lo, hi, ctr, 0, Fq, cm.device) % fb_act.shape[2]
        N = qi.shape[0]
        xb = torch.cat([cm[:, :, qi], fb_act[:, :, qf]], dim=-1).reshape(T, B * N, -1)
        xb = _utterance_norm(xb, B, model.norm_type)
        act, trace = run_sequence_model(m, xb.contiguous())
        return coef_layout(act, B, N, model.sb_df_orders[i], 1), trace

    for c, tr in _fork_join(cm.device, [lambda i=i, m=m, lo=lo, hi=hi: band(i, m, lo, hi)
                                        for i, (m, (lo, hi)) in enumerate(zip(sbm.sb_models, sbm.band_edges(Fq)))]):
        coefs.append(c)
        sb_all.append(tr)
    enh, lo = [], 0
    for coef, order in zip(coefs, model.sb_df_orders):
        nf = coef.shape[3]
        enh.append(deepfilter(cmp[:, lo:lo + nf], coef, order))
        lo += nf
    enh = torch.cat(enh + [cmp[:, None, lo:]], dim=2)[:, 0]
    y = torch.istft(enh, model.n_fft, model.hop_length, model.win_length,
                    window=torch.hann_window(model.win_length, device=wave.device), length=L)
    return y, enh.abs(), fb_all, sb_all


def cirm_gsn_forward(model, wave):
    """cirm_gsn `Model.forward` (CGN:206-244) on the autograd path."""
    from .modeling import _istft, _stft
    B, L = wave.shape
    cmp = _stft(wave, model.n_fft, model.hop_length, model.win_length)
    cm = (cmp.abs() ** model.fdrc).permute(2, 0, 1)  # all bins (CGN:226)
    fbm = model.fb_model
    x = fbm.pre_layer_norm(cm) if fbm.use_pre_layer_norm else cm
    act, all_out = run_sequence_model(fbm, x.contiguous())
    T, _, P = act.shape
    d, S = model.df_order, model.num_spks
    coef = act.reshape(T, B, 2, d, S, P // (2 * d * S)).permute(1, 3, 4, 5, 0, 2)  # b d s f t c
    enh = deepfilter(cmp, coef, d)  # [B,S,F,T]
    if S > 1:
        y = _istft(enh.reshape(B * S, *enh.shape[2:]), model.n_fft, model.hop_length, model.win_length, L)
        return y.reshape(B, S, L), [all_out]
    enh = enh[:, 0]
    return _istft(enh, model.n_fft, model.hop_length, model.win_length, L), enh.abs()
