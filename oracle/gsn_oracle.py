"""CPU oracle for the GSN hot path of Spiking-FullSubNet -- TEST INFRASTRUCTURE ONLY.

This module is a numpy restatement of the reference's algorithm for the path named in
BASELINE.json `north_star`.  It is the *checker*: only `tests/`,
`__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of `bench.py` may import
it.  Nothing under `spiking_fullsubnet_b200/` (the product) imports it, and the product has no CPU
fallback.

Parity pin: the reference's own tests hold no golden vectors for this path (SURVEY.md section 4,
8c).  The oracle is therefore pinned against outputs of the reference itself, generated in the
build container by `tests/golden/make_golden.py` (imports `/root/reference` read-only) and committed
under `tests/golden/*.npz`; `tests/test_oracle_golden.py` checks every fixture.

All citations are relative to the reference root (`/root/reference/`):
  ESN = audiozen/models/spiking_fullsubnet/efficient_spiking_neuron.py
  MSF = audiozen/models/spiking_fullsubnet/modeling_spiking_fullsubnet.py
  CGN = audiozen/models/cirm_gsn/modeling_cirm_gsn.py

Parameters are passed as a flat dict keyed by the reference's state_dict names
(e.g. "fb_model.sequence_model.layers.0.cell.weight_ih"), values numpy arrays.
Every function takes `dtype` (np.float32 reproduces the reference's arithmetic type; np.float64 gives
the high-precision arbiter used to decide which side of a disagreement is closer to exact).
"""
from __future__ import annotations

import numpy as np

BN_EPS = 1e-5  # torch.nn.BatchNorm1d default, ESN:122-123
LN_EPS = 1e-5  # torch.nn.LayerNorm default, MSF:27


# --------------------------------------------------------------------------------------------
# a1/a2: one frame of one layer (ESN:132-153) and the Heaviside spike (ESN:84-92)
# --------------------------------------------------------------------------------------------
def sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def gsu_cell_step(x_t, h_prev, c_prev, w_ih, w_hh, bias, bn, shared, dtype=np.float32):
    """One GSUCell.forward (ESN:132-153), eval-mode BatchNorm.

    x_t [R,K], h_prev/c_prev [R,H]; w_ih [gH,K], w_hh [gH,H] with g = 1 (shared) or 2; bias [2H];
    bn = None or dict(weight, bias, running_mean, running_var) each [H].
    Returns (h_t, c_t).  Threshold 0, no reset (ESN:151, :89).
    """
    H = h_prev.shape[1]
    z = x_t @ w_ih.T + h_prev @ w_hh.T  # ESN:140-145 (bias added below; same for both halves)
    if shared:  # ESN:134-136 -- repeat((2,1)) means both gate halves see the same product
        f_hat = z + bias[:H]
        g_hat = z + bias[H:]
    else:
        f_hat = z[:, :H] + bias[:H]
        g_hat = z[:, H:] + bias[H:]
    f = sigmoid(f_hat.astype(dtype))
    c = f * c_prev + (1.0 - f) * g_hat  # ESN:148
    if bn is not None:  # ESN:149-150, eval mode: running statistics
        inv = 1.0 / np.sqrt(bn["running_var"].astype(dtype) + dtype(BN_EPS))
        c = (c - bn["running_mean"]) * inv * bn["weight"] + bn["bias"]
    c = c.astype(dtype)
    h = (c >= 0).astype(dtype)  # ESN:89
    return h, c


def gsu_cell_step_train_bn(x_t, h_prev, c_prev, w_ih, w_hh, bias, bn, shared, dtype=np.float32):
    """Same as gsu_cell_step but BatchNorm in training mode (batch statistics over rows, biased
    variance for the normalisation; ESN:149-150 with nn.BatchNorm1d.training=True).  Returns
    (h_t, c_t, batch_mean, batch_var_biased)."""
    H = h_prev.shape[1]
    z = x_t @ w_ih.T + h_prev @ w_hh.T
    if shared:
        f_hat, g_hat = z + bias[:H], z + bias[H:]
    else:
        f_hat, g_hat = z[:, :H] + bias[:H], z[:, H:] + bias[H:]
    f = sigmoid(f_hat.astype(dtype))
    c = f * c_prev + (1.0 - f) * g_hat
    mu = c.mean(axis=0)
    var = c.var(axis=0)
    c = (c - mu) / np.sqrt(var + dtype(BN_EPS)) * bn["weight"] + bn["bias"]
    c = c.astype(dtype)
    return (c >= 0).astype(dtype), c, mu, var


# --------------------------------------------------------------------------------------------
# a3/a4: time loop + layer loop (ESN:50-62, 75-81)
# --------------------------------------------------------------------------------------------
def _layer_params(params, prefix, l, dtype):
    p = f"{prefix}layers.{l}.cell."
    bn = None
    if p + "batchnorm.weight" in params:
        bn = {k: np.asarray(params[p + "batchnorm." + k], dtype=dtype)
              for k in ("weight", "bias", "running_mean", "running_var")}
    return (np.asarray(params[p + "weight_ih"], dtype=dtype),
            np.asarray(params[p + "weight_hh"], dtype=dtype),
            np.asarray(params[p + "bias_ih"], dtype=dtype), bn)


def gsn_stack_forward(x, params, prefix, num_layers, shared, dtype=np.float32, return_c=False,
                      teacher=None):
    """StackedGSU.forward (ESN:50-62): layer-outer, time-inner; zero initial state (MSF:100-106).

    x [T,R,K] -> (out [T,R,H], all_layer_output = [x, h1, ..., hL], c_traces (list of [T,R,H]) if
    return_c).  `teacher`, if given, is a list (per layer) of (h_prev_trace, c_prev_trace) arrays
    [T,R,H] that REPLACE the carried state at each step (protocol P1).
    """
    x = np.asarray(x, dtype=dtype)
    T, R, _ = x.shape
    all_out = [x]
    c_traces = []
    cur = x
    for l in range(num_layers):
        w_ih, w_hh, bias, bn = _layer_params(params, prefix, l, dtype)
        H = w_hh.shape[1]
        h = np.zeros((R, H), dtype=dtype)
        c = np.zeros((R, H), dtype=dtype)
        hs = np.empty((T, R, H), dtype=dtype)
        cs = np.empty((T, R, H), dtype=dtype)
        for t in range(T):  # ESN:78-80
            if teacher is not None:
                h, c = teacher[l][0][t].astype(dtype), teacher[l][1][t].astype(dtype)
            h, c = gsu_cell_step(cur[t], h, c, w_ih, w_hh, bias, bn, shared, dtype)
            hs[t], cs[t] = h, c
        all_out.append(hs)
        c_traces.append(cs)
        cur = hs
    if return_c:
        return cur, all_out, c_traces
    return cur, all_out


# --------------------------------------------------------------------------------------------
# a6: SequenceModel.forward (MSF:81-125)
# --------------------------------------------------------------------------------------------
def layer_norm(x, weight, bias, eps=LN_EPS):
    """nn.LayerNorm over the last dim (biased variance), MSF:27,111-112."""
    mu = x.mean(axis=-1, keepdims=True)
    var = ((x - mu) ** 2).mean(axis=-1, keepdims=True)
    return (x - mu) / np.sqrt(var + x.dtype.type(eps)) * weight + bias


_ACT = {
    "tanh": np.tanh,
    "sigmoid": sigmoid,
    "relu": lambda v: np.maximum(v, 0),
}


def sequence_model_forward(inp, params, prefix, num_layers, shared, activation=None,
                           dtype=np.float32, return_c=False, proj="proj"):
    """inp [R,K,T] -> (out [R,P,T], all_layer_outputs [x_norm, h1..hL, proj_out] each [T,R,.]).

    Follows MSF:81-125: 'b f t -> t b f', optional pre_layer_norm, stack, proj (Linear or Identity),
    append proj output to the trace list, activation, back to 'b f t'.
    """
    x = np.ascontiguousarray(np.transpose(np.asarray(inp, dtype=dtype), (2, 0, 1)))  # MSF:108
    if prefix + "pre_layer_norm.weight" in params:  # MSF:111-112
        x = layer_norm(x, np.asarray(params[prefix + "pre_layer_norm.weight"], dtype=dtype),
                       np.asarray(params[prefix + "pre_layer_norm.bias"], dtype=dtype)).astype(dtype)
    res = gsn_stack_forward(x, params, prefix + "sequence_model.", num_layers, shared, dtype,
                            return_c=return_c)
    out, all_out = res[0], res[1]
    if prefix + proj + ".weight" in params:  # MSF:118 (surface B: fc_output_layer, model_low_freq.py:126)
        out = out @ np.asarray(params[prefix + proj + ".weight"], dtype=dtype).T \
            + np.asarray(params[prefix + proj + ".bias"], dtype=dtype)
    all_out = all_out + [out]  # MSF:119
    if activation in _ACT:  # MSF:54-61,122
        out = _ACT[activation](out)
    out = np.transpose(out, (1, 2, 0))  # MSF:124
    if return_c:
        return out, all_out, res[2]
    return out, all_out


# --------------------------------------------------------------------------------------------
# a8: sub-band unfold as an index map (MSF:265-312)
# --------------------------------------------------------------------------------------------
def freq_unfold_index(lo, hi, ctr, nbr, num_freqs):
    """Bin indices [N, ctr+2*nbr] that `_freq_unfold` gathers for the band [lo,hi).

    Reflect padding (edge bin not repeated) only at the spectrum edges (MSF:290-299); interior bands
    read real neighbours.  Raises ValueError like MSF:283-287.
    """
    if (hi - lo) % ctr != 0:
        raise ValueError(f"Number of frequency bins must be divisible by the center frequency. "
                         f"GOT: ctr_freq={ctr}, upper_cutoff_freq={hi}, lower_cutoff_freq={lo}")
    n_sub = (hi - lo) // ctr
    j = np.arange(ctr + 2 * nbr)[None, :]
    n = np.arange(n_sub)[:, None]
    q = lo + n * ctr - nbr + j
    if lo == 0:
        q = np.where(q < 0, -q, q)
    elif hi == num_freqs:
        q = np.where(q > num_freqs - 1, 2 * (num_freqs - 1) - q, q)
    if q.min() < 0 or q.max() > num_freqs - 1:
        raise ValueError("sub-band neighbourhood leaves the spectrum")
    return q


def subband_inputs(noisy, fb_out, lo, hi, ctr, nbr):
    """[B,F,T] noisy magnitude + [B,F,T] tiled full-band output -> [B*N, (ctr+2nbr)+ctr, T]
    (MSF:241-258 concat, then 'b n c fs t -> (b n) (c fs) t' MSF:155)."""
    B, F, T = noisy.shape
    qi = freq_unfold_index(lo, hi, ctr, nbr, F)  # [N, ctr+2nbr]
    qf = freq_unfold_index(lo, hi, ctr, 0, F)  # [N, ctr]
    a = noisy[:, qi, :]  # [B,N,fs,T]
    b = fb_out[:, qf, :]
    x = np.concatenate([a, b], axis=2)
    return x.reshape(B * qi.shape[0], x.shape[2], T)


def subband_coef_layout(out, B, ctr, df, num_spks=1):
    """'(b n) (c fc df s) t -> b df s (n fc) t c' (MSF:160-167). out [B*N, P, T]."""
    BN, P, T = out.shape
    N = BN // B
    o = out.reshape(B, N, 2, ctr, df, num_spks, T)
    o = np.transpose(o, (0, 4, 5, 1, 3, 6, 2))  # b df s n fc t c
    return np.ascontiguousarray(o.reshape(B, df, num_spks, N * ctr, T, 2))


# --------------------------------------------------------------------------------------------
# a9: network part of SpikingFullSubNet.forward (MSF:434-447): magnitude in -> coefficients out
# --------------------------------------------------------------------------------------------
def compress_mag(mag, fdrc, dtype=np.float32):
    mag = np.asarray(mag, dtype=dtype)
    if fdrc == 0.5:  # torch.pow(x, 0.5) is evaluated as sqrt
        return np.sqrt(mag)
    if fdrc == 1.0:
        return mag
    return np.power(mag, dtype(fdrc))


def spiking_fullsubnet_network(mag, params, cfg, dtype=np.float32):
    """mag [B, n_fft//2+1, T] (STFT magnitude) -> (df_coef_list, fb_all_layer_outputs,
    sb_all_layer_outputs) exactly as MSF:434-447 produce them.  cfg = the TOML [model.args] dict."""
    num_spks = cfg.get("num_spks", 1)
    shared = cfg.get("shared_weights", False)
    cm = compress_mag(mag, cfg["fdrc"], dtype)[:, :-1, :]  # MSF:435-436
    fb_in = cm[:, : cfg["fb_input_size"], :]  # MSF:439
    act = cfg.get("fb_output_activate_function")
    fb_out, fb_all = sequence_model_forward(fb_in, params, "fb_model.", cfg["fb_num_layers"], shared,
                                            act if isinstance(act, str) else None, dtype)
    rep = (cfg["n_fft"] // 2 + 1) // cfg["fb_input_size"]  # MSF:443
    fb_tiled = np.tile(fb_out, (1, rep, 1))
    coefs, sb_all = [], []
    cuts = cfg["freq_cutoffs"]
    for i, (ctr, nbr, df) in enumerate(zip(cfg["center_freq_sizes"], cfg["neighbor_freq_sizes"],
                                           cfg["df_orders"])):
        x = subband_inputs(cm, fb_tiled, cuts[i], cuts[i + 1], ctr, nbr)
        out, all_out = sequence_model_forward(x, params, f"sb_model.sb_models.{i}.",
                                              cfg["sb_num_layers"], shared, None, dtype)
        coefs.append(subband_coef_layout(out, mag.shape[0], ctr, df, num_spks))
        sb_all.append(all_out)
    return coefs, fb_all, sb_all


# --------------------------------------------------------------------------------------------
# a10: surface B, recipes/intel_ndns/spiking_fullsubnet_freeze_phase/model_low_freq.py (MLF)
# --------------------------------------------------------------------------------------------
EPSILON = np.finfo(float).eps  # audiozen/constant.py:11


def offline_laplace_norm(x):
    """MLF:146-172: divide by the utterance-level mean over every non-batch dim."""
    mu = x.mean(axis=tuple(range(1, x.ndim)), keepdims=True, dtype=x.dtype)
    return (x / (mu + x.dtype.type(EPSILON))).astype(x.dtype)


def cumulative_laplace_norm(x):
    """recipes/intel_ndns/spiking_fullsubnet_freeze_phase/model_low_freq_count_time.py:173-204 on [..., K, T]: every
    leading index is divided by the mean of its own K x (t+1) entries up to frame t."""
    K, T = x.shape[-2], x.shape[-1]
    cum = np.cumsum(x.sum(axis=-2, dtype=x.dtype), axis=-1, dtype=x.dtype)            # [..., T]
    count = np.arange(K, K * T + 1, K, dtype=x.dtype)
    return (x / ((cum / count)[..., None, :] + x.dtype.type(EPSILON))).astype(x.dtype)


def separator_network(mag, params, cfg, dtype=np.float32):
    """MLF:574-586: mag [B, n_fft//2+1, T] -> (coef list [B,df,F_i,T,2], fb_all, sb_all); offline laplace norm, or the
    cumulative one of the count_time variant."""
    assert cfg["norm_type"] in ("offline_laplace_norm", "cumulative_laplace_norm")
    shared = cfg.get("shared_weights", False)
    B = mag.shape[0]
    cm = compress_mag(mag, cfg["fdrc"], dtype)[:, :-1, :]
    nf = cm.shape[1]
    cumulative = cfg["norm_type"] == "cumulative_laplace_norm"
    fb_raw = np.ascontiguousarray(cm[:, : cfg["fb_freqs"], :])
    fb_in = cumulative_laplace_norm(fb_raw) if cumulative else offline_laplace_norm(fb_raw)
    act = {"Tanh": "tanh", "ReLU": "relu"}.get(cfg.get("fb_output_activate_function") or None)
    fb_out, fb_all = sequence_model_forward(fb_in, params, "fb_model.", 2, shared, act, dtype,
                                            proj="fc_output_layer")
    fb_tiled = np.tile(fb_out, (1, cfg["num_freqs"] // cfg["fb_freqs"], 1))
    cuts = [0] + list(cfg["freq_cutoffs"]) + [nf]
    coefs, sb_all = [], []
    for i, (ctr, nbr, df) in enumerate(zip(cfg["sb_num_center_freqs"], cfg["sb_num_neighbor_freqs"],
                                           cfg["sb_df_orders"])):
        x = subband_inputs(cm, fb_tiled, cuts[i], cuts[i + 1], ctr, nbr)  # [B*N, K, T]
        if cumulative:
            x = cumulative_laplace_norm(x)                               # count_time :173-204: per (b, n) row
        else:
            x = offline_laplace_norm(x.reshape(B, -1)).reshape(x.shape)  # MLF:475 over (N, 1, K, T) jointly
        sact = {"Tanh": "tanh", "ReLU": "relu"}.get(cfg.get("sb_output_activate_function") or None)
        out, all_out = sequence_model_forward(x, params, f"sb_model.sb_models.{i}.", 2, shared, sact, dtype,
                                              proj="fc_output_layer")
        coefs.append(subband_coef_layout(out, B, ctr, df, 1)[:, :, 0])  # MLF:257-263 (no speaker dim)
        sb_all.append(all_out)
    return coefs, fb_all, sb_all


def cirm_gsn_network(mag, params, cfg, dtype=np.float32):
    """CGN:226-230: all bins, one SequenceModel, 'b (c d s f) t -> b d s f t c'."""
    num_spks = cfg.get("num_spks", 2)
    cm = compress_mag(mag, cfg["fdrc"], dtype)
    act = cfg.get("output_activate_function")
    out, all_out = sequence_model_forward(cm, params, "fb_model.", cfg["num_layers"],
                                          cfg.get("shared_weights", False),
                                          act if isinstance(act, str) else None, dtype)
    B, P, T = out.shape
    d = cfg["df_order"]
    F = P // (2 * d * num_spks)
    o = out.reshape(B, 2, d, num_spks, F, T)
    coef = np.ascontiguousarray(np.transpose(o, (0, 2, 3, 4, 5, 1)))
    return coef, all_out


# --------------------------------------------------------------------------------------------
# f1: deep filtering (MSF:315-346)
# --------------------------------------------------------------------------------------------
def deepfiltering(spec, coef, order):
    """spec complex [B,F,T]; coef [B,df,S,F,T,2] -> complex [B,S,F,T].
    Y[f,t] = sum_d X[f, t-(df-1)+d] * C[d,f,t] with zero left padding."""
    B, F, T = spec.shape
    cc = coef[..., 0] + 1j * coef[..., 1]
    pad = np.concatenate([np.zeros((B, F, order - 1), dtype=spec.dtype), spec], axis=2)
    out = np.zeros((B, coef.shape[2], F, T), dtype=np.result_type(spec.dtype, cc.dtype))
    for d in range(order):
        out += pad[:, None, :, d: d + T] * cc[:, d]
    return out


# --------------------------------------------------------------------------------------------
# backward -- surrogate gradient, eval-mode BN; used by gradient parity tests
# --------------------------------------------------------------------------------------------
def gsn_layer_backward(x, h_trace, c_trace, d_h_out, w_ih, w_hh, bias, bn, shared, dtype=np.float64):
    """Backward of one GSULayer given its forward traces.

    x [T,R,K], h_trace/c_trace [T,R,H] (outputs of the forward), d_h_out [T,R,H] = dL/dh_t from above.
    Returns dict(dx, dw_ih, dw_hh, dbias).  Triangle surrogate max(0, 1-|c|) (ESN:95-101).
    """
    T, R, K = x.shape
    H = w_hh.shape[1]
    x = x.astype(dtype); h_trace = h_trace.astype(dtype); c_trace = c_trace.astype(dtype)
    w_ih = w_ih.astype(dtype); w_hh = w_hh.astype(dtype); bias = bias.astype(dtype)
    if bn is not None:
        bn_scale = bn["weight"].astype(dtype) / np.sqrt(bn["running_var"].astype(dtype) + BN_EPS)
        bn_shift = bn["bias"].astype(dtype) - bn["running_mean"].astype(dtype) * bn_scale
    dx = np.zeros_like(x)
    dw_ih = np.zeros_like(w_ih); dw_hh = np.zeros_like(w_hh); dbias = np.zeros_like(bias)
    dc_next = np.zeros((R, H), dtype); dh_next = np.zeros((R, H), dtype)
    for t in range(T - 1, -1, -1):
        h_prev = h_trace[t - 1] if t > 0 else np.zeros((R, H), dtype)
        c_prev = c_trace[t - 1] if t > 0 else np.zeros((R, H), dtype)
        z = x[t] @ w_ih.T + h_prev @ w_hh.T
        if shared:
            f_hat, g_hat = z + bias[:H], z + bias[H:]
        else:
            f_hat, g_hat = z[:, :H] + bias[:H], z[:, H:] + bias[H:]
        f = sigmoid(f_hat)
        dh = d_h_out[t] + dh_next
        dc = dc_next + dh * np.maximum(0.0, 1.0 - np.abs(c_trace[t]))
        dct = dc * bn_scale if bn is not None else dc
        df = dct * (c_prev - g_hat) * f * (1 - f)
        dg = dct * (1 - f)
        dc_next = dct * f
        if shared:
            dz = df + dg
            dh_next = dz @ w_hh
            dx[t] = dz @ w_ih
            dw_hh += dz.T @ h_prev
            dw_ih += dz.T @ x[t]
        else:
            dgates = np.concatenate([df, dg], axis=1)
            dh_next = dgates @ w_hh
            dx[t] = dgates @ w_ih
            dw_hh += dgates.T @ h_prev
            dw_ih += dgates.T @ x[t]
        dbias += np.concatenate([df.sum(0), dg.sum(0)])
    return dict(dx=dx, dw_ih=dw_ih, dw_hh=dw_hh, dbias=dbias)


# --------------------------------------------------------------------------------------------
# bookkeeping used by bench.py / DESIGN.md
# --------------------------------------------------------------------------------------------
def model_rows_and_shapes(cfg, batch):
    """[(name, rows, K, H, P)] for every sequence model of a surface-A config."""
    out = [("fb", batch, cfg["fb_input_size"], cfg["fb_hidden_size"], cfg["fb_proj_size"])]
    cuts = cfg["freq_cutoffs"]
    S = cfg.get("num_spks", 1)
    for i, (ctr, nbr, df) in enumerate(zip(cfg["center_freq_sizes"], cfg["neighbor_freq_sizes"],
                                           cfg["df_orders"])):
        n = (cuts[i + 1] - cuts[i]) // ctr
        out.append((f"sb{i}", batch * n, 2 * ctr + 2 * nbr, cfg["sb_hidden_size"], 2 * ctr * df * S))
    return out


def algorithmic_flops_per_frame(cfg):
    """Dense 2*MAC count per frame per utterance (SURVEY.md 8d formula)."""
    g = 1 if cfg.get("shared_weights", False) else 2
    total = 0
    for name, rows, K, H, P in model_rows_and_shapes(cfg, 1):
        L = cfg["fb_num_layers"] if name == "fb" else cfg["sb_num_layers"]
        total += rows * (g * 2 * H * K + (L - 1) * g * 2 * H * H + L * g * 2 * H * H + 2 * H * P)
    return total


# --------------------------------------------------------------------------------------------
# f4: SynOps / NeuronOps accounting (audiozen/metric.py:303-340)
# --------------------------------------------------------------------------------------------
def compute_synops(fb_all, sb_all, shared_weights=True):
    s = 0.0
    for trace in [fb_all] + list(sb_all):
        for i in range(1, len(trace) - 1):
            s += float((trace[i] > 0).mean()) * trace[i].shape[-1] * (trace[i + 1].shape[-1] + trace[i].shape[-1])
    return s if shared_weights else 2 * s


def compute_neuronops(fb_all, sb_all):
    return float(sum(t.shape[-1] for t in fb_all) + sum(t.shape[-1] for tr in sb_all for t in tr))
