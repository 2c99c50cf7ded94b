"""torch-CPU port of the reference's path -- TEST/BENCH INFRASTRUCTURE ONLY (see gsn_oracle.py header).

Purpose: the `cpu_baseline` / `--impl reference` legs of bench.py.  The reference IS eager PyTorch on
CPU, so the honest CPU baseline executes the same ATen kernel sequence per frame
(mm, add, sigmoid, mul, batch_norm, ge; ESN:132-153) inside the same layer-outer / time-inner Python
loops (ESN:50-62, 75-81), with all host threads.  It is a restatement, not a copy: one `mm` per
operand with the gate halves sliced from it instead of the reference's per-frame `weight.repeat`
(ESN:134-136), which only makes this baseline FASTER than the reference (19 % of its CPU time,
SURVEY.md fact 2).  Checked against the numpy oracle in tests/test_oracle_golden.py.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import gsn_oracle as O


class _Triangle(torch.autograd.Function):
    """Heaviside forward, triangular surrogate max(0, 1-|c|) backward (restates ESN:84-101, gamma = 1)."""

    @staticmethod
    def forward(ctx, c):
        ctx.save_for_backward(c)
        return c.ge(0.0).float()

    @staticmethod
    def backward(ctx, grad):
        (c,) = ctx.saved_tensors
        return grad * (1.0 - c.abs()).clamp(min=0)


def _stack(x, params, prefix, L, shared, train_state=None):
    T, R, _ = x.shape
    trace = [x]
    cur = x
    for l in range(L):
        p = f"{prefix}layers.{l}.cell."
        w_ih, w_hh, bias = params[p + "weight_ih"], params[p + "weight_hh"], params[p + "bias_ih"]
        H = w_hh.shape[1]
        bn = p + "batchnorm.weight" in params
        h = torch.zeros(R, H)
        c = torch.zeros(R, H)
        w_ih_t, w_hh_t = w_ih.t(), w_hh.t()
        outs = []
        for t in range(T):
            z = torch.mm(cur[t], w_ih_t) + torch.mm(h, w_hh_t)
            if shared:
                f = torch.sigmoid(z + bias[:H])
                g = z + bias[H:]
            else:
                f = torch.sigmoid(z[:, :H] + bias[:H])
                g = z[:, H:] + bias[H:]
            c = f * c + (1 - f) * g
            if bn:
                c = F.batch_norm(c, params[p + "batchnorm.running_mean"], params[p + "batchnorm.running_var"],
                                 params[p + "batchnorm.weight"], params[p + "batchnorm.bias"],
                                 train_state is not None, 0.1, 1e-5)
            h = _Triangle.apply(c) if train_state is not None else c.ge(0.0).float()
            outs.append(h)
        cur = torch.stack(outs)
        trace.append(cur)
    return cur, trace


def _sequence_model(inp, params, prefix, L, shared, act, train_state=None):
    x = inp.permute(2, 0, 1)
    if prefix + "pre_layer_norm.weight" in params:
        x = F.layer_norm(x, (x.shape[-1],), params[prefix + "pre_layer_norm.weight"],
                         params[prefix + "pre_layer_norm.bias"])
    out, trace = _stack(x.contiguous(), params, prefix + "sequence_model.", L, shared, train_state)
    if prefix + "proj.weight" in params:
        out = F.linear(out, params[prefix + "proj.weight"], params[prefix + "proj.bias"])
    trace = trace + [out]
    if act == "tanh":
        out = torch.tanh(out)
    elif act == "sigmoid":
        out = torch.sigmoid(out)
    elif act == "relu":
        out = torch.relu(out)
    return out.permute(1, 2, 0), trace


def to_torch(params):
    import numpy as np
    return {k: torch.from_numpy(np.array(v)) for k, v in params.items()}


def spiking_fullsubnet_train_step(mag, params, cfg):
    """Forward in training mode (batch-statistics BatchNorm) + backward of a stand-in loss (mean square of the
    coefficients) through the surrogate gradient: the CPU cost of the path inside one training step."""
    leaf = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v.clone())
            for k, v in params.items()}
    coefs, _, _ = _network(mag, leaf, cfg, train_state=True)
    loss = sum(c.pow(2).mean() for c in coefs)
    loss.backward()
    return float(loss.detach())


@torch.no_grad()
def spiking_fullsubnet_network(mag, params, cfg):
    """mag torch [B,F,T] -> (coef list, fb_all, sb_all); MSF:434-447 on torch CPU."""
    return _network(mag, params, cfg, None)


def _network(mag, params, cfg, train_state):
    shared = cfg.get("shared_weights", False)
    S = cfg.get("num_spks", 1)
    cm = (mag ** cfg["fdrc"])[:, :-1, :]
    act = cfg.get("fb_output_activate_function")
    fb_out, fb_all = _sequence_model(cm[:, : cfg["fb_input_size"], :], params, "fb_model.",
                                     cfg["fb_num_layers"], shared, act if isinstance(act, str) else None, train_state)
    rep = (cfg["n_fft"] // 2 + 1) // cfg["fb_input_size"]
    fb_tiled = fb_out.repeat(1, rep, 1)
    B, Fq, T = cm.shape
    coefs, sb_all = [], []
    cuts = cfg["freq_cutoffs"]
    for i, (ctr, nbr, df) in enumerate(zip(cfg["center_freq_sizes"], cfg["neighbor_freq_sizes"],
                                           cfg["df_orders"])):
        qi = torch.from_numpy(O.freq_unfold_index(cuts[i], cuts[i + 1], ctr, nbr, Fq))
        qf = torch.from_numpy(O.freq_unfold_index(cuts[i], cuts[i + 1], ctr, 0, Fq))
        x = torch.cat([cm[:, qi, :], fb_tiled[:, qf, :]], dim=2)
        N = qi.shape[0]
        out, trace = _sequence_model(x.reshape(B * N, x.shape[2], T), params, f"sb_model.sb_models.{i}.",
                                     cfg["sb_num_layers"], shared, None, train_state)
        o = out.reshape(B, N, 2, ctr, df, S, T).permute(0, 4, 5, 1, 3, 6, 2).reshape(B, df, S, N * ctr, T, 2)
        coefs.append(o)
        sb_all.append(trace)
    return coefs, fb_all, sb_all
