"""Shared helpers for the parity tests (test infrastructure; may import oracle/)."""
import json
import os

import numpy as np

from oracle import synth

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SURFACE_A = ["tiny_shared_bn", "tiny_unshared_nobn", "tiny_spk2_tanh", "cfg1_baseline_m_1s"]


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    d = {k: z[k] for k in z.files}
    d["cfg"] = json.loads(str(d["cfg"]))
    d["seed"] = int(d["seed"])
    return d


def unpack(bits, H):
    """inverse of make_golden.pack: packed uint8 [..., ceil(H/8)] -> float32 {0,1} [..., H]."""
    return np.unpackbits(bits, axis=-1)[..., :H].astype(np.float32)


def golden_params(g):
    return synth.make_params(g["cfg"], g["seed"])


def spike_flip_stats(a, b):
    """fraction of differing spikes and index of the first frame with a difference (or -1)."""
    diff = a != b
    frac = float(diff.mean())
    first = -1
    if diff.any():
        first = int(np.argmax(diff.reshape(diff.shape[0], -1).any(axis=1)))
    return frac, first


def load_golden_weights(name):
    z = np.load(os.path.join(GOLDEN, name + "_weights.npz"), allow_pickle=False)
    return {k: z[k] for k in z.files}


def loss_waveforms(n=2, length=12000, seed=20220815):
    """Seeded (estimate, clean) waveform pair of the loss fixture (tests/golden/make_golden_loss.py)."""
    rs = np.random.RandomState(seed)
    t = np.arange(length) / 16000.0
    clean = (0.1 * np.sin(2 * np.pi * (200 + 300 * t) * t))[None, :].repeat(n, 0) * rs.uniform(0.5, 1.5, (n, 1))
    clean = clean.astype(np.float32)
    est = (clean + 0.03 * rs.standard_normal(clean.shape)).astype(np.float32)
    return est, clean


# ---- BASELINE-sized free-running fixtures (tests/golden/make_golden.py long) ------------------------------------
LONG = ["cfgS_2x4s", "zoo_s_2x4s", "zoo_s_1x10s", "zoo_l_2x4s", "zoo_l_1x10s"]


def load_long(name):
    g = load_golden(name)
    g["surface"] = str(g["surface"])
    if "weights_file" in g:
        z = np.load(os.path.join(GOLDEN, str(g["weights_file"]) + ".npz"), allow_pickle=False)
        g["params"] = {k: z[k] for k in z.files}
    else:
        g["params"] = synth.make_params(g["cfg"], 5)
    return g


def _np(t):
    return t.detach().cpu().numpy() if hasattr(t, "detach") else np.asarray(t)


def compare_long(g, coefs, fb_all, sb_all):
    """Spike flips (count, total, first frame) and coefficient max|delta| / max|ref| of a free-running result against a
    long fixture.  Coefficient tensors are compared on the frames the fixture stores (`coef_tail` last frames, or all)."""
    cfg = g["cfg"]
    Hf, Hs = cfg["fb_hidden_size"], cfg["sb_hidden_size"]
    flips = total = 0
    first = -1
    layers = [(fb_all[1 + l], unpack(g[f"fb_h{l}"], Hf)) for l in range(2)]
    for i in range(len(sb_all)):
        layers += [(sb_all[i][1 + l], unpack(g[f"sb{i}_h{l}"], Hs)) for l in range(2)]
    for got, ref in layers:
        d = _np(got) != ref
        flips += int(d.sum())
        total += ref.size
        if d.any():
            f = int(np.argmax(d.reshape(d.shape[0], -1).any(axis=1)))
            first = f if first < 0 else min(first, f)
    tail = int(g["coef_tail"])
    rel = 0.0
    for i, c in enumerate(coefs):
        c = _np(c)
        ref = g[f"coef{i}"]
        if tail:
            c = c[..., c.shape[-2] - tail:, :]
        assert c.shape == ref.shape, (c.shape, ref.shape)
        rel = max(rel, float(np.abs(c - ref).max() / (np.abs(ref).max() + 1e-30)))
    return {"flips": flips, "total": total, "first_flip_frame": first, "coef_rel": rel,
            "floor_flips": int(g["floor_flips"]), "floor_coef_rel": float(g["floor_coef_rel"])}


def assert_long(st, what):
    """Protocol P3 with the reference's own noise floor (reference vs itself with the input scaled by 1 + 1e-6,
    measured by make_golden.py in the same run): spike flips <= 4 x floor + 1e-5 of all spikes; coefficients within
    1e-4 relative when nothing flipped (everything after layer 1 is a function of the spikes only), otherwise
    within 4 x the floor's coefficient delta (+ the north-star 1e-3)."""
    floor = st["floor_flips"] / st["total"]
    frac = st["flips"] / st["total"]
    # one threshold event decorrelates the rest of its utterance, so the count is heavy-tailed: the floor itself moves
    # by 10x between perturbations of the same size (0 ... 1.2e-3 on zoo-S); the bound is 4x the worst of four
    assert frac <= 4 * floor + 1e-5, f"{what}: {st['flips']} of {st['total']} spikes differ (floor {st['floor_flips']})"
    if st["flips"] == 0:
        assert st["coef_rel"] < 1e-4, f"{what}: coefficients {st['coef_rel']:.2e} with identical spikes"
    else:
        assert st["coef_rel"] <= 4 * st["floor_coef_rel"] + 1e-3, \
            f"{what}: coefficients {st['coef_rel']:.2e} vs floor {st['floor_coef_rel']:.2e}"


def record_parity(name, st):
    """Append a parity record to gpurun_out/parity_counts.json (copied to profiles/ by hand after a GPU run)."""
    import json as _json
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    path = os.path.join(root, "gpurun_out", "parity_counts.json")
    try:
        os.makedirs(os.path.dirname(path), exist_ok=True)
        data = _json.load(open(path)) if os.path.exists(path) else {}
        data[name] = st
        _json.dump(data, open(path, "w"), indent=1, sort_keys=True)
    except OSError:
        pass


def block_forced_check(g, xs, run_layer):
    """Block teacher-forced protocol at full T (between P1 and P3 of SURVEY 8c): every SNAP = 32 frames the
    recurrence restarts from the REFERENCE'S OWN state (spikes of the previous frame from the fixture's trace, membrane
    from its `*_c{l}` snapshots) and is driven by the reference's own layer input (the stored spikes of the layer
    below; `xs[tag]` [T,R,K] for layer 0), so a threshold-chaos flip can only damage the rest of ITS 32-frame block of
    ITS row instead of the rest of the utterance.  All blocks of a layer are independent recurrences and run as ONE
    call: rows' = blocks x rows.

    run_layer(inp [SNAP, R', K], w_ih, w_hh, bias, bn, shared, h0 [R',H], c0 [R',H]) -> h [SNAP, R', H] (numpy).
    Returns {"flips", "total", "per_layer": {tag_l: (flips, total)}}."""
    cfg = g["cfg"]
    snap = int(g["snap"])
    shared = cfg.get("shared_weights", False)
    nb_models = len(cfg["sb_df_orders"] if "sb_df_orders" in cfg else cfg["df_orders"])
    models = [("fb", "fb_model.", cfg["fb_hidden_size"])]
    models += [(f"sb{i}", f"sb_model.sb_models.{i}.", cfg["sb_hidden_size"]) for i in range(nb_models)]
    out = {"flips": 0, "total": 0, "per_layer": {}}
    for tag, prefix, H in models:
        inp = np.asarray(xs[tag], dtype=np.float32)
        T, R, _ = inp.shape
        nb = (T + snap - 1) // snap
        for l in range(2):
            q = f"{prefix}sequence_model.layers.{l}.cell."
            p = g["params"]
            bn = None
            if q + "batchnorm.weight" in p:
                bn = {k: p[q + "batchnorm." + k] for k in ("weight", "bias", "running_mean", "running_var")}
            href = unpack(g[f"{tag}_h{l}"], H)                       # [T,R,H]
            csnap = g[f"{tag}_c{l}"]                                  # [T//snap, R, H]: c after frame k*snap-1
            pad = np.zeros((nb * snap,) + inp.shape[1:], dtype=np.float32)
            pad[:T] = inp
            blocked = pad.reshape(nb, snap, R, -1).transpose(1, 0, 2, 3).reshape(snap, nb * R, -1)
            h0 = np.zeros((nb, R, H), dtype=np.float32)
            c0 = np.zeros((nb, R, H), dtype=np.float32)
            for b in range(1, nb):
                h0[b] = href[b * snap - 1]
                c0[b] = csnap[b - 1]
            h = run_layer(np.ascontiguousarray(blocked), p[q + "weight_ih"], p[q + "weight_hh"], p[q + "bias_ih"], bn,
                          shared, h0.reshape(nb * R, H), c0.reshape(nb * R, H))
            h = np.asarray(h).reshape(snap, nb, R, H).transpose(1, 0, 2, 3).reshape(nb * snap, R, H)[:T]
            f = int((h != href).sum())
            out["per_layer"][f"{tag}_{l}"] = (f, href.size)
            out["flips"] += f
            out["total"] += href.size
            inp = href
    return out


# ---- "first divergence must sit on the threshold" audit ----------------------------------------------------------
# Two fp32 implementations that sum h.W^T in different orders differ by ~1e-7 in the membrane potential, and the density
# of |c| near 0 is ~1.6 per unit: a spike flips about once per 5e6 neuron-steps even between the
# reference and itself on another BLAS, and then decorrelates the rest of its utterance (the fixture cfgS_2x4s holds a
# membrane potential of EXACTLY 0.0 at frame 57).  So free-running spike counts cannot be a pass/fail quantity at
# T = 501 / 1 251; what can be is WHERE each trajectory first leaves the reference: only at a neuron whose reference
# membrane potential is within 1e-5 of the threshold.
def _models_of(cfg):
    nb = len(cfg["sb_df_orders"] if "sb_df_orders" in cfg else cfg["df_orders"])
    return [("fb", "fb_model.", cfg["fb_hidden_size"])] + \
        [(f"sb{i}", f"sb_model.sb_models.{i}.", cfg["sb_hidden_size"]) for i in range(nb)]


def layer0_inputs_from_reference(g):
    """Layer-0 inputs x [T,R,K] of every sequence model, computed with the oracle's front end from the fixture's
    magnitude and the REFERENCE's full-band spikes (so they stay valid after the oracle's own first flip)."""
    from oracle import gsn_oracle as O
    cfg, p = g["cfg"], g["params"]
    surf_b = g["surface"] == "B"
    cm = O.compress_mag(g["mag"], cfg["fdrc"])[:, :-1, :]
    B, F, T = cm.shape
    Hf = cfg["fb_hidden_size"]
    h_last = unpack(g["fb_h1"], Hf)                                   # [T,B,Hf] reference spikes, last fb layer
    xs = {}
    if surf_b:
        fbk = cfg["fb_freqs"]
        xs["fb"] = np.ascontiguousarray(np.transpose(O.offline_laplace_norm(np.ascontiguousarray(cm[:, :fbk])), (2, 0, 1)))
        w, b = p["fb_model.fc_output_layer.weight"], p["fb_model.fc_output_layer.bias"]
        act = {"Tanh": "tanh", "ReLU": "relu"}.get(cfg.get("fb_output_activate_function") or None)
        rep = cfg["num_freqs"] // fbk
        cuts = [0] + list(cfg["freq_cutoffs"]) + [F]
        ctrs, nbrs = cfg["sb_num_center_freqs"], cfg["sb_num_neighbor_freqs"]
    else:
        fbk = cfg["fb_input_size"]
        x = np.ascontiguousarray(np.transpose(cm[:, :fbk], (2, 0, 1)))
        if "fb_model.pre_layer_norm.weight" in p:
            x = O.layer_norm(x, p["fb_model.pre_layer_norm.weight"], p["fb_model.pre_layer_norm.bias"]).astype(np.float32)
        xs["fb"] = x
        w, b = p["fb_model.proj.weight"], p["fb_model.proj.bias"]
        act = cfg.get("fb_output_activate_function")
        rep = (cfg["n_fft"] // 2 + 1) // fbk
        cuts = cfg["freq_cutoffs"]
        ctrs, nbrs = cfg["center_freq_sizes"], cfg["neighbor_freq_sizes"]
    fb_out = h_last @ w.T + b                                          # [T,B,P]
    if isinstance(act, str) and act in O._ACT:
        fb_out = O._ACT[act](fb_out)
    fb_tiled = np.tile(np.transpose(fb_out, (1, 2, 0)), (1, rep, 1)).astype(np.float32)
    for i, (ctr, nbr) in enumerate(zip(ctrs, nbrs)):
        x = O.subband_inputs(cm, fb_tiled, cuts[i], cuts[i + 1], ctr, nbr)     # [B*N, K, T]
        if surf_b:
            x = O.offline_laplace_norm(x.reshape(B, -1)).reshape(x.shape)
            x = np.transpose(x, (2, 0, 1))
        else:
            x = np.transpose(x, (2, 0, 1))
            q = f"sb_model.sb_models.{i}.pre_layer_norm."
            if q + "weight" in p:
                x = O.layer_norm(x, p[q + "weight"], p[q + "bias"])
        xs[f"sb{i}"] = np.ascontiguousarray(x, dtype=np.float32)
    return xs


def reference_membrane(g):
    """c_hat[f"{tag}_{l}"] [T,R,H]: the reference's membrane potentials reconstructed by the numpy oracle under the
    block teacher-forced protocol (restart from the fixture's snapshots every SNAP frames, reference spikes as layer
    inputs): equal to the reference's own values to ~1e-7 except inside the few blocks where the oracle itself flips.
    The returned stats also carry "noise_abs_c": the largest |c| at which the numpy fp32 oracle ITSELF first leaves the
    reference inside a block -- the distance from the threshold at which two independent fp32 implementations of this
    model (trained weights: folded BatchNorm scales up to ~10^2) can disagree about a spike."""
    from oracle import gsn_oracle as O
    c_hat = {}
    order, order_h = [], []

    def run_layer(inp, w_ih, w_hh, bias, bn, shared, h0, c0):
        hs, cs = [], []
        h, c = h0, c0
        for t in range(inp.shape[0]):
            h, c = O.gsu_cell_step(inp[t], h, c, w_ih, w_hh, bias, bn, shared)
            hs.append(h)
            cs.append(c)
        order.append(np.stack(cs))
        order_h.append(np.stack(hs))
        return order_h[-1]

    st = block_forced_check(g, layer0_inputs_from_reference(g), run_layer)
    snap = int(g["snap"])
    k = 0
    noise = 0.0
    for tag, _, H in _models_of(g["cfg"]):
        for l in range(2):
            cs, hs = order[k], order_h[k]
            k += 1
            href = unpack(g[f"{tag}_h{l}"], H)
            T = href.shape[0]
            R = cs.shape[1] // ((T + snap - 1) // snap)
            nb = cs.shape[1] // R
            c_hat[f"{tag}_{l}"] = cs.reshape(snap, nb, R, H).transpose(1, 0, 2, 3).reshape(nb * snap, R, H)[:T]
            hb = hs.reshape(snap, nb, R, H).transpose(1, 0, 2, 3).reshape(nb * snap, R, H)[:T]
            d = hb != href
            if d.any():
                pad = np.zeros((nb * snap, R, H), dtype=bool)
                pad[:T] = d
                blk = pad.reshape(nb, snap, R, H)
                anyrow = blk.any(axis=3)                                   # [nb, snap, R]
                first = anyrow.argmax(axis=1)                              # [nb, R] first flipped frame of the block
                cb = np.zeros((nb * snap, R, H), dtype=np.float32)
                cb[:T] = c_hat[f"{tag}_{l}"]
                cb = cb.reshape(nb, snap, R, H)
                for b, r in zip(*np.nonzero(anyrow.any(axis=1))):
                    t = first[b, r]
                    noise = max(noise, float(np.abs(cb[b, t, r][blk[b, t, r]]).max()))
    st["noise_abs_c"] = noise
    return c_hat, st


def membrane_noise(g):
    """D[f"{tag}_{l}"] [T,R,H] = |c_fp32 - c_fp64| of the numpy oracle run CONTINUOUSLY over all frames with the
    reference's spike history forced (h_{t-1} and the layer input from the fixture, its own membrane potential carried):
    how far two valid floating-point evaluations of the SAME equations on the SAME spike history drift apart.  The
    leaky recursion c_t = a (f c_{t-1} + (1-f) g) + b is expanding wherever a f > 1 (folded BatchNorm scales of the
    trained zoo-L checkpoint reach 2.1), so this drift is not bounded by an ulp: on zoo-L it reaches 1e-5 after 100
    frames and 1e-2 after 431 with identical spikes.  A free-running trajectory can therefore leave the reference at any
    neuron whose reference membrane potential is within a few D of the threshold."""
    from oracle import gsn_oracle as O
    cfg, p = g["cfg"], g["params"]
    shared = cfg.get("shared_weights", False)
    xs = layer0_inputs_from_reference(g)
    out = {}
    for tag, prefix, H in _models_of(cfg):
        inp = np.asarray(xs[tag], dtype=np.float32)
        for l in range(2):
            q = f"{prefix}sequence_model.layers.{l}.cell."
            bn = None
            if q + "batchnorm.weight" in p:
                bn = {k: p[q + "batchnorm." + k] for k in ("weight", "bias", "running_mean", "running_var")}
            href = unpack(g[f"{tag}_h{l}"], H)
            T, R, _ = href.shape
            traces = []
            for dt in (np.float32, np.float64):
                w_ih, w_hh, bias = (p[q + k].astype(dt) for k in ("weight_ih", "weight_hh", "bias_ih"))
                bnd = None if bn is None else {k: v.astype(dt) for k, v in bn.items()}
                c = np.zeros((R, H), dt)
                cs = np.empty((T, R, H), np.float64)
                for t in range(T):
                    hin = href[t - 1].astype(dt) if t > 0 else np.zeros((R, H), dt)
                    _, c = O.gsu_cell_step(inp[t].astype(dt), hin, c, w_ih, w_hh, bias, bnd, shared)
                    cs[t] = c
                traces.append(cs)
            out[f"{tag}_{l}"] = np.abs(traces[0] - traces[1]).astype(np.float32)
            inp = href
    return out


def _first_true(a):
    """a [T, ...] bool -> first index along axis 0 where any is True, per trailing index of axis 1 (rows); T if never.
    a is [T,R,H] -> returns [R]."""
    anyrow = a.any(axis=2)                      # [T,R]
    T = a.shape[0]
    first = np.where(anyrow.any(axis=0), anyrow.argmax(axis=0), T)
    return first


def divergence_audit(g, c_hat, fb_all, sb_all, thr=1e-5, noise=None, noise_factor=8.0):
    """For every row trajectory of a FREE-RUNNING result: the frame at which it first leaves the reference, and whether
    the spikes that flipped there ("root" flips: not explained by an earlier flip of the same row, of the layer below in
    the same row, or of the utterance's full-band model) belong to neurons whose reference membrane potential is within
    `thr` of the threshold.  Returns counts and per-row divergence frames {tag: [R]} (T = never)."""
    cfg = g["cfg"]
    out = {"root_flips": 0, "bad_root_flips": 0, "worst_root_abs_c": 0.0, "rows": 0, "rows_diverged": 0}
    div = {}

    def audit(tag, H, got, limit):
        """limit [R]: frames >= limit[r] are already excused for row r (upstream divergence)."""
        d0 = _np(got[1]) != unpack(g[f"{tag}_h0"], H)
        d1 = _np(got[2]) != unpack(g[f"{tag}_h1"], H)
        T, R, _ = d0.shape
        t0, t1 = _first_true(d0), _first_true(d1)
        for r in range(R):
            lim = min(limit[r], T)
            roots = []
            if t0[r] < lim:
                roots.append((0, t0[r]))
            if t1[r] < min(lim, t0[r]):          # same frame as a layer-0 flip: explained by it
                roots.append((1, t1[r]))
            for l, t in roots:
                d = (d0 if l == 0 else d1)[t, r]
                cabs = np.abs(c_hat[f"{tag}_{l}"][t, r][d])
                # "on the threshold": within thr, or within noise_factor x the fp32-vs-fp64 drift of that very neuron
                # at that frame under the reference's own spike history (membrane_noise)
                lim_c = np.full(cabs.shape, thr)
                if noise is not None:
                    lim_c = np.maximum(lim_c, noise_factor * noise[f"{tag}_{l}"][t, r][d])
                    out["worst_root_c_over_noise"] = max(out.get("worst_root_c_over_noise", 0.0),
                                                         float((cabs / np.maximum(lim_c, 1e-30)).max()))
                cabs = np.where(cabs < lim_c, 0.0, cabs)      # excused roots count as 0 below
                out["root_flips"] += int(d.sum())
                out["bad_root_flips"] += int((cabs >= thr).sum())
                out["worst_root_abs_c"] = max(out["worst_root_abs_c"], float(cabs.max()))
                if (cabs >= thr).any():
                    out.setdefault("bad_roots", []).append(
                        {"model": tag, "layer": l, "row": int(r), "frame": int(t), "limit": int(min(lim, 10 ** 9)),
                         "abs_c": [float(v) for v in cabs[cabs >= thr][:4]]})
        first = np.minimum(np.minimum(t0, t1), limit)
        out["rows"] += R
        out["rows_diverged"] += int((first < T).sum())
        return first, T

    fb_first, T = audit("fb", cfg["fb_hidden_size"], fb_all, np.full(_np(fb_all[1]).shape[1], 10 ** 9))
    div["fb"] = np.minimum(fb_first, T)
    B = len(fb_first)
    for i in range(len(sb_all)):
        R = _np(sb_all[i][1]).shape[1]
        N = R // B
        lim = np.repeat(div["fb"], N)
        if str(g.get("surface", "A")) == "B":
            # surface B normalises the full-band part of the sub-band input by its mean over ALL frames of the utterance
            # (offline_laplace_norm, model_low_freq.py:146-171): a full-band divergence at frame t moves the sub-band
            # input of EVERY frame by ~1e-4 relative, so it excuses the utterance's sub-band rows from frame 0 on
            lim = np.where(lim < T, 0, lim)
        lim = np.where(lim >= T, 10 ** 9, lim)
        first, _ = audit(f"sb{i}", cfg["sb_hidden_size"], sb_all[i], lim)
        div[f"sb{i}"] = np.minimum(first, T)
    out["frames"] = int(T)
    out["first_divergence_frame"] = int(min(int(v.min()) for v in div.values()))
    return out, div


def coef_rel_before_divergence(g, coefs, div):
    """max|coef - ref| / max|ref| restricted to (sub-band row, frame) pairs before that row left the reference."""
    cfg = g["cfg"]
    ctrs = cfg["sb_num_center_freqs"] if "sb_num_center_freqs" in cfg else cfg["center_freq_sizes"]
    tail = int(g["coef_tail"])
    worst = 0.0
    for i, c in enumerate(coefs):
        c = _np(c)
        ref = g[f"coef{i}"]
        Tfull = c.shape[-2]
        t_off = Tfull - tail if tail else 0
        c = c[..., t_off:, :]
        # coefficient tensors end with (..., F_band = N*ctr, T, 2); batch first
        B = c.shape[0]
        d = div[f"sb{i}"].reshape(B, -1)                               # [B,N]
        N = d.shape[1]
        tt = np.arange(t_off, Tfull)
        ok = tt[None, None, :] < d[:, :, None]                         # [B,N,T']
        ok = np.repeat(ok, ctrs[i], axis=1)                            # [B,F_band,T']
        shape = [B] + [1] * (c.ndim - 4) + [ok.shape[1], ok.shape[2], 1]
        okb = np.broadcast_to(ok.reshape(shape), c.shape)
        if okb.any():
            worst = max(worst, float(np.abs(c - ref)[okb].max() / (np.abs(ref).max() + 1e-30)))
    return worst
