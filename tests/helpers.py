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
