"""Pins the CPU oracle (oracle/gsn_oracle.py) against the reference-generated golden fixtures.

Tolerances: the oracle runs numpy (OpenBLAS) fp32, the fixtures come from torch (MKL) fp32; only the
summation order differs, so real-valued tensors agree to ~1e-5 relative and spikes are identical on
these fixtures (verified when the fixtures were generated; a flip would need |c| < ~1e-6).
"""
import numpy as np
import pytest

from oracle import gsn_oracle as O
from oracle import synth
from tests.helpers import SURFACE_A, golden_params, load_golden, unpack


def _rel(a, b):
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30))


@pytest.mark.parametrize("name", SURFACE_A)
def test_network_free_running(name):
    g = load_golden(name)
    cfg = g["cfg"]
    params = golden_params(g)
    coefs, fb_all, sb_all = O.spiking_fullsubnet_network(g["mag"], params, cfg)
    Hf, Hs = cfg["fb_hidden_size"], cfg["sb_hidden_size"]
    assert _rel(fb_all[0], g["fb_xnorm"]) < 2e-5
    for l in range(cfg["fb_num_layers"]):
        assert np.array_equal(fb_all[1 + l], unpack(g[f"fb_h{l}"], Hf)), f"fb layer {l} spikes"
    assert _rel(fb_all[-1], g["fb_proj"]) < 2e-5
    for i in range(len(cfg["center_freq_sizes"])):
        assert _rel(sb_all[i][0], g[f"sb{i}_xnorm"]) < 5e-5
        for l in range(cfg["sb_num_layers"]):
            assert np.array_equal(sb_all[i][1 + l], unpack(g[f"sb{i}_h{l}"], Hs)), f"sb{i} layer {l}"
        assert coefs[i].shape == g[f"coef{i}"].shape
        # 1e-3 relative fp32 is the north_star tolerance; the oracle is far inside it
        assert _rel(coefs[i], g[f"coef{i}"]) < 1e-4


@pytest.mark.parametrize("name", ["tiny_shared_bn", "tiny_unshared_nobn", "tiny_spk2_tanh"])
def test_membrane_traces_teacher_forced(name):
    """Protocol P1: feed the reference's own (h_{t-1}, c_{t-1}) at every step; c_t must agree to 1e-5
    and h_t exactly wherever |c_t^ref| > 1e-5."""
    g = load_golden(name)
    cfg = g["cfg"]
    params = golden_params(g)
    models = [("fb_model.", cfg["fb_num_layers"], cfg["fb_hidden_size"], "fb")]
    models += [(f"sb_model.sb_models.{i}.", cfg["sb_num_layers"], cfg["sb_hidden_size"], f"sb{i}")
               for i in range(len(cfg["center_freq_sizes"]))]
    for prefix, L, H, tag in models:
        x = g[f"{tag}_xnorm"]
        teacher = []
        for l in range(L):
            c = g[f"c__{prefix}sequence_model.layers.{l}.cell"]
            h = unpack(g[f"{tag}_h{l}"], H)
            z = np.zeros_like(c[:1])
            teacher.append((np.concatenate([z, h[:-1]]), np.concatenate([z, c[:-1]])))
        _, all_out, cs = O.gsn_stack_forward(x, params, prefix + "sequence_model.", L,
                                             cfg["shared_weights"], return_c=True, teacher=teacher)
        # teacher forcing feeds layer l+1 with the oracle's own h of layer l, which equals the
        # reference's as long as spikes agree -- asserted below layer by layer
        for l in range(L):
            cref = g[f"c__{prefix}sequence_model.layers.{l}.cell"]
            assert np.abs(cs[l] - cref).max() <= 1e-5 * max(1.0, np.abs(cref).max())
            safe = np.abs(cref) > 1e-5
            assert np.array_equal(all_out[1 + l][safe], unpack(g[f"{tag}_h{l}"], H)[safe])


def test_cirm_gsn_network():
    g = load_golden("tiny_cirm")
    cfg = g["cfg"]
    params = synth.make_params_cirm(cfg, g["seed"])
    coef, all_out = O.cirm_gsn_network(g["mag"], params, cfg)
    for l in range(cfg["num_layers"]):
        assert np.array_equal(all_out[1 + l], unpack(g[f"fb_h{l}"], cfg["hidden_size"]))
    ref = g["fb_out"]  # [B, P, T]
    B, P, T = ref.shape
    d, S = cfg["df_order"], cfg["num_spks"]
    ref_coef = np.transpose(ref.reshape(B, 2, d, S, P // (2 * d * S), T), (0, 2, 3, 4, 5, 1))
    assert _rel(coef, ref_coef) < 1e-4


def test_fp64_oracle_agrees_with_fp32_reference_on_cfg1():
    """The fp64 arbiter reproduces the fp32 reference spikes on config 1 (no near-threshold events)."""
    g = load_golden("cfg1_baseline_m_1s")
    cfg = g["cfg"]
    coefs, fb_all, _ = O.spiking_fullsubnet_network(g["mag"], golden_params(g), cfg, dtype=np.float64)
    flips = float((fb_all[1] != unpack(g["fb_h0"], cfg["fb_hidden_size"])).mean())
    assert flips < 1e-3
    if flips == 0:
        assert _rel(coefs[2], g["coef2"]) < 1e-4


def test_unfold_index_errors_and_reflection():
    q = O.freq_unfold_index(0, 32, 4, 15, 256)
    assert q.shape == (8, 34) and q[0, 0] == 15 and q[0, 14] == 1 and q[0, 15] == 0
    q = O.freq_unfold_index(128, 256, 64, 15, 256)
    assert q[-1, -1] == 255 - 15 and q[-1, -16] == 255
    with pytest.raises(ValueError):
        O.freq_unfold_index(0, 30, 4, 15, 256)


def test_flops_table_matches_survey():
    # S 2 897 280, M 5 530 240, L 19 117 056, XL 10 503 424
    assert O.algorithmic_flops_per_frame(synth.CFG_S) == 2897280
    assert O.algorithmic_flops_per_frame(synth.CFG_M) == 5530240
    assert O.algorithmic_flops_per_frame(synth.CFG_L) == 19117056
    assert O.algorithmic_flops_per_frame(synth.CFG_XL) == 10503424


def test_deepfilter_and_backward_consistency():
    """deepfiltering restatement vs a direct loop; backward equations vs finite differences (fp64)."""
    rs = np.random.RandomState(0)
    spec = rs.standard_normal((2, 4, 9)) + 1j * rs.standard_normal((2, 4, 9))
    coef = rs.standard_normal((2, 3, 1, 4, 9, 2))
    y = O.deepfiltering(spec, coef, 3)
    t, f = 5, 2
    want = sum(spec[1, f, t - 2 + d] * (coef[1, d, 0, f, t, 0] + 1j * coef[1, d, 0, f, t, 1]) for d in range(3))
    assert abs(y[1, 0, f, t] - want) < 1e-12


@pytest.mark.parametrize("name", ["tiny_surface_b", "tiny_surface_b_cumnorm", "zoo_s_1s"])
def test_surface_b_network(name):
    """Surface B (`Separator`, offline laplace norm); `zoo_s_1s` uses the TRAINED model-zoo S checkpoint.
    With trained weights the path is chaotic: numpy-vs-MKL summation order may flip a
    spike, so the bound is the reference's own noise floor (3.5e-4 flips)."""
    from tests.helpers import load_golden_weights
    g = load_golden(name)
    cfg = g["cfg"]
    params = load_golden_weights(name) if name.startswith("zoo") else synth.make_params_b(cfg, g["seed"])
    coefs, fb_all, sb_all = O.separator_network(g["mag"], params, cfg)
    assert _rel(fb_all[0], g["fb_x"]) < 1e-5
    flips = total = 0
    for l in range(2):
        ref = unpack(g[f"fb_h{l}"], cfg["fb_hidden_size"])
        flips += (fb_all[1 + l] != ref).sum()
        total += ref.size
    for i in range(3):
        for l in range(2):
            ref = unpack(g[f"sb{i}_h{l}"], cfg["sb_hidden_size"])
            flips += (sb_all[i][1 + l] != ref).sum()
            total += ref.size
    assert flips / total <= 3.5e-4, f"{flips} of {total} spikes differ"
    if flips == 0:
        for i in range(3):
            assert coefs[i].shape == g[f"coef{i}"].shape
            assert _rel(coefs[i], g[f"coef{i}"]) < 1e-4


# ---- BASELINE-sized fixtures (T = 501 / 1 251; zoo-S, zoo-L, surface-A S) -----------------------------------------
def test_torch_port_is_bit_identical_to_the_reference_at_full_T():
    """oracle/gsn_oracle_torch.py is the `kind: "port"` CPU arm of bench.py: on surface-A S at 2 clips x 4 s (T = 501)
    it must reproduce the reference exactly -- every spike and every coefficient bit (same ATen kernels in the same
    order; the port only drops the per-frame weight.repeat)."""
    import torch
    from oracle import gsn_oracle_torch as OT
    from tests.helpers import compare_long, load_long
    g = load_long("cfgS_2x4s")
    torch.set_num_threads(4)
    coefs, fb_all, sb_all = OT.spiking_fullsubnet_network(torch.from_numpy(g["mag"]), OT.to_torch(g["params"]), g["cfg"])
    st = compare_long(g, coefs, fb_all, sb_all)
    assert st["flips"] == 0 and st["coef_rel"] == 0.0, st


@pytest.mark.parametrize("name", ["cfgS_2x4s", "zoo_s_2x4s", "zoo_l_2x4s"])
def test_numpy_oracle_free_running_at_T501(name):
    """The numpy oracle against the reference at T = 501 (numpy/OpenBLAS vs torch/MKL summation order only): flips
    bounded by the reference's own 1 + 1e-6 noise floor, coefficients per tests.helpers.assert_long."""
    from tests.helpers import assert_long, compare_long, load_long
    g = load_long(name)
    if g["surface"] == "A":
        coefs, fb_all, sb_all = O.spiking_fullsubnet_network(g["mag"], g["params"], g["cfg"])
    else:
        coefs, fb_all, sb_all = O.separator_network(g["mag"], g["params"], g["cfg"])
    st = compare_long(g, coefs, fb_all, sb_all)
    print(name, st)
    assert_long(st, name)


@pytest.mark.parametrize("name", ["zoo_s_2x4s", "zoo_l_2x4s"])
def test_numpy_oracle_block_teacher_forced_at_T501(name):
    """Trained zoo weights at T = 501, restarted from the reference's state every 32 frames
    (tests.helpers.block_forced_check): the oracle's cell arithmetic reproduces the reference's spikes up to isolated
    threshold events (< 1e-4 of all spikes; free-running, the same events would decorrelate whole utterances)."""
    from tests.helpers import block_forced_check, load_long
    g = load_long(name)
    coefs, fb_all, sb_all = O.separator_network(g["mag"], g["params"], g["cfg"])
    xs = {"fb": fb_all[0]}
    xs.update({f"sb{i}": al[0] for i, al in enumerate(sb_all)})

    def run_layer(inp, w_ih, w_hh, bias, bn, shared, h0, c0):
        hs = []
        h, c = h0, c0
        for t in range(inp.shape[0]):
            h, c = O.gsu_cell_step(inp[t], h, c, w_ih, w_hh, bias, bn, shared)
            hs.append(h)
        return np.stack(hs)

    st = block_forced_check(g, xs, run_layer)
    print(name, st)
    assert st["flips"] / st["total"] < 1e-4, st


@pytest.mark.parametrize("name", ["cfgS_2x4s", "zoo_l_2x4s"])
def test_numpy_oracle_first_divergence_sits_on_the_threshold(name):
    """The free-running numpy oracle leaves the reference (if at all) only at neurons whose reference membrane potential
    is within 1e-5 of the threshold, and its coefficients agree to 1e-4 before that (tests.helpers.divergence_audit)."""
    from tests.helpers import coef_rel_before_divergence, divergence_audit, load_long, reference_membrane
    g = load_long(name)
    if g["surface"] == "A":
        coefs, fb_all, sb_all = O.spiking_fullsubnet_network(g["mag"], g["params"], g["cfg"])
    else:
        coefs, fb_all, sb_all = O.separator_network(g["mag"], g["params"], g["cfg"])
    c_hat, _ = reference_membrane(g)
    st, div = divergence_audit(g, c_hat, fb_all, sb_all)
    st["coef_rel_before_divergence"] = coef_rel_before_divergence(g, coefs, div)
    print(name, st)
    assert st["bad_root_flips"] == 0, st
    assert st["coef_rel_before_divergence"] < 1e-4, st
