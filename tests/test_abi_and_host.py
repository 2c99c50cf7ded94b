"""CPU-side checks: the C-ABI library loads and exports every symbol include/gsn_b200.h declares,
the facade keeps the reference's state_dict contract, and the product refuses to run without CUDA."""
import re

import numpy as np
import pytest
import torch

from oracle import synth
from spiking_fullsubnet_b200 import CirmGSN, SpikingFullSubNet, _lib
from spiking_fullsubnet_b200.modeling import coef_layout
from tests.helpers import load_golden


def test_header_symbols_all_exported_and_bound():
    text = open(_lib.HEADER_PATH).read()
    declared = set(re.findall(r"GSN_API\s+[\w\s\*]+?\b(gsn_\w+)\s*\(", text))
    assert declared, "no GSN_API declarations found"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name)
    assert lib.gsn_abi_version() == 1


def test_argument_validation_without_gpu():
    lib = _lib.load()
    rc = lib.gsn_compress_mag(None, None, 1, 1, 1, 1, 0.5, None)
    assert rc == _lib.GSN_EINVAL and b"null" in lib.gsn_last_error()
    with pytest.raises(ValueError):
        _lib.check(rc)
    assert lib.gsn_layer_recurrence_workspace_bytes(8, 64, 1, _lib.BACKEND_SIMT) == 64 * 64 * 4
    assert lib.gsn_layer_recurrence_workspace_bytes(8, 64, 0, _lib.BACKEND_SIMT) == 64 * 128 * 4


@pytest.mark.parametrize("name", ["S", "M", "L", "XL"])
def test_state_dict_contract(name):
    """Same parameter/buffer names and shapes as the reference (so zoo checkpoints load strict=True);
    parameter counts of SURVEY.md section 6 (S 520 920 ... are surface B; surface A M = 954 412)."""
    cfg = synth.CONFIGS[name]
    model = SpikingFullSubNet(**cfg)
    params = synth.make_params(cfg, 1)
    sd = model.state_dict()
    assert set(sd) == set(params)
    for k, v in params.items():
        assert tuple(sd[k].shape) == tuple(np.shape(v)), k
    model.load_state_dict({k: torch.from_numpy(np.array(v)) for k, v in params.items()}, strict=True)
    if name == "M":
        assert sum(p.numel() for p in model.parameters()) == 954412


def test_golden_state_dict_loads():
    g = load_golden("cfg1_baseline_m_1s")
    m = SpikingFullSubNet(**g["cfg"])
    m.load_state_dict({k: torch.from_numpy(np.array(v)) for k, v in synth.make_params(g["cfg"], g["seed"]).items()})
    c = CirmGSN(**synth.CFG_CIRM)
    assert set(c.state_dict()) == set(synth.make_params_cirm(synth.CFG_CIRM, 0))


def test_init_distribution_matches_reference_rule():
    m = SpikingFullSubNet(**synth.CFG_S)
    cell = m.fb_model.sequence_model.layers[0].cell
    bound = 1.0 / np.sqrt(240)
    for p in (cell.weight_ih, cell.weight_hh, cell.bias_ih):
        assert p.abs().max() <= bound and p.abs().max() > 0.9 * bound
    assert cell.weight_ih.shape == (240, 64) and cell.bias_ih.shape == (480,)


def test_no_cpu_fallback():
    m = SpikingFullSubNet(**synth.tiny_cfg()).eval()
    with pytest.raises(RuntimeError, match="no CPU path"):
        m(torch.zeros(1, 320))
    with pytest.raises(NotImplementedError):
        SpikingFullSubNet(**dict(synth.tiny_cfg(), sequence_model="LSTM"))
    with pytest.raises(NotImplementedError):
        SpikingFullSubNet(**dict(synth.tiny_cfg(), sequence_model="GRU"))


def test_coef_layout_matches_oracle_index_map():
    from oracle import gsn_oracle as O
    rs = np.random.RandomState(0)
    B, N, ctr, df, S, T = 2, 3, 4, 3, 2, 5
    P = 2 * ctr * df * S
    proj = rs.standard_normal((T, B * N, P)).astype(np.float32)
    want = O.subband_coef_layout(np.transpose(proj, (1, 2, 0)), B, ctr, df, S)
    got = coef_layout(torch.from_numpy(proj), B, N, df, S).numpy()
    assert np.array_equal(got, want)


def test_shipped_shapes_run_on_tcgen05():
    """No silent fp32-SIMT fallback for any shipped model size (S/M/L/XL, cirm_gsn) at bench batch sizes."""
    from oracle import gsn_oracle as O
    from spiking_fullsubnet_b200 import ops
    for name, cfg in synth.CONFIGS.items():
        for B in (1, 32, 64):
            for _, rows, _, H, _ in O.model_rows_and_shapes(cfg, B):
                assert ops.pick_backend(rows, H, cfg["shared_weights"]) == "tcgen05", (name, B, rows, H)
    assert ops.pick_backend(32, 268, True) == "tcgen05"  # cirm_gsn default
    assert ops.pick_backend(480, 512, True) == "tcgen05_i8"  # H > 320: bf16x3 planes do not fit TMEM, int8 do


def test_synops_accounting_matches_reference_formula():
    """metrics.compute_synops / compute_neuronops (audiozen/metric.py:303-340) on the golden traces."""
    from oracle import gsn_oracle as O
    from spiking_fullsubnet_b200 import metrics
    from tests.helpers import golden_params
    g = load_golden("tiny_shared_bn")
    _, fb_all, sb_all = O.spiking_fullsubnet_network(g["mag"], golden_params(g), g["cfg"])
    want_s, want_n = O.compute_synops(fb_all, sb_all), O.compute_neuronops(fb_all, sb_all)
    tf = [torch.from_numpy(t) for t in fb_all]
    ts = [[torch.from_numpy(t) for t in tr] for tr in sb_all]
    assert abs(metrics.compute_synops(tf, ts) - want_s) <= 1e-4 * want_s
    assert metrics.compute_neuronops(tf, ts) == want_n
    assert abs(metrics.compute_synops(tf, ts, shared_weights=False) - 2 * want_s) <= 2e-4 * want_s


def test_sync_free_istft_matches_torch_istft():
    """The inference path's iSTFT (no host-synchronising envelope check, graph-capturable) against torch.istft."""
    from spiking_fullsubnet_b200.modeling import _istft, _istft_nosync, _stft
    torch.manual_seed(0)
    for n, h, L in [(512, 128, 16000), (64, 16, 624), (256, 64, 8000), (512, 128, 16123)]:
        x = torch.randn(2, L)
        spec = _stft(x, n, h, n) * (1 + 0.1 * torch.randn(2, n // 2 + 1, 1 + L // h))
        a, b = _istft(spec, n, h, n, L), _istft_nosync(spec, n, h, n, L)
        assert a.shape == b.shape and float((a - b).abs().max() / a.abs().max()) < 2e-6


def test_launch_options_validate_without_gpu():
    """gsn_set_option: known options are accepted (and reset), unknown ones are GSN_EINVAL with a message."""
    lib = _lib.load()
    for opt in (_lib.OPT_PDL, _lib.OPT_F32_MAX_CTAS):
        assert lib.gsn_set_option(opt, 1) == _lib.GSN_OK
        assert lib.gsn_set_option(opt, 0) == _lib.GSN_OK
    rc = lib.gsn_set_option(99, 1)
    assert rc == _lib.GSN_EINVAL and b"unknown option" in lib.gsn_last_error()
    rc = lib.gsn_pack_spikes(None, None, 1, 1, None)
    assert rc == _lib.GSN_EINVAL
    rc = lib.gsn_linear_spike_bits(None, None, None, None, None, 0, 1, 16, 16, 0, None)
    assert rc == _lib.GSN_EINVAL


def test_wavefront_chunk_bounds_cover_all_frames(monkeypatch):
    """Frame chunks of the wavefront schedule: contiguous, non-empty, cover [0, T) exactly; more chunks than
    frames collapses to one frame per chunk; GSN_WF_WEIGHTS reshapes them without losing frames."""
    from spiking_fullsubnet_b200.modeling import _chunk_bounds
    monkeypatch.delenv("GSN_WF_WEIGHTS", raising=False)
    for T, n in [(501, 12), (501, 1), (126, 16), (5, 12), (1251, 16), (1, 3)]:
        b = _chunk_bounds(T, n)
        assert b[0][0] == 0 and b[-1][1] == T and len(b) == min(n, T)
        assert all(a1 == b0 for (_, a1), (b0, _) in zip(b[:-1], b[1:])) and all(hi > lo for lo, hi in b)
        assert max(hi - lo for lo, hi in b) - min(hi - lo for lo, hi in b) <= 1
    monkeypatch.setenv("GSN_WF_WEIGHTS", "1,2,4,2,1")
    b = _chunk_bounds(501, 12)
    assert len(b) == 5 and b[0][0] == 0 and b[-1][1] == 501 and (b[2][1] - b[2][0]) > 3 * (b[0][1] - b[0][0])


def test_spike_bits_layout_helpers_on_cpu():
    """ops.spike_bits_buffer sizes the packed trace as ceil(H/32) int32 words per row."""
    from spiking_fullsubnet_b200 import ops
    for H, W in [(16, 1), (32, 1), (33, 2), (160, 5), (240, 8), (320, 10)]:
        assert tuple(ops.spike_bits_buffer((7, 3), H, "cpu").shape) == (7, 3, W)


def test_hooks_on_gsn_submodules_are_reported_once():
    """The reference's debug mode hooks every sub-module (audiozen/trainer.py:354-356); the fused kernels never call
    GSUCell.forward per frame, so the model says so (once) instead of silently skipping the hooks."""
    import warnings
    import torch
    from oracle import synth
    from spiking_fullsubnet_b200 import SpikingFullSubNet
    m = SpikingFullSubNet(**synth.tiny_cfg())
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        m._warn_cell_hooks()
    assert not w
    del m.__dict__["_cell_hooks_checked"]
    cell = next(mod for mod in m.modules() if type(mod).__name__ == "GSUCell")
    cell.register_forward_hook(lambda mod, inp, out: None)
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        m._warn_cell_hooks()
        m._warn_cell_hooks()
    assert len(w) == 1 and "forward hooks" in str(w[0].message)
