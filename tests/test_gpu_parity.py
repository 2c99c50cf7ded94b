"""GPU parity tests: the CUDA path (through the C ABI) against the golden fixtures and the CPU oracle.

Tolerances: north_star asks for outputs within 1e-3 relative fp32 of the reference CPU path; real-valued
tensors are checked at 1e-3 * max|ref| (they agree to ~1e-5 when no spike flips) and spikes must be
IDENTICAL on the golden fixtures (protocol P2).
"""
import numpy as np
import pytest
import torch

from oracle import gsn_oracle as O
from oracle import synth
from spiking_fullsubnet_b200 import CirmGSN, SpikingFullSubNet, ops
from tests.helpers import SURFACE_A, golden_params, load_golden, spike_flip_stats, unpack

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
BACKENDS = ["simt", "auto", "tcgen05_i8"]


def _rel(a, b):
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30))


def _model(cfg, params, backend="auto"):
    m = SpikingFullSubNet(**cfg)
    m.load_state_dict({k: torch.from_numpy(np.array(v)) for k, v in params.items()}, strict=True)
    return m.eval().to(DEV).set_backend(backend)


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


@pytest.mark.parametrize("backend", BACKENDS)
@pytest.mark.parametrize("name", SURFACE_A)
def test_network_vs_golden(name, backend):
    g = load_golden(name)
    cfg = g["cfg"]
    m = _model(cfg, golden_params(g), backend)
    with torch.no_grad():
        coefs, fb_all, sb_all = m.coefficients(_t(g["mag"]))
    Hf, Hs = cfg["fb_hidden_size"], cfg["sb_hidden_size"]
    assert _rel(fb_all[0].cpu().numpy(), g["fb_xnorm"]) < 1e-4
    for l in range(cfg["fb_num_layers"]):
        frac, first = spike_flip_stats(fb_all[1 + l].cpu().numpy(), unpack(g[f"fb_h{l}"], Hf))
        assert frac == 0, f"fb layer {l}: {frac:.2e} spikes flipped, first at frame {first}"
    assert _rel(fb_all[-1].cpu().numpy(), g["fb_proj"]) < 1e-3
    for i in range(len(cfg["center_freq_sizes"])):
        assert _rel(sb_all[i][0].cpu().numpy(), g[f"sb{i}_xnorm"]) < 1e-4
        for l in range(cfg["sb_num_layers"]):
            frac, first = spike_flip_stats(sb_all[i][1 + l].cpu().numpy(), unpack(g[f"sb{i}_h{l}"], Hs))
            assert frac == 0, f"sb{i} layer {l}: {frac:.2e} spikes flipped, first at frame {first}"
        assert coefs[i].shape == g[f"coef{i}"].shape
        assert _rel(coefs[i].cpu().numpy(), g[f"coef{i}"]) < 1e-3  # the "cIRM max|delta|" bar


@pytest.mark.parametrize("name", SURFACE_A)
def test_full_forward_vs_golden(name):
    """forward(wave) end to end: STFT -> network -> deep filter -> iSTFT, same return tuple."""
    g = load_golden(name)
    cfg = g["cfg"]
    m = _model(cfg, golden_params(g))
    with torch.no_grad():
        out = m(_t(g["wave"]))
    if cfg["num_spks"] > 1:
        assert len(out) == 3
        enh_y = out[0]
    else:
        assert len(out) == 4
        enh_y, enh_mag = out[0], out[1]
        assert _rel(enh_mag.cpu().numpy(), g["enh_mag"]) < 1e-3
    assert enh_y.shape == g["enh_y"].shape
    assert _rel(enh_y.cpu().numpy(), g["enh_y"]) < 1e-3
    L = cfg["fb_num_layers"]
    assert len(out[-2]) == L + 2 and len(out[-1]) == len(cfg["center_freq_sizes"])


@pytest.mark.parametrize("backend", BACKENDS)
@pytest.mark.parametrize("name", ["tiny_shared_bn", "tiny_unshared_nobn", "tiny_spk2_tanh"])
def test_recurrence_teacher_forced(name, backend):
    """Protocol P1 per layer: one frame at a time from the reference's own (x_t, h_{t-1}, c_{t-1})."""
    g = load_golden(name)
    cfg = g["cfg"]
    params = golden_params(g)
    shared = cfg["shared_weights"]
    models = [("fb_model.", cfg["fb_num_layers"], cfg["fb_hidden_size"], "fb")]
    models += [(f"sb_model.sb_models.{i}.", cfg["sb_num_layers"], cfg["sb_hidden_size"], f"sb{i}")
               for i in range(len(cfg["center_freq_sizes"]))]
    for prefix, L, H, tag in models:
        inp = g[f"{tag}_xnorm"]
        for l in range(L):
            q = f"{prefix}sequence_model.layers.{l}.cell."
            cref = g["c__" + q[:-1]]
            href = unpack(g[f"{tag}_h{l}"], H)
            T, R, _ = cref.shape
            w_ih, w_hh, bias = _t(params[q + "weight_ih"]), _t(params[q + "weight_hh"]), _t(params[q + "bias_ih"])
            a = b = None
            if cfg["bn"]:
                inv = 1.0 / np.sqrt(params[q + "batchnorm.running_var"] + np.float32(1e-5))
                al = (inv * params[q + "batchnorm.weight"]).astype(np.float32)
                a, b = _t(al), _t((params[q + "batchnorm.bias"] - params[q + "batchnorm.running_mean"] * al).astype(np.float32))
            # all T single-frame problems at once: fold time into rows (each row is independent)
            hprev = np.concatenate([np.zeros_like(href[:1]), href[:-1]]).reshape(1, T * R, H)
            cprev = np.concatenate([np.zeros_like(cref[:1]), cref[:-1]]).reshape(1, T * R, H)
            xproj = ops.linear(_t(inp.reshape(1, T * R, -1)), w_ih)
            h, c, _ = ops.layer_recurrence(xproj, w_hh, bias, a, b, shared=shared, want_c=True,
                                           h0=_t(hprev[0]), c0=_t(cprev[0]), backend=backend)
            c = c.cpu().numpy().reshape(T, R, H)
            h = h.cpu().numpy().reshape(T, R, H)
            assert np.abs(c - cref).max() <= 1e-5 * max(1.0, np.abs(cref).max()), (tag, l)
            safe = np.abs(cref) > 1e-5
            assert np.array_equal(h[safe], href[safe]), (tag, l)
            inp = href


def test_cirm_gsn_vs_golden():
    g = load_golden("tiny_cirm")
    cfg = g["cfg"]
    params = synth.make_params_cirm(cfg, g["seed"])
    m = CirmGSN(**cfg)
    m.load_state_dict({k: torch.from_numpy(np.array(v)) for k, v in params.items()}, strict=True)
    m = m.eval().to(DEV)
    with torch.no_grad():
        act, all_out = m.network(_t(g["mag"]))
        enh_y, enh_mag = m(_t(g["wave"]))
    for l in range(cfg["num_layers"]):
        assert np.array_equal(all_out[1 + l].cpu().numpy(), unpack(g[f"fb_h{l}"], cfg["hidden_size"]))
    assert _rel(act.permute(1, 2, 0).cpu().numpy(), g["fb_out"]) < 1e-3
    assert _rel(enh_y.cpu().numpy(), g["enh_y"]) < 1e-3
    assert _rel(enh_mag.cpu().numpy(), g["enh_mag"]) < 1e-3


@pytest.mark.parametrize("backend", BACKENDS)
@pytest.mark.parametrize("R,K,H,L,shared,bn", [
    (1, 5, 16, 1, True, False),      # single row, tiny
    (3, 38, 160, 2, True, True),     # S sub-band shape, ragged row tile
    (130, 64, 240, 2, True, True),   # S full-band width, rows spill over one 128-row tile
    (37, 94, 224, 2, False, True),   # XL-style unshared gates
    (9, 33, 268, 3, True, True),     # cirm_gsn hidden size (not a multiple of 16)
    (70, 20, 320, 2, True, False),   # M/L full-band width
])
def test_stack_vs_oracle_shapes(R, K, H, L, shared, bn, backend):
    """StackedGSU.forward (the reference's operator API) on ragged / edge shapes vs the oracle."""
    from spiking_fullsubnet_b200 import efficient_spiking_neuron
    rs = np.random.RandomState(R * 1000 + H)
    T = 24
    p = synth._seq_model_params(rs, "m.", K, H, L, 0, shared, bn, False)
    stack = efficient_spiking_neuron(K, H, L, shared_weights=shared, bn=bn)
    stack.load_state_dict({k[len("m.sequence_model."):]: torch.from_numpy(np.array(v)) for k, v in p.items()})
    stack = stack.eval().to(DEV)
    stack.backend = backend
    x = rs.standard_normal((T, R, K)).astype(np.float32)
    with torch.no_grad():
        out, states, trace = stack(_t(x), None, want_c=True)
    ref_out, ref_trace, ref_c = O.gsn_stack_forward(x, p, "m.sequence_model.", L, shared, return_c=True)
    assert len(trace) == L + 1 and len(states) == L
    for l in range(L):
        c = stack.last_c[l].cpu().numpy()
        frac, first = spike_flip_stats(trace[1 + l].cpu().numpy(), ref_trace[1 + l])
        if frac == 0:
            assert np.abs(c - ref_c[l]).max() < 1e-4
        else:  # a flip is only legitimate where the oracle's membrane potential is at the threshold
            assert np.abs(ref_c[l][first]).min() < 1e-5, f"layer {l}: flip at frame {first} away from threshold"
            break
    assert np.array_equal(states[-1].hx.cpu().numpy(), trace[-1][-1].cpu().numpy())


def test_initial_state_and_chunked_equivalence():
    """Carrying (h,c) across two calls equals one call over the concatenated frames (streaming use)."""
    rs = np.random.RandomState(3)
    T, R, K, H = 20, 5, 12, 64
    from spiking_fullsubnet_b200 import efficient_spiking_neuron
    stack = efficient_spiking_neuron(K, H, 2, shared_weights=True, bn=False).eval().to(DEV)
    x = _t(rs.standard_normal((T, R, K)).astype(np.float32))
    with torch.no_grad():
        full, st_full, _ = stack(x, None)
        a, st_a, _ = stack(x[:9].contiguous(), None)
        b, st_b, _ = stack(x[9:].contiguous(), st_a)
    assert torch.equal(torch.cat([a, b]), full)
    assert torch.equal(st_b[1].cx, st_full[1].cx)


def test_properties_at_full_size_S():
    """BASELINE config 2 size (S, batch 32 x 4 s): size-independent properties.
    (1) batch-composition independence (SURVEY 8c determinism fact): clip 0 alone == clip 0 in the batch;
    (2) spikes are exactly {0,1}; (3) the two recurrence back ends agree on every spike."""
    cfg = synth.CFG_S
    params = synth.make_params(cfg, 5)
    m = _model(cfg, params, "auto")
    mag = _t(synth.make_mag(32, 257, 501, 11))
    with torch.no_grad():
        projs, fb_all, sb_all = m.network(mag)
        projs1, fb1, sb1 = m.network(mag[:1].contiguous())
    assert torch.equal(fb_all[1][:, :1], fb1[1])
    n0 = projs1[0].shape[1]
    assert torch.equal(projs[0][:, :n0], projs1[0])
    for t in fb_all[1:-1] + [x for al in sb_all for x in al[1:-1]]:
        assert bool(((t == 0) | (t == 1)).all())
    m.set_backend("simt")
    with torch.no_grad():
        projs_s, fb_s, sb_s = m.network(mag)
    flips = sum(float((a != b).float().sum()) for a, b in zip(fb_all[1:-1], fb_s[1:-1]))
    total = sum(a.numel() for a in fb_all[1:-1])
    assert flips / total < 1e-4, f"back ends disagree on {flips / total:.2e} of full-band spikes"


def test_cuda_graph_replay_matches_eager():
    """The captured CUDA graph of the hot path (plain and frame-chunked wavefront schedule) reproduces the
    eager launch sequence bit for bit, also after the input changes and after BatchNorm statistics change
    (the fold is part of the graph)."""
    cfg = synth.tiny_cfg()
    m = _model(cfg, synth.make_params(cfg, 9))
    mags = [_t(synth.make_mag(2, 33, 30, s)) for s in (1, 2)]
    with torch.no_grad():
        eager = [[p.clone() for p in m.network(x)[0]] for x in mags]
        m.enable_cuda_graph(True, frame_chunks=1)
        for x, ref in zip(mags, eager):
            out = m.network(x)[0]
            assert all(torch.equal(a, b) for a, b in zip(out, ref))
        m.enable_cuda_graph(True, frame_chunks=4)  # frame-chunked wavefront schedule
        for x, ref in zip(mags + mags, eager + eager):
            out = m.network(x)[0]
            assert all(torch.equal(a, b) for a, b in zip(out, ref))
        bn = m.fb_model.sequence_model.layers[0].cell.batchnorm
        bn.running_mean.add_(0.05)
        got = [p.clone() for p in m.network(mags[0])[0]]
        m.enable_cuda_graph(False)
        want = m.network(mags[0])[0]
    assert all(torch.equal(a, b) for a, b in zip(got, want))
    assert not all(torch.equal(a, b) for a, b in zip(got, eager[0]))


@pytest.mark.parametrize("backend", BACKENDS)
@pytest.mark.parametrize("name", ["tiny_surface_b", "tiny_surface_b_cumnorm", "zoo_s_1s"])
def test_surface_b_vs_golden(name, backend):
    """Surface B `Separator`: tiny fixture and the TRAINED model-zoo S checkpoint (loads strict=True).
    Trained weights are chaotic (protocol P3): spike flips are bounded by the reference's own
    noise floor (3.5e-4) and the waveform must stay within 1e-3 relative when nothing flipped."""
    from spiking_fullsubnet_b200 import Separator
    from tests.helpers import load_golden_weights
    g = load_golden(name)
    cfg = g["cfg"]
    params = load_golden_weights(name) if name.startswith("zoo") else synth.make_params_b(cfg, g["seed"])
    m = Separator(**cfg)
    m.load_state_dict({k: torch.from_numpy(np.array(v)) for k, v in params.items()}, strict=True)
    m = m.eval().to(DEV).set_backend(backend)
    with torch.no_grad():
        coefs, fb_all, sb_all = m.coefficients(_t(g["mag"]))
        enh_y, enh_mag, _, _ = m(_t(g["wave"]))
    assert _rel(fb_all[0].cpu().numpy(), g["fb_x"]) < 1e-4
    flips = total = 0
    for l in range(2):
        ref = unpack(g[f"fb_h{l}"], cfg["fb_hidden_size"])
        flips += (fb_all[1 + l].cpu().numpy() != ref).sum()
        total += ref.size
    for i in range(3):
        for l in range(2):
            ref = unpack(g[f"sb{i}_h{l}"], cfg["sb_hidden_size"])
            flips += (sb_all[i][1 + l].cpu().numpy() != ref).sum()
            total += ref.size
    print(f"{name}/{backend}: {flips} of {total} spikes differ from the reference")
    assert flips / total <= 3.5e-4
    if flips == 0:
        for i in range(3):
            assert coefs[i].shape == g[f"coef{i}"].shape
            assert _rel(coefs[i].cpu().numpy(), g[f"coef{i}"]) < 1e-3
        assert _rel(enh_y.cpu().numpy(), g["enh_y"]) < 1e-3
        assert _rel(enh_mag.cpu().numpy(), g["enh_mag"]) < 1e-3


@pytest.mark.parametrize("M,K,N", [(1, 16, 8), (130, 160, 160), (1000, 240, 64), (777, 256, 256), (300, 320, 320),
                                   (515, 268, 1542), (64, 224, 448), (5000, 160, 24)])
def test_linear_spikes_is_exact(M, K, N):
    """gsn_linear_spikes (tcgen05, weights as exact bf16x3 planes): with {0,1} inputs the result is the fp32 sum of
    exact products -- compared with a float64 product it agrees to fp32 accumulation accuracy, and with the fp32
    CUDA-core kernel to a few ulp."""
    rs = np.random.RandomState(M + K)
    a = (rs.uniform(size=(M, K)) < 0.45).astype(np.float32)
    w = rs.uniform(-0.1, 0.1, (N, K)).astype(np.float32)
    b = rs.uniform(-0.1, 0.1, N).astype(np.float32)
    ref = a.astype(np.float64) @ w.astype(np.float64).T + b
    out_tc, act_tc = ops.linear(_t(a), _t(w), _t(b), act="tanh", spikes=True)
    out_f32 = ops.linear(_t(a), _t(w), _t(b))
    scale = np.abs(ref).max()
    assert np.abs(out_tc.cpu().numpy() - ref).max() <= 2e-6 * scale
    assert np.abs(out_tc.cpu().numpy() - out_f32.cpu().numpy()).max() <= 2e-6 * scale
    assert np.abs(act_tc.cpu().numpy() - np.tanh(ref)).max() <= 1e-5


def _pack_np(h):
    """numpy restatement of the bit-packed trace layout: neuron n = bit n%32 of word n/32."""
    H = h.shape[-1]
    W = (H + 31) // 32
    padded = np.zeros(h.shape[:-1] + (W * 32,), dtype=np.uint64)
    padded[..., :H] = h != 0
    words = (padded.reshape(h.shape[:-1] + (W, 32)) << np.arange(32, dtype=np.uint64)).sum(-1)
    return words.astype(np.uint32)


@pytest.mark.parametrize("M,K,N", [(1, 16, 8), (130, 160, 160), (1000, 240, 64), (777, 256, 256), (300, 320, 320),
                                   (515, 268, 1542), (64, 224, 448), (5000, 160, 24), (333, 38, 40), (70, 250, 96)])
def test_linear_spike_bits_matches_fp32_trace(M, K, N):
    """gsn_linear_spike_bits reads the bit-packed trace (gsn_pack_spikes layout) and must give BIT-IDENTICAL
    results to gsn_linear_spikes on the fp32 trace (same MMAs on the same operands); K need not be a multiple
    of 4 or 32 here."""
    rs = np.random.RandomState(M * 7 + K)
    a = (rs.uniform(size=(M, K)) < 0.4).astype(np.float32)
    w = rs.uniform(-0.1, 0.1, (N, K)).astype(np.float32)
    b = rs.uniform(-0.1, 0.1, N).astype(np.float32)
    bits = ops.pack_spikes(_t(a))
    assert np.array_equal(bits.cpu().numpy().view(np.uint32), _pack_np(a))
    out_bits, act_bits = ops.linear(_t(a), _t(w), _t(b), act="sigmoid", spikes=True, bits=bits)
    ref = a.astype(np.float64) @ w.astype(np.float64).T + b
    assert np.abs(out_bits.cpu().numpy() - ref).max() <= 2e-6 * np.abs(ref).max()
    if K % 4 == 0:
        out_f, act_f = ops.linear(_t(a), _t(w), _t(b), act="sigmoid", spikes=True)
        assert torch.equal(out_bits, out_f) and torch.equal(act_bits, act_f)


@pytest.mark.parametrize("backend", ["simt", "tcgen05", "tcgen05_i8"])
@pytest.mark.parametrize("T,R,H,shared", [(9, 37, 160, True), (5, 16, 240, True), (7, 70, 72, False),
                                          (4, 130, 320, True), (6, 33, 100, True)])
def test_recurrence_bit_packed_trace(T, R, H, shared, backend):
    """gsn_layer_recurrence_bits: the bit-packed trace equals the fp32 trace packed on the host, for every back
    end (tcgen05 writes its ballot words, the others pack afterwards)."""
    rs = np.random.RandomState(T * 100 + R)
    gH = H if shared else 2 * H
    s = 1 / np.sqrt(H)
    xproj = rs.uniform(-1, 1, (T, R, gH)).astype(np.float32)
    w = rs.uniform(-s, s, (gH, H)).astype(np.float32)
    b = rs.uniform(-s, s, 2 * H).astype(np.float32)
    bits = ops.spike_bits_buffer((T, R), H, DEV)
    bits.fill_(-1)
    try:
        h, _, _ = ops.layer_recurrence(_t(xproj), _t(w), _t(b), shared=shared, backend=backend, out_bits=bits)
    except NotImplementedError:
        pytest.skip(f"{backend} does not support R={R} H={H} shared={shared}")
    assert np.array_equal(bits.cpu().numpy().view(np.uint32), _pack_np(h.cpu().numpy()))
    assert 0.02 < float(h.mean()) < 0.98


@pytest.mark.parametrize("name", ["tiny_train_shared_bn", "tiny_train_unshared_nobn"])
def test_training_step_vs_golden(name):
    """Protocol P4: one training step (train-mode BatchNorm, BPTT with the Triangle surrogate) against the
    reference's autograd on CPU: same spikes, loss, updated BatchNorm buffers and gradients of EVERY parameter
    (max|delta| <= 2e-3 * max|ref| per tensor: fp32 BPTT through ~20 frames in a different summation order)."""
    g = load_golden(name)
    cfg = g["cfg"]
    m = SpikingFullSubNet(**cfg)
    m.load_state_dict({k: torch.from_numpy(np.array(v)) for k, v in golden_params(g).items()}, strict=True)
    m = m.to(DEV).train()
    enh_y, enh_mag, fb_all, sb_all = m(_t(g["wave"]))
    loss = (enh_y * _t(g["target"])).sum() + enh_mag.pow(2).mean()
    loss.backward()
    for l in range(cfg["fb_num_layers"]):
        assert np.array_equal(fb_all[1 + l].detach().cpu().numpy(), unpack(g[f"fb_h{l}"], cfg["fb_hidden_size"]))
    for i in range(3):
        for l in range(cfg["sb_num_layers"]):
            assert np.array_equal(sb_all[i][1 + l].detach().cpu().numpy(),
                                  unpack(g[f"sb{i}_h{l}"], cfg["sb_hidden_size"]))
    assert abs(float(loss) - float(g["loss"])) <= 1e-3 * abs(float(g["loss"]))
    assert _rel(enh_y.detach().cpu().numpy(), g["enh_y"]) < 1e-3
    for k, b in m.named_buffers():
        ref = g["buf__" + k]
        got = b.detach().cpu().numpy()
        if ref.dtype.kind == "i":
            assert int(got) == int(ref), k
        else:
            assert np.abs(got - ref).max() <= 1e-4 * max(1.0, np.abs(ref).max()), k
    worst = 0.0
    for k, p in m.named_parameters():
        ref = g["grad__" + k]
        assert p.grad is not None, k
        err = np.abs(p.grad.cpu().numpy() - ref).max() / (np.abs(ref).max() + 1e-12)
        worst = max(worst, err)
        assert err <= 2e-3, f"{k}: gradient rel err {err:.2e}"
    print(f"{name}: worst gradient rel err {worst:.2e}")


def test_layer_backward_vs_oracle_eval_bn():
    """gsn_layer_train_backward with EVAL-mode BatchNorm against the oracle's BPTT restatement (fp64)."""
    from spiking_fullsubnet_b200.training import GSNLayerFn
    from spiking_fullsubnet_b200.modeling import GSUCell
    rs = np.random.RandomState(5)
    T, R, K, H = 12, 11, 7, 40
    for shared in (True, False):
        p = synth._seq_model_params(rs, "m.", K, H, 1, 0, shared, True, False)
        q = "m.sequence_model.layers.0.cell."
        cell = GSUCell(K, H, shared, True)
        cell.load_state_dict({k[len(q):]: torch.from_numpy(np.array(v)) for k, v in p.items()})
        cell = cell.to(DEV).eval()
        x = rs.standard_normal((T, R, K)).astype(np.float32)
        d_out = rs.standard_normal((T, R, H)).astype(np.float32)
        xt = _t(x).requires_grad_(True)
        xproj = torch.nn.functional.linear(xt, cell.weight_ih)
        h = GSNLayerFn.apply(xproj, cell.weight_hh, cell.bias_ih, cell.batchnorm.weight, cell.batchnorm.bias, cell)
        (h * _t(d_out)).sum().backward()
        _, trace, cs = O.gsn_stack_forward(x, p, "m.sequence_model.", 1, shared, return_c=True)
        assert np.array_equal(h.detach().cpu().numpy(), trace[1])
        bn = {k: p[q + "batchnorm." + k] for k in ("weight", "bias", "running_mean", "running_var")}
        ref = O.gsn_layer_backward(x, trace[1], cs[0], d_out, p[q + "weight_ih"], p[q + "weight_hh"], p[q + "bias_ih"],
                                   bn, shared)
        for got, want, nm in [(xt.grad, ref["dx"], "dx"), (cell.weight_ih.grad, ref["dw_ih"], "dw_ih"),
                              (cell.weight_hh.grad, ref["dw_hh"], "dw_hh"), (cell.bias_ih.grad, ref["dbias"], "dbias")]:
            err = np.abs(got.cpu().numpy() - want).max() / (np.abs(want).max() + 1e-12)
            assert err < 1e-4, (shared, nm, err)


def test_cuda_graph_full_forward_matches_eager():
    """forward() with STFT -> wavefront network -> deep filter replayed from a CUDA graph equals the eager call."""
    g = load_golden("tiny_shared_bn")
    m = _model(g["cfg"], golden_params(g))
    waves = [_t(g["wave"]), _t(g["wave"][:, ::-1].copy())]
    with torch.no_grad():
        eager = [[t.clone() for t in m(w)[:2]] for w in waves]
        m.enable_cuda_graph(True, frame_chunks=4)
        for w, ref in zip(waves + waves, eager + eager):
            out = m(w)
            assert torch.equal(out[0], ref[0]) and torch.equal(out[1], ref[1])
    assert _rel(eager[0][0].cpu().numpy(), g["enh_y"]) < 1e-3


def test_separator_and_cirm_graph_and_autograd_paths():
    """Surface B and cirm_gsn: CUDA-graph replay equals the eager launches; the autograd path (gradients enabled,
    eval-mode BatchNorm) reproduces the inference path's spikes / waveform and gives every weight a finite gradient."""
    from spiking_fullsubnet_b200 import Separator
    g = load_golden("tiny_surface_b")
    sep = Separator(**g["cfg"])
    sep.load_state_dict({k: torch.from_numpy(np.array(v)) for k, v in synth.make_params_b(g["cfg"], g["seed"]).items()})
    gc = load_golden("tiny_cirm")
    cirm = CirmGSN(**gc["cfg"])
    cirm.load_state_dict({k: torch.from_numpy(np.array(v))
                          for k, v in synth.make_params_cirm(gc["cfg"], gc["seed"]).items()})
    for m, gg in ((sep, g), (cirm, gc)):
        m = m.eval().to(DEV)
        wave, mag = _t(gg["wave"]), _t(gg["mag"])
        with torch.no_grad():
            eager = m.network(mag)
            ref_y = m(wave)[0].clone()
            m.enable_cuda_graph(True)
            graphed = m.network(mag)
            a = eager[0] if isinstance(eager[0], list) else [eager[0]]
            b = graphed[0] if isinstance(graphed[0], list) else [graphed[0]]
            assert all(torch.equal(x, y) for x, y in zip(a, b))
            m.enable_cuda_graph(False)
        out = m(wave)  # gradients enabled -> autograd path
        assert _rel(out[0].detach().cpu().numpy(), ref_y.cpu().numpy()) < 1e-4
        (out[0].square().mean() + out[1].mean()).backward()
        for k, p in m.named_parameters():
            if "batchnorm" in k:
                continue  # eval-mode BatchNorm affine is frozen on this path
            assert p.grad is not None and torch.isfinite(p.grad).all() and float(p.grad.abs().max()) > 0, k


@pytest.mark.parametrize("R,H,bn,training", [(300, 160, True, True), (37, 240, True, True), (64, 320, True, False),
                                             (130, 224, False, False)])
def test_train_forward_tcgen05_vs_simt(R, H, bn, training):
    """The tcgen05 training forward (batch-statistics BatchNorm through a grid barrier) against the fp32 CUDA-core
    training kernel: identical spikes, saved tensors and running statistics within fp32 summation noise."""
    rs = np.random.RandomState(R + H)
    T = 40
    s = 1 / np.sqrt(H)
    xproj = _t(rs.uniform(-1, 1, (T, R, H)).astype(np.float32))
    w = _t(rs.uniform(-s, s, (H, H)).astype(np.float32))
    b = _t(rs.uniform(-s, s, 2 * H).astype(np.float32))
    outs = []
    for tc_on in (True, False):
        ops.TC_TRAIN[0] = tc_on
        g_ = _t(rs.uniform(0.6, 1.0, H).astype(np.float32)) if bn else None
        rs2 = np.random.RandomState(1)
        gam = _t(rs2.uniform(0.6, 1.0, H).astype(np.float32)) if bn else None
        bet = _t(rs2.normal(0, 0.1, H).astype(np.float32)) if bn else None
        rm = _t(rs2.normal(0, 0.1, H).astype(np.float32)) if bn else None
        rv = _t(rs2.uniform(1.0, 2.0, H).astype(np.float32)) if bn else None
        res = ops.layer_train_forward(xproj, w, b, gam, bet, rm, rv, training, 0.1, 1e-5, True)
        outs.append((res, rm, rv))
    ops.TC_TRAIN[0] = True
    (a, rma, rva), (bb, rmb, rvb) = outs
    assert torch.equal(a[0], bb[0]), "spikes differ between the tcgen05 and the CUDA-core training forward"
    for x, y in zip(a[1:], bb[1:]):
        if x is not None:
            assert float((x - y).abs().max()) <= 2e-5 * max(1.0, float(y.abs().max()))
    if bn:
        assert float((rma - rmb).abs().max()) < 1e-5 and float((rva - rvb).abs().max()) < 1e-5


def test_layer_training_large_rows_vs_torch_autograd():
    """Rows > 256 exercise the 16-row backward CTAs and multi-tile batch statistics: one GSU layer, train-mode
    BatchNorm, against a float64 torch-autograd restatement (Triangle surrogate) on the same device."""
    from spiking_fullsubnet_b200.modeling import GSUCell
    from spiking_fullsubnet_b200.training import GSNLayerFn

    class Tri(torch.autograd.Function):
        @staticmethod
        def forward(ctx, c):
            ctx.save_for_backward(c)
            return (c >= 0).to(c.dtype)

        @staticmethod
        def backward(ctx, g):
            (c,) = ctx.saved_tensors
            return g * (1 - c.abs()).clamp(min=0)

    torch.manual_seed(0)
    T, R, K, H = 7, 1100, 12, 64
    cell = GSUCell(K, H, True, True).to(DEV).train()
    with torch.no_grad():
        cell.batchnorm.weight.uniform_(0.6, 1.0)
        cell.batchnorm.bias.normal_(0, 0.1)
    x = torch.randn(T, R, K, device=DEV)
    d_out = torch.randn(T, R, H, device=DEV)
    # reference in float64
    p64 = {k: v.detach().double().requires_grad_(True) for k, v in cell.named_parameters()}
    h = torch.zeros(R, H, device=DEV, dtype=torch.float64)
    c = torch.zeros_like(h)
    hs = []
    for t in range(T):
        z = x[t].double() @ p64["weight_ih"].t() + h @ p64["weight_hh"].t()
        f = torch.sigmoid(z + p64["bias_ih"][:H])
        g = z + p64["bias_ih"][H:]
        c = f * c + (1 - f) * g
        c = torch.nn.functional.batch_norm(c, None, None, p64["batchnorm.weight"], p64["batchnorm.bias"], True, 0.1, 1e-5)
        h = Tri.apply(c)
        hs.append(h)
    href = torch.stack(hs)
    (href * d_out.double()).sum().backward()
    # product path
    xproj = torch.nn.functional.linear(x, cell.weight_ih)
    hp = GSNLayerFn.apply(xproj, cell.weight_hh, cell.bias_ih, cell.batchnorm.weight, cell.batchnorm.bias, cell)
    (hp * d_out).sum().backward()
    assert torch.equal(hp.detach().double(), href.detach()), "spikes differ from the float64 reference"
    for k, p in cell.named_parameters():
        ref = p64[k].grad
        err = float((p.grad.double() - ref).abs().max() / (ref.abs().max() + 1e-12))
        assert err < 1e-3, (k, err)


# ---- BASELINE-sized parity (VERDICT r01 "close the parity holes") ---------------------------------------------------
from tests.helpers import LONG, block_forced_check, compare_long, load_long, record_parity  # noqa: E402


def _long_model(g, backend="auto"):
    from spiking_fullsubnet_b200 import Separator
    cls = SpikingFullSubNet if g["surface"] == "A" else Separator
    m = cls(**g["cfg"])
    m.load_state_dict({k: torch.from_numpy(np.array(v)) for k, v in g["params"].items()}, strict=True)
    return m.eval().to(DEV).set_backend(backend)


@pytest.mark.parametrize("name", LONG)
def test_long_free_running_vs_reference(name):
    """Free-running parity with the REFERENCE at BASELINE sizes: T = 501 (config 2) and T = 1 251 (config 3) frames,
    surface-A S (the bench's weights), trained zoo-S and zoo-L checkpoints.  fp32 threshold chaos makes the raw flip
    count a heavy-tailed quantity (tests/helpers.py, divergence audit), so the pass/fail statements are:
    (1) every row trajectory leaves the reference -- if at all -- only through neurons whose reference membrane
    potential is within 1e-5 of the threshold, or within 8 x the drift between an fp32 and an fp64 evaluation of that
    neuron under the reference's own spike history (tests.helpers.membrane_noise: the trained zoo-L recursion is
    expanding, a f > 1, and that drift reaches 1e-2 by frame 431 with identical spikes); (2) until then the coefficients agree to 1e-4 of max|ref| (north star:
    1e-3).  Raw counts, the reference's own 1 +- 2e-6 noise floor and the audit go to gpurun_out/parity_counts.json."""
    from tests.helpers import coef_rel_before_divergence, divergence_audit, membrane_noise, reference_membrane
    g = load_long(name)
    m = _long_model(g)
    with torch.no_grad():
        coefs, fb_all, sb_all = m.coefficients(_t(g["mag"]))
    st = compare_long(g, coefs, fb_all, sb_all)
    c_hat, _ = reference_membrane(g)
    audit, div = divergence_audit(g, c_hat, fb_all, sb_all, noise=membrane_noise(g))
    audit["coef_rel_before_divergence"] = coef_rel_before_divergence(g, coefs, div)
    st.update(audit)
    print(name, st)
    record_parity(f"free_running/{name}", st)
    assert st["bad_root_flips"] == 0, f"{name}: a trajectory left the reference away from the threshold: {st}"
    assert st["coef_rel_before_divergence"] < 1e-4, st
    if st["flips"] == 0:
        assert st["coef_rel"] < 1e-4


def _gpu_run_layer(backend, nt_want, tiles_seen):
    def run_layer(inp, w_ih, w_hh, bias, bn, shared, h0, c0):
        a = b = None
        if bn is not None:
            inv = 1.0 / np.sqrt(bn["running_var"] + np.float32(1e-5))
            al = (inv * bn["weight"]).astype(np.float32)
            a, b = _t(al), _t((bn["bias"] - bn["running_mean"] * al).astype(np.float32))
        x = _t(inp)
        spikes = bool(((inp == 0) | (inp == 1)).all())
        xproj = ops.linear(x, _t(w_ih), spikes=spikes)
        _, R, _ = inp.shape
        H = w_hh.shape[1]
        # SM budget that makes the tile picker choose `nt_want` rows per cluster (or the largest tile that fits TMEM)
        C = (H + 127) // 128 if shared else (H + 63) // 64
        budget = 0
        if nt_want > 16:
            budget = max(1, ((R + nt_want // 2 - 1) // (nt_want // 2)) * C - 1)
        tiles_seen.add(ops.recurrence_tile(R, H, shared, backend, budget))
        h, _, _ = ops.layer_recurrence(xproj, _t(w_hh), _t(bias), a, b, shared=shared, h0=_t(h0), c0=_t(c0),
                                       backend=backend, sm_budget=budget)
        return h.cpu().numpy()
    return run_layer


@pytest.mark.parametrize("nt", [16, 32, 64])
@pytest.mark.parametrize("backend", ["tcgen05", "tcgen05_i8"])
@pytest.mark.parametrize("name", ["zoo_s_2x4s", "zoo_l_2x4s", "zoo_l_1x10s"])
def test_block_teacher_forced_tiles_vs_reference(name, backend, nt):
    """Every row tile (NT = 16 / 32 / 64) of both tcgen05 back ends against the REFERENCE with TRAINED weights at
    H = 160 / 240 / 256 / 320 over T = 501 / 1 251 frames: the recurrence restarts from the reference's state every
    32 frames (tests.helpers.block_forced_check), all blocks in one launch (rows' = blocks x rows: up to 1 872
    rows), layer by layer with the reference's own layer inputs.  Bound: < 1e-4 of all spikes (the numpy oracle under
    the same protocol: 8e-6 on zoo-L)."""
    g = load_long(name)
    m = _long_model(g)
    with torch.no_grad():
        _, fb_all, sb_all = m.coefficients(_t(g["mag"]))
    xs = {"fb": fb_all[0].cpu().numpy()}
    xs.update({f"sb{i}": al[0].cpu().numpy() for i, al in enumerate(sb_all)})
    tiles = set()
    try:
        st = block_forced_check(g, xs, _gpu_run_layer(backend, nt, tiles))
    except NotImplementedError:
        pytest.skip(f"{backend} does not support a shape of {name}")
    st["tiles_used"] = sorted(tiles)
    print(name, backend, nt, st)
    record_parity(f"block_forced/{name}/{backend}/nt{nt}", st)
    assert max(tiles) == nt or nt == 64, f"tile picker never chose NT={nt}: {tiles}"
    assert st["flips"] / st["total"] < 1e-4, st


@pytest.mark.parametrize("R,H,shared,backend", [(1536, 256, True, "tcgen05"), (4096, 256, True, "tcgen05"),
                                                (4096, 320, True, "tcgen05"), (1536, 224, False, "tcgen05"),
                                                (4096, 256, True, "tcgen05_i8"), (2048, 448, True, "tcgen05_i8")])
def test_large_row_tiles_vs_oracle(R, H, shared, backend):
    """The row counts of BASELINE config 3 (L at batch 64: R = 1 024 / 1 536; 4 096 for the config-5 sweep): with the
    whole device as budget the picker leaves NT = 16 (R = 1 536 -> 32, R = 4 096 -> 64 at H = 256); two-layer stack
    against the numpy oracle, every spike."""
    from spiking_fullsubnet_b200 import efficient_spiking_neuron
    rs = np.random.RandomState(R + H)
    T, K, L = 12, 38, 2
    p = synth._seq_model_params(rs, "m.", K, H, L, 0, shared, True, False)
    stack = efficient_spiking_neuron(K, H, L, shared_weights=shared, bn=True)
    stack.load_state_dict({k[len("m.sequence_model."):]: torch.from_numpy(np.array(v)) for k, v in p.items()})
    stack = stack.eval().to(DEV)
    stack.backend = backend
    nt = ops.recurrence_tile(R, H, shared, backend, 0)
    x = rs.standard_normal((T, R, K)).astype(np.float32)
    try:
        with torch.no_grad():
            out, states, trace = stack(_t(x), None, want_c=True)
    except NotImplementedError:
        pytest.skip(f"{backend} does not support R={R} H={H}")
    _, ref_trace, ref_c = O.gsn_stack_forward(x, p, "m.sequence_model.", L, shared, return_c=True)
    worst = 0.0
    for l in range(L):
        got = trace[1 + l].cpu().numpy()
        diff = got != ref_trace[1 + l]
        if diff.any():  # only legitimate where the oracle's membrane potential sits on the threshold
            first = int(np.argmax(diff.reshape(T, -1).any(axis=1)))
            bad = diff[first]
            assert np.abs(ref_c[l][first][bad]).max() < 1e-5, f"NT={nt} layer {l}: flip away from the threshold"
            break
        worst = max(worst, float(np.abs(stack.last_c[l].cpu().numpy() - ref_c[l]).max()))
    record_parity(f"large_rows/R{R}_H{H}_{'sh' if shared else 'un'}_{backend}", {"nt": nt, "max_c_err": worst})
    if H <= 320:  # (kind::i8 at H = 448 only fits the NT = 16 tile in tensor memory)
        assert nt > 16, f"expected a coarse row tile for R={R}, got NT={nt}"
    assert worst < 1e-4


def test_wavefront_graph_matches_eager_at_full_size_S():
    """The schedule bench.py times (CUDA graph, 12-chunk frame wavefront, programmatic dependent launches) against the
    eager launch sequence at BASELINE config-2 size (S, batch 32 x 4 s, T = 501): bit-identical coefficients."""
    cfg = synth.CFG_S
    m = _model(cfg, synth.make_params(cfg, 5))
    mag = _t(synth.make_mag(32, 257, 501, 11))
    with torch.no_grad():
        eager = [p.clone() for p in m.network(mag)[0]]
        m.enable_cuda_graph(True, frame_chunks=12)
        for _ in range(2):
            out = m.network(mag)[0]
            assert all(torch.equal(a, b) for a, b in zip(out, eager))


def test_exported_submodules_are_autograd_safe():
    """The drop-in sub-modules (SequenceModel / SubbandModel / GSULayer / GSUCell) route to the autograd kernels when
    gradients are recorded: outputs carry a grad_fn, EVERY parameter (pre-LayerNorm and proj included) receives a
    gradient, eval-mode results equal the inference kernels', and a non-zero initial state is refused instead of being
    silently ignored."""
    from spiking_fullsubnet_b200.modeling import GSUCell, GSULayer, SequenceModel, SubbandModel
    torch.manual_seed(0)
    sm = SequenceModel(input_size=12, hidden_size=40, num_layers=2, sequence_model="GSN", proj_size=6,
                       shared_weights=True, output_activate_function="tanh", bn=True, use_pre_layer_norm=True).to(DEV)
    x = torch.randn(5, 12, 30, device=DEV)
    sm.train()
    out, all_out = sm(x)
    assert out.grad_fn is not None and out.shape == (5, 6, 30) and len(all_out) == 4
    out.square().mean().backward()
    missing = [n for n, p in sm.named_parameters() if p.grad is None or not torch.isfinite(p.grad).all()]
    assert not missing, f"no gradient for {missing}"
    sm.eval()
    with torch.no_grad():
        ref, _ = sm(x)
    out_eval, _ = sm(x)  # eval BatchNorm, gradients on: autograd kernels with running statistics
    assert out_eval.grad_fn is not None
    # (F.linear / cuBLAS vs the fp32 FFMA kernel: a near-threshold spike may flip and move a few outputs)
    assert float(((out_eval - ref).abs() > 1e-4).float().mean()) < 0.02

    sb = SubbandModel(freq_cutoffs=[0, 8, 24, 32], center_freq_sizes=[2, 4, 8], neighbor_freq_sizes=[3, 3, 3],
                      df_orders=[3, 2, 1], num_spks=1, hidden_size=40, num_layers=2, shared_weights=True,
                      sequence_model="GSN", bn=True, use_pre_layer_norm=True).to(DEV).train()
    noisy = torch.rand(2, 1, 32, 20, device=DEV)
    fbo = torch.randn(2, 1, 32, 20, device=DEV)
    coefs, traces = sb(noisy, fbo)
    assert all(c.grad_fn is not None for c in coefs) and len(traces) == 3
    sum(c.square().mean() for c in coefs).backward()
    missing = [n for n, p in sb.named_parameters() if p.grad is None]
    assert not missing, f"no gradient for {missing}"

    layer = GSULayer(GSUCell, 12, 40, True, True).to(DEV).train()
    xin = torch.randn(7, 5, 12, device=DEV)
    h, st = layer(xin, None)
    assert h.grad_fn is not None and h.shape == (7, 5, 40)
    with pytest.raises(NotImplementedError):
        layer(xin, (torch.ones(5, 40, device=DEV), torch.zeros(5, 40, device=DEV)))
    h1, _ = layer.cell(xin[0], (torch.zeros(5, 40, device=DEV), torch.zeros(5, 40, device=DEV)))
    assert h1.grad_fn is not None and h1.shape == (5, 40)


@pytest.mark.parametrize("B,F,T,fdrc", [(3, 257, 101, 0.5), (2, 33, 34, 0.5), (1, 257, 40, 0.3), (2, 129, 77, 1.0)])
def test_spectral_front_end_on_the_complex_stft(B, F, T, fdrc):
    """gsn_compress_spec: |X|^fdrc straight from the complex STFT equals torch.abs(X)**fdrc in the network's
    time-major layout (MSF:434-436, 108), and equals gsn_compress_mag on the materialised magnitude to 1 ulp."""
    g = torch.Generator(device="cpu").manual_seed(B + T)
    spec = torch.complex(torch.randn(B, F, T, generator=g), torch.randn(B, F, T, generator=g)).to(DEV)
    cm = ops.compress_mag(spec, F - 1, fdrc)
    ref = (spec.abs() ** fdrc)[:, :-1, :].permute(2, 0, 1)
    assert cm.shape == (T, B, F - 1)
    assert float((cm - ref).abs().max()) <= 2e-6 * float(ref.abs().max())
    cm2 = ops.compress_mag(spec.abs().contiguous(), F - 1, fdrc)
    assert float((cm - cm2).abs().max()) <= 2e-7 * float(ref.abs().max())
    spec_tm = spec.transpose(1, 2).contiguous().transpose(1, 2)  # time-major view, as torch.stft returns it
    assert not spec_tm.is_contiguous() and torch.equal(ops.compress_mag(spec_tm, F - 1, fdrc), cm)


@pytest.mark.parametrize("layout", [0, 1])
@pytest.mark.parametrize("B,N,ctr,df,S,lo,F", [(2, 8, 4, 3, 1, 0, 257), (3, 3, 32, 1, 2, 32, 257), (2, 1, 33, 3, 2, 0, 33)])
def test_deepfilter_spec_matches_the_reference_formula(B, N, ctr, df, S, lo, F, layout):
    """gsn_deepfilter_spec (complex in / complex out) against the reference's deep filter evaluated with torch in
    float64 (MSF:315-346: pad df-1 frames on the left, sum_d spec[t - (df-1) + d] * coef[d]) for both proj feature
    orders, plus the Nyquist pass-through."""
    g = torch.Generator(device="cpu").manual_seed(ctr + df)
    T = 29
    spec = torch.complex(torch.randn(B, F, T, generator=g), torch.randn(B, F, T, generator=g)).to(DEV)
    P = 2 * ctr * df * S
    proj = torch.randn(T, B * N, P, generator=g).to(DEV)
    out = torch.zeros(B, S, F, T, dtype=torch.complex64, device=DEV)
    mag = torch.full(out.shape, -1.0, device=DEV)
    ops.deepfilter_spec(proj, spec, out, N, ctr, df, S, lo, layout=layout, mag=mag)
    ops.spec_passthrough(spec, out, lo + N * ctr, mag=mag)
    written = out[:, :, lo:]  # |.| of everything the two kernels wrote, as torch.abs computes it (enh_mag, MSF:472)
    assert float((mag[:, :, lo:] - written.abs()).abs().max()) <= 2e-7 * float(written.abs().max())
    v = proj.double().reshape(T, B, N, *((2, ctr, df, S) if layout == 0 else (2, df, S, ctr)))
    v = v.permute(1, 2, 3, 4, 5, 6, 0) if layout == 0 else v.permute(1, 2, 3, 6, 4, 5, 0)   # [B,N,c,fc,df,S,T]
    coef = torch.complex(v[:, :, 0], v[:, :, 1])                                           # [B,N,fc,df,S,T]
    band = spec[:, lo:lo + N * ctr].to(torch.complex128).reshape(B, N, ctr, T)
    pad = torch.nn.functional.pad(band, (df - 1, 0))
    ref = sum(pad[:, :, :, None, d:d + T] * coef[:, :, :, d] for d in range(df))           # [B,N,fc,S,T]
    ref = ref.permute(0, 3, 1, 2, 4).reshape(B, S, N * ctr, T)
    got = out[:, :, lo:lo + N * ctr].to(torch.complex128)
    assert float((got - ref).abs().max()) <= 1e-5 * float(ref.abs().max())
    assert torch.equal(out[:, :, lo + N * ctr:], spec[:, None, lo + N * ctr:].expand(-1, S, -1, -1))
    # time-major spectra (transposed views of [B,T,F] / [B,S,T,F]): same values
    spec_tm = spec.transpose(1, 2).contiguous().transpose(1, 2)
    out_tm = torch.zeros(B, S, T, F, dtype=torch.complex64, device=DEV).transpose(2, 3)
    ops.deepfilter_spec(proj, spec_tm, out_tm, N, ctr, df, S, lo, layout=layout)
    ops.spec_passthrough(spec_tm, out_tm, lo + N * ctr)
    assert torch.equal(out_tm[:, :, lo:], out[:, :, lo:])


@pytest.mark.parametrize("B,n_fft,hop,L", [(3, 512, 128, 16000), (2, 64, 16, 528), (1, 512, 128, 64000), (2, 512, 128, 1000),
                                           (2, 512, 128, 1001), (2, 64, 16, 531)])
def test_fused_istft_matches_torch_istft(B, n_fft, hop, L):
    """cuFFT inverse real FFT + gsn_overlap_add against torch.istft (audio_feature.py:297-347)."""
    from spiking_fullsubnet_b200.modeling import _istft_fused
    g = torch.Generator(device="cpu").manual_seed(L)
    wave = torch.randn(B, L, generator=g).to(DEV)
    window = torch.hann_window(n_fft, device=DEV)
    spec = torch.stft(wave, n_fft, hop, n_fft, window=window, return_complex=True, pad_mode="constant")
    spec = spec * torch.complex(torch.rand_like(spec.real) + 0.5, torch.rand_like(spec.real) - 0.5)  # not a valid STFT
    ref = torch.istft(spec, n_fft, hop, n_fft, window=window, length=L)
    got = _istft_fused(spec.contiguous(), n_fft, hop, n_fft, L)
    assert got.shape == ref.shape
    assert float((got - ref).abs().max()) <= 2e-5 * float(ref.abs().max())
    assert not spec.is_contiguous()  # torch.stft hands out the time-major view: no transpose copy on that path
    got_tm = _istft_fused(spec, n_fft, hop, n_fft, L)  # n_fft = 512: gsn_irfft_frames instead of cuFFT
    assert float((got_tm - ref).abs().max()) <= 2e-5 * float(ref.abs().max())
    if n_fft != ops.FFT_FUSED_N:
        assert torch.equal(got_tm, got)


@pytest.mark.parametrize("B,n_fft,hop,L", [(3, 512, 128, 16000), (2, 64, 16, 801), (1, 512, 128, 64000), (2, 128, 32, 1000),
                                           (2, 512, 128, 300)])
def test_fused_stft_matches_torch_stft(B, n_fft, hop, L):
    """Row f2, analysis side: gsn_frame_signal (zero padding + framing + window in one pass) + cuFFT's batched real FFT
    against torch.stft(center=True, pad_mode="constant", hann window) (audio_feature.py:236-294): same frame count, same
    [B,F,T] time-major view, values within fp32 FFT rounding."""
    from spiking_fullsubnet_b200.modeling import _stft, _stft_fused
    rs = np.random.RandomState(L + n_fft)
    y = _t(rs.standard_normal((B, L)).astype(np.float32))
    want = _stft(y, n_fft, hop, n_fft)
    launches = ops.LAUNCHES[0]
    got = _stft_fused(y, n_fft, hop, n_fft)
    assert ops.LAUNCHES[0] == launches + 1, "the framing kernel did not run"
    assert got.shape == want.shape and got.stride() == want.stride()
    err = (got - want).abs().max().item()
    assert err <= 2e-6 * want.abs().max().item(), err
    frames = ops.frame_signal(y, torch.hann_window(n_fft, device=DEV), hop)
    pad = torch.nn.functional.pad(y, (n_fft // 2, n_fft // 2))
    ref = pad.unfold(1, n_fft, hop) * torch.hann_window(n_fft, device=DEV)
    assert torch.equal(frames, ref)


@pytest.mark.parametrize("B,hop,L,f_keep,fdrc", [(3, 128, 16000, 256, 0.5), (2, 128, 300, 257, 1.0), (1, 256, 64000, 256, 0.3),
                                                 (2, 100, 1001, 256, 0.5), (5, 128, 40037, 64, 0.5)])
def test_stft_compress_kernel_matches_torch_stft_and_compress(B, hop, L, f_keep, fdrc):
    """Row f2: gsn_stft_compress (zero padding + framing + window + 512-point real FFT + |X|**fdrc, one kernel) against
    torch.stft(center=True, pad_mode="constant", hann window) (audio_feature.py:236-294) and mag**fdrc in the network's
    layout (MSF:434-436, MSF:108); ragged lengths, frame counts that are not a multiple of the block's four frames."""
    rs = np.random.RandomState(L + hop)
    y = _t(rs.standard_normal((B, L)).astype(np.float32))
    window = torch.hann_window(512, device=DEV)
    want = torch.stft(y, 512, hop, 512, window=window, return_complex=True, pad_mode="constant")
    spec, cm = ops.stft_compress(y, window, hop, f_keep, fdrc)
    assert spec.shape == want.shape and spec.stride() == want.stride() and cm.shape == (want.shape[2], B, f_keep)
    err = (spec - want).abs().max().item()
    assert err <= 2e-6 * want.abs().max().item(), err
    # the compressed magnitude is that of the spectrum the kernel wrote (to the last ulp or two: |.| without hypotf's
    # scaling on the common path), and within FFT rounding of the reference's
    own = ops.compress_mag(spec, f_keep, fdrc)
    assert float(((cm - own).abs() / own.clamp_min(1e-20)).max()) <= 4e-7
    ref_cm = (want.abs()[:, :f_keep] ** fdrc).permute(2, 0, 1)
    assert float((cm - ref_cm).abs().max()) <= 1e-4 * float(ref_cm.abs().max())
    spec2, none = ops.stft_compress(y, window, hop)
    assert none is None and torch.equal(spec2, spec)


@pytest.mark.parametrize("B,T", [(2, 126), (1, 3), (3, 501)])
def test_irfft_frames_kernel_matches_torch_irfft(B, T):
    """gsn_irfft_frames against torch.fft.irfft(n=512) of every frame, including the C2R convention that the imaginary
    parts of the DC and Nyquist bins are ignored."""
    g = torch.Generator(device="cpu").manual_seed(B * 1000 + T)
    spec_tm = torch.complex(torch.randn(B, T, 257, generator=g), torch.randn(B, T, 257, generator=g)).to(DEV)
    want = torch.fft.irfft(spec_tm, n=512, dim=-1)
    got = ops.irfft_frames(spec_tm.transpose(1, 2))
    assert got.shape == want.shape
    assert float((got - want).abs().max()) <= 2e-6 * float(want.abs().max())
    with pytest.raises(ValueError):
        ops.irfft_frames(spec_tm.transpose(1, 2).contiguous())


@pytest.mark.parametrize("layout,bands", [(0, [(4, 8, 5), (6, 16, 3), (2, 64, 1)]), (0, [(8, 32, 2)]), (1, [(1, 257, 3)]),
                                          (0, [(2, 16, 3), (2, 16, 1), (2, 32, 2), (4, 32, 1)])])
def test_deepfilter_irfft_kernel_matches_the_band_kernels_and_irfft(layout, bands):
    """gsn_deepfilter_irfft (deep filter of all bands + pass-through + inverse FFT, the enhanced spectrum never in
    memory) against gsn_deepfilter_spec / gsn_spec_passthrough + torch.fft.irfft (MSF:315-346, 449-472; CGN:230)."""
    B, T, F = 3, 37, 257
    g = torch.Generator(device="cpu").manual_seed(len(bands) * 7 + layout)
    spec_tm = torch.complex(torch.randn(B, T, F, generator=g), torch.randn(B, T, F, generator=g)).to(DEV)
    spec = spec_tm.transpose(1, 2)
    projs = [torch.randn(T, B * n, 2 * c * d, generator=g).to(DEV) for n, c, d in bands]
    Ns, ctrs, dfs = ([b[i] for b in bands] for i in range(3))
    enh = torch.empty((B, 1, T, F), dtype=torch.complex64, device=DEV).transpose(2, 3)
    mag = torch.empty((B, 1, T, F), dtype=torch.float32, device=DEV).transpose(2, 3)
    lo = 0
    for p, n, c, d in zip(projs, Ns, ctrs, dfs):
        ops.deepfilter_spec(p, spec, enh, n, c, d, 1, lo, layout=layout, mag=mag)
        lo += n * c
    ops.spec_passthrough(spec, enh, lo, mag=mag)
    frames, got_mag, got_enh = ops.deepfilter_irfft(projs, spec, Ns, ctrs, dfs, layout=layout, want_enh=True)
    assert float((got_enh - enh).abs().max()) <= 1e-6 * float(enh.abs().max())
    assert float((got_mag - mag).abs().max()) <= 1e-6 * float(mag.abs().max())
    want = torch.fft.irfft(enh[:, 0].transpose(1, 2), n=512, dim=-1)
    assert float((frames - want).abs().max()) <= 2e-6 * float(want.abs().max())
    frames2, m2, e2 = ops.deepfilter_irfft(projs, spec, Ns, ctrs, dfs, layout=layout, want_mag=False)
    assert m2 is None and e2 is None and torch.equal(frames2, frames)


def test_forward_with_the_fused_fft_kernels_equals_its_parts():
    """forward() at the recipes' n_fft = 512 (S, surface A): the fused front end hands the network the compressed
    magnitude of its own spectrum, and the result is the fused back end applied to the network's coefficients; against
    the per-band kernels + cuFFT on the SAME spectrum and coefficients the waveform agrees to FFT rounding."""
    from spiking_fullsubnet_b200 import modeling
    cfg = synth.CONFIGS["S"]
    m = _model(cfg, synth.make_params(cfg, 5))
    wave = _t(synth.make_wave(2, 16000, 3))
    with torch.no_grad():
        n0 = ops.LAUNCHES[0]
        y, mag, fb_all, sb_all = m(wave)
        fused_launches = ops.LAUNCHES[0] - n0
        cmp = modeling._stft_fused(wave, 512, cfg["hop_length"], 512, f_keep=256, fdrc=cfg["fdrc"])
        own = ops.compress_mag(cmp.transpose(1, 2).contiguous().transpose(1, 2), 256, cfg["fdrc"])
        assert float(((cmp._gsn_cm[2] - own).abs() / own.clamp_min(1e-20)).max()) <= 4e-7
        projs, _, _ = m.network(cmp)
        # the reference composition on the same spectrum: per-band deep filter, pass-through, cuFFT, overlap-add
        enh = modeling._empty_spec_like(cmp, 1)
        ref_mag = torch.empty_strided(enh.shape, enh.stride(), dtype=torch.float32, device=DEV)
        cuts, ctrs = m.sb_model.freq_cutoffs, m.sb_model.center_freq_sizes
        lo = 0
        for i, p in enumerate(projs):
            n = (cuts[i + 1] - cuts[i]) // ctrs[i]
            ops.deepfilter_spec(p, cmp, enh, n, ctrs[i], m.df_orders[i], 1, lo, mag=ref_mag)
            lo += n * ctrs[i]
        ops.spec_passthrough(cmp, enh, lo, mag=ref_mag)
        window = torch.hann_window(512, device=DEV)
        ref_y = torch.istft(enh[:, 0], 512, cfg["hop_length"], 512, window=window, length=wave.shape[1])
    assert float((y - ref_y).abs().max()) <= 2e-5 * float(ref_y.abs().max())
    assert float((mag - ref_mag[:, 0]).abs().max()) <= 1e-6 * float(ref_mag.abs().max())
    os_env = __import__("os").environ
    os_env["GSN_FFT_FUSED"] = "0"
    try:
        with torch.no_grad():
            n0 = ops.LAUNCHES[0]
            y0 = m(wave)[0]
            assert ops.LAUNCHES[0] - n0 > fused_launches  # the cuFFT path launches more of the library's kernels
    finally:
        del os_env["GSN_FFT_FUSED"]
    assert y0.shape == y.shape and torch.isfinite(y0).all()


@pytest.mark.parametrize("streaming", [False, True])
def test_graph_replay_compresses_the_callers_tensor(streaming):
    """network() in graph mode keeps the input compression outside the captured graph (it runs on the caller's tensor in
    front of every replay, no copy into a static input): replays on DIFFERENT tensors, real magnitudes and complex
    spectra of either layout, must equal the eager results on those tensors."""
    cfg = synth.CONFIGS["S"]
    m = _model(cfg, synth.make_params(cfg, 5))
    if streaming:
        m.enable_streaming(True)
    window = torch.hann_window(512, device=DEV)
    specs = [torch.stft(_t(synth.make_wave(2, 8000, seed)), 512, 128, 512, window=window, return_complex=True,
                        pad_mode="constant") for seed in (3, 4, 5)]
    inputs = [specs[0].abs().contiguous(), specs[1].abs().contiguous(), specs[2], specs[1].contiguous(), specs[0]]
    with torch.no_grad():
        m.enable_cuda_graph(False)
        want = [[p.clone() for p in m.network(x)[0]] for x in inputs]
        m.enable_cuda_graph(True, frame_chunks=4)
        for _ in range(2):
            for x, w in zip(inputs, want):
                got = m.network(x)[0]
                for a, b in zip(got, w):
                    assert torch.equal(a, b)
    assert m.graph_launches > 0


def test_loss_terms_on_the_gpu_match_reference_values_and_gradient():
    """Row f3 on the device the training step runs on: freq_MAE / mag_MAE / SISNRLoss and the recipe's combined loss
    (audiozen/loss.py:138-190, 11-40; recipes/.../trainer.py:33-37) against values and the waveform gradient the
    reference produced (tests/golden/loss_ref.npz); cuFFT vs the CPU FFT: 1e-5 relative."""
    import os
    from spiking_fullsubnet_b200 import losses
    from tests.helpers import loss_waveforms
    G = np.load(os.path.join(os.path.dirname(__file__), "golden", "loss_ref.npz"))
    est, clean = (_t(a) for a in loss_waveforms())

    def close(a, b, rel=1e-5):
        return abs(float(a) - float(b)) <= rel * max(1.0, abs(float(b)))

    assert close(losses.freq_MAE(est, clean), G["loss_freq_mae"])
    assert close(losses.mag_MAE(est, clean), G["loss_mag_mae"])
    assert close(losses.SISNRLoss()(est, clean), G["loss_sdr"], 1e-4)
    est.requires_grad_(True)
    out = losses.ndns_training_loss(est, clean)
    assert close(out["loss"], G["loss"])
    out["loss"].backward()
    ref = G["grad"]
    assert np.abs(est.grad.cpu().numpy() - ref).max() <= 1e-4 * np.abs(ref).max()
