"""Generate the golden fixtures in this directory by RUNNING THE REFERENCE (read-only import from
/root/reference, CPU, torch fp32).  Run in the build container only:

    python tests/golden/make_golden.py

The GPU box has no /root/reference; tests only read the committed .npz files.
Import recipe: SURVEY.md Appendix C (stub modules for librosa/soundfile/matplotlib, which
audiozen/acoustics/audio_feature.py:4-7 imports at module top but the path never uses).
"""
import json
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("GSN_REFERENCE_ROOT", "/root/reference")
sys.path.insert(0, ROOT)

for n in ["librosa", "soundfile", "matplotlib", "matplotlib.pyplot", "onnxruntime", "pesq", "pystoi", "accelerate"]:
    sys.modules.setdefault(n, types.ModuleType(n))
sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
sys.modules["pesq"].pesq = sys.modules["pystoi"].stoi = None
sys.modules["accelerate"].__version__ = "stub"
sys.path.insert(0, REF)
sys.path.insert(0, REF + "/recipes/intel_ndns/spiking_fullsubnet_freeze_phase")  # surface B imports by bare name

from audiozen.models.spiking_fullsubnet import modeling_spiking_fullsubnet as MSF  # noqa: E402
from audiozen.models.spiking_fullsubnet.efficient_spiking_neuron import GSUCell  # noqa: E402
from audiozen.models.cirm_gsn import modeling_cirm_gsn as CGN  # noqa: E402
from audiozen.models.cirm_gsn.efficient_spiking_neuron import GSUCell as GSUCellC  # noqa: E402

from oracle import synth  # noqa: E402


def pack(h):
    """[T,R,H] {0,1} float -> packed bits along H."""
    a = h.detach().numpy()
    assert ((a == 0) | (a == 1)).all()
    return np.packbits(a.astype(np.uint8), axis=-1)


def load_params(model, params):
    sd = {k: torch.from_numpy(np.array(v)) for k, v in params.items()}
    model.load_state_dict(sd, strict=True)
    model.eval()


def hook_cells(model, cell_cls):
    """Collect c_t per step for every GSUCell, keyed by module name."""
    traces = {}

    def mk(name):
        def hook(_m, _inp, out):
            traces.setdefault(name, []).append(out[1][1].detach().numpy().copy())
        return hook

    hs = [m.register_forward_hook(mk(n)) for n, m in model.named_modules() if isinstance(m, cell_cls)]
    return traces, hs


def run_surface_a(name, cfg, seed, batch, num_samples, with_c):
    torch.manual_seed(0)
    model = MSF.SpikingFullSubNet(**cfg)
    params = synth.make_params(cfg, seed)
    load_params(model, params)
    wave = synth.make_wave(batch, num_samples, seed + 1)
    traces, hooks = hook_cells(model, GSUCell) if with_c else ({}, [])
    x = torch.from_numpy(wave)
    with torch.no_grad():
        mag = model.stft(x)[0]
        out = model(x)
    for h in hooks:
        h.remove()
    d = {"cfg": json.dumps(cfg), "seed": seed, "wave": wave, "mag": mag.numpy()}
    if cfg.get("num_spks", 1) > 1:
        enh_y, fb_all, sb_all = out
    else:
        enh_y, enh_mag, fb_all, sb_all = out
        d["enh_mag"] = enh_mag.numpy()
    d["enh_y"] = enh_y.numpy()
    # coefficient tensors (the "cIRM" of BASELINE.json): re-run the network part to fetch them
    with torch.no_grad():
        cm = (mag.unsqueeze(1) ** cfg["fdrc"])[..., :-1, :]
        fb_in = cm[..., : cfg["fb_input_size"], :].squeeze(1)
        fb_out, _ = model.fb_model(fb_in)
        fb_out = fb_out.unsqueeze(1).repeat(1, 1, (cfg["n_fft"] // 2 + 1) // cfg["fb_input_size"], 1)
        coefs, _ = model.sb_model(cm, fb_out)
    for i, c in enumerate(coefs):
        d[f"coef{i}"] = c.numpy()
    L = cfg["fb_num_layers"]
    d["fb_xnorm"] = fb_all[0].numpy()
    for l in range(L):
        d[f"fb_h{l}"] = pack(fb_all[1 + l])
    d["fb_proj"] = fb_all[-1].numpy()
    for i, al in enumerate(sb_all):
        d[f"sb{i}_xnorm"] = al[0].numpy()
        for l in range(cfg["sb_num_layers"]):
            d[f"sb{i}_h{l}"] = pack(al[1 + l])
        if with_c:
            d[f"sb{i}_proj"] = al[-1].numpy()
    if with_c:
        for k, v in traces.items():  # hooks were removed before the second (network-only) pass
            d["c__" + k] = np.stack(v)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **d)
    print(name, {k: (v.shape if hasattr(v, "shape") else v) for k, v in d.items() if k != "cfg"})


def run_surface_b(name, cfg, params, seed, batch, num_samples, store_weights=False, module="model_low_freq"):
    """Surface B `Separator` (recipes/.../model_low_freq.py:485-618).  module="model_low_freq_count_time" is the
    variant of the same recipe directory whose cumulative_laplace_norm accepts the 5-D sub-band input (:173-204)."""
    import importlib
    MLF = importlib.import_module(module)  # reference module, imported from the recipe directory
    torch.manual_seed(0)
    model = MLF.Separator(**cfg)
    load_params(model, params)
    wave = synth.make_wave(batch, num_samples, seed + 1)
    x = torch.from_numpy(wave)
    with torch.no_grad():
        mag = model.stft(x)[0]
        enh_y, enh_mag, fb_all, sb_all = model(x)
        # coefficient tensors: re-run the network part (MLF:574-586)
        cm = (mag.unsqueeze(1) ** cfg["fdrc"])[..., :-1, :]
        fb_in = model.norm(cm[..., : cfg["fb_freqs"], :]).squeeze(1)
        fb_out, _ = model.fb_model(fb_in)
        coefs, _ = model.sb_model(cm, fb_out.unsqueeze(1).repeat(1, 1, cfg["num_freqs"] // cfg["fb_freqs"], 1))
    d = {"cfg": json.dumps(cfg), "seed": seed, "wave": wave, "mag": mag.numpy(), "enh_y": enh_y.numpy(),
         "enh_mag": enh_mag.numpy(), "fb_x": fb_all[0].numpy(), "fb_proj": fb_all[-1].numpy()}
    for i, c in enumerate(coefs):
        d[f"coef{i}"] = c.numpy()
    for l in range(2):
        d[f"fb_h{l}"] = pack(fb_all[1 + l])
    for i, al in enumerate(sb_all):
        d[f"sb{i}_x"] = al[0].numpy()
        for l in range(2):
            d[f"sb{i}_h{l}"] = pack(al[1 + l])
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **d)
    if store_weights:
        np.savez_compressed(os.path.join(HERE, name + "_weights.npz"), **{k: np.asarray(v) for k, v in params.items()})
    print(name, {k: (v.shape if hasattr(v, "shape") else v) for k, v in d.items() if k != "cfg"})


def _flip_frac(all_a, all_b):
    """fraction of differing spikes between two all_layer_outputs structures (layers 1..L of every model)."""
    flips = total = 0
    for a, b in zip(all_a, all_b):
        flips += int((a != b).sum())
        total += a.numel()
    return flips, total


def _spike_layers(fb_all, sb_all, L=2):
    return [fb_all[1 + l] for l in range(L)] + [al[1 + l] for al in sb_all for l in range(L)]


def _coef_rel(ca, cb):
    return max(float((a - b).abs().max() / (b.abs().max() + 1e-30)) for a, b in zip(ca, cb))


SNAP = 32


def run_long(name, surface, cfg, params, seed, batch, num_samples, coef_tail=0, weights_file=None):
    """Free-running parity fixture at BASELINE sizes (T = 501 / 1 251): full bit-packed spike traces, coefficient
    tensors (all frames, or the last `coef_tail` frames -- chaos grows with T, so the tail is the telling part),
    the enhanced waveform, and the REFERENCE'S OWN NOISE FLOOR measured in the same run: the reference against
    itself with the input scaled by 1 + 1e-6 (SURVEY 8c protocol P3) -- spike-flip fraction and coefficient
    max|delta| / max|ref|.  surface: "A" (SpikingFullSubNet) or "B" (Separator)."""
    from oracle import run_reference as RR
    torch.manual_seed(0)
    model = (RR.build_surface_a if surface == "A" else RR.build_surface_b)(cfg, params)
    net = RR.network_a if surface == "A" else RR.network_b
    wave = synth.make_wave(batch, num_samples, seed + 1)
    x = torch.from_numpy(wave)
    with torch.no_grad():
        mag = model.stft(x)[0]
        enh_y, enh_mag, fb_all, sb_all = model(x)
        # membrane snapshots every SNAP frames (state after frames SNAP-1, 2*SNAP-1, ...) for the block
        # teacher-forced protocol: the CUDA recurrence restarts from the reference's own (h, c) every SNAP frames
        snaps, counters, hooks = {}, {}, []
        cell_types = tuple(c.GSUCell for c in (RR.load().ESN, sys.modules.get("efficient_spiking_neuron"))
                           if c is not None and hasattr(c, "GSUCell"))

        def mk(key):
            def hook(_m, _inp, out):
                t = counters.get(key, 0)
                counters[key] = t + 1
                if (t + 1) % SNAP == 0:
                    snaps.setdefault(key, []).append(out[1][1].detach().numpy().copy())
            return hook

        for n_, m_ in model.named_modules():
            if isinstance(m_, cell_types):
                tag = "fb" if n_.startswith("fb_model") else "sb" + n_.split("sb_models.")[1].split(".")[0]
                hooks.append(m_.register_forward_hook(mk(f"{tag}_c{n_.split('layers.')[1].split('.')[0]}")))
        coefs, fb2, sb2 = net(model, mag, cfg)
        for h_ in hooks:
            h_.remove()
        assert _flip_frac(_spike_layers(fb_all, sb_all), _spike_layers(fb2, sb2))[0] == 0  # run-to-run deterministic
        # noise floor: the reference against itself under input scalings within +-2e-6 (worst of four)
        flips, total, floor_coef = 0, 1, 0.0
        for eps in (1e-6, -1e-6, 2e-6, -2e-6):
            coefs_n, fb_n, sb_n = net(model, mag * (1.0 + eps), cfg)
            f_, total = _flip_frac(_spike_layers(fb_n, sb_n), _spike_layers(fb_all, sb_all))
            flips = max(flips, f_)
            floor_coef = max(floor_coef, _coef_rel(coefs_n, coefs))
    d = {"cfg": json.dumps(cfg), "seed": seed, "surface": surface, "wave": wave, "mag": mag.numpy(), "enh_y": enh_y.numpy(),
         "floor_flips": flips, "floor_total": total, "floor_coef_rel": floor_coef,
         "coef_tail": coef_tail, "snap": SNAP}
    for k, v in snaps.items():
        d[k] = np.stack(v)
    T = mag.shape[-1]
    t0 = T - coef_tail if coef_tail else 0
    for i, c in enumerate(coefs):  # A: [B,df,S,F,T,2]; B: [B,df,F,T,2]
        d[f"coef{i}"] = c[..., t0:, :].contiguous().numpy()
    for l in range(2):
        d[f"fb_h{l}"] = pack(fb_all[1 + l])
    for i, al in enumerate(sb_all):
        for l in range(2):
            d[f"sb{i}_h{l}"] = pack(al[1 + l])
    if weights_file:
        d["weights_file"] = weights_file
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **d)
    print(name, f"T={T} floor: {flips}/{total} flips ({flips / total:.2e}), coef rel {d['floor_coef_rel']:.2e}",
          {k: (v.shape if hasattr(v, "shape") else v) for k, v in d.items() if k not in ("cfg", "wave", "mag")})


def run_train(name, cfg, seed, batch, num_samples):
    """One TRAINING step of surface A on the reference: train-mode BatchNorm (batch statistics per frame,
    running-stat updates) and autograd BPTT through the Triangle surrogate (ESN:84-101)."""
    torch.manual_seed(0)
    model = MSF.SpikingFullSubNet(**cfg)
    params = synth.make_params(cfg, seed)
    load_params(model, params)
    model.train()
    wave = synth.make_wave(batch, num_samples, seed + 1)
    target = synth.make_wave(batch, num_samples, seed + 2)
    x = torch.from_numpy(wave)
    enh_y, enh_mag, fb_all, sb_all = model(x)
    loss = (enh_y * torch.from_numpy(target)).sum() + enh_mag.pow(2).mean()
    loss.backward()
    d = {"cfg": json.dumps(cfg), "seed": seed, "wave": wave, "target": target, "loss": loss.detach().numpy(),
         "enh_y": enh_y.detach().numpy()}
    for l in range(cfg["fb_num_layers"]):
        d[f"fb_h{l}"] = pack(fb_all[1 + l])
    for i, al in enumerate(sb_all):
        for l in range(cfg["sb_num_layers"]):
            d[f"sb{i}_h{l}"] = pack(al[1 + l])
    for k, p in model.named_parameters():
        d["grad__" + k] = p.grad.numpy()
    for k, b in model.named_buffers():
        d["buf__" + k] = b.detach().numpy()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **d)
    print(name, "loss", float(loss), {k: v.shape for k, v in d.items() if k.startswith("grad__")}.__len__(), "grads")


def run_cirm(name, cfg, seed, batch, num_samples):
    torch.manual_seed(0)
    model = CGN.Model(**cfg)
    params = synth.make_params_cirm(cfg, seed)
    load_params(model, params)
    wave = synth.make_wave(batch, num_samples, seed + 1)
    traces, hooks = hook_cells(model, GSUCellC)
    x = torch.from_numpy(wave)
    with torch.no_grad():
        mag = model.stft(x)[0]
        enh_y, enh_mag = model(x)
        fb_out, all_out = model.fb_model(mag ** cfg["fdrc"])
    for h in hooks:
        h.remove()
    d = {"cfg": json.dumps(cfg), "seed": seed, "wave": wave, "mag": mag.numpy(), "enh_y": enh_y.numpy(),
         "enh_mag": enh_mag.numpy(), "fb_out": fb_out.numpy(), "fb_xnorm": all_out[0].numpy()}
    for l in range(cfg["num_layers"]):
        d[f"fb_h{l}"] = pack(all_out[1 + l])
    for k, v in traces.items():  # hooks fired twice (full forward, then fb_model alone): keep pass 1
        T = len(v) // 2
        d["c__" + k] = np.stack(v[:T])
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **d)
    print(name, {k: (v.shape if hasattr(v, "shape") else v) for k, v in d.items() if k != "cfg"})


def main_long():
    """BASELINE config-2 / config-3 sized fixtures: zoo-S and zoo-L `Separator` at 2 clips x 4 s
    (T = 501) and 1 clip x 10 s (T = 1 251), surface-A S (random init) at 2 x 4 s."""
    zoo_s = torch.load(REF + "/model_zoo/intel_ndns/spike_fsb/baseline_s/checkpoints/best/pytorch_model.bin",
                       map_location="cpu")
    zoo_l = torch.load(REF + "/model_zoo/intel_ndns/spike_fsb/baseline_l/checkpoints/best/pytorch_model.bin",
                       map_location="cpu")
    ps = {k: v.numpy() for k, v in zoo_s.items()}
    pl = {k: v.numpy() for k, v in zoo_l.items()}
    np.savez_compressed(os.path.join(HERE, "zoo_l_weights.npz"), **pl)
    run_long("zoo_s_2x4s", "B", synth.CFG_ZOO_S, ps, 201, 2, 64000, weights_file="zoo_s_1s_weights")
    run_long("zoo_s_1x10s", "B", synth.CFG_ZOO_S, ps, 202, 1, 160000, coef_tail=128, weights_file="zoo_s_1s_weights")
    run_long("zoo_l_2x4s", "B", synth.CFG_ZOO_L, pl, 203, 2, 64000, coef_tail=128, weights_file="zoo_l_weights")
    run_long("zoo_l_1x10s", "B", synth.CFG_ZOO_L, pl, 204, 1, 160000, coef_tail=64, weights_file="zoo_l_weights")
    run_long("cfgS_2x4s", "A", synth.CFG_S, synth.make_params(synth.CFG_S, 5), 205, 2, 64000)


if __name__ == "__main__":
    torch.set_num_threads(4)
    if len(sys.argv) > 1 and sys.argv[1] == "long":
        main_long()
        sys.exit(0)
    # tiny structural variants with per-step membrane traces (teacher-forced protocol P1)
    run_surface_a("tiny_shared_bn", synth.tiny_cfg(), 101, 2, 16 * 39, True)
    run_surface_a("tiny_unshared_nobn", synth.tiny_cfg(shared_weights=False, bn=False,
                                                       use_pre_layer_norm_fb=False,
                                                       use_pre_layer_norm_sb=False), 102, 3, 16 * 24, True)
    run_surface_a("tiny_spk2_tanh", synth.tiny_cfg(num_spks=2, fb_output_activate_function="tanh",
                                                   sb_num_layers=1, fb_num_layers=3), 103, 2, 16 * 20, True)
    run_cirm("tiny_cirm", dict(synth.CFG_CIRM, n_fft=64, hop_length=16, win_length=64, input_size=33,
                               hidden_size=40, num_layers=3, proj_size=33, df_order=2), 104, 2, 16 * 30)
    # training step (config 4 in miniature): gradients of every parameter + updated BatchNorm buffers
    run_train("tiny_train_shared_bn", synth.tiny_cfg(), 107, 3, 16 * 19)
    run_train("tiny_train_unshared_nobn", synth.tiny_cfg(shared_weights=False, bn=False), 108, 2, 16 * 15)
    # surface B: tiny structural fixture + the TRAINED model-zoo S checkpoint on a 1 s clip (protocol P3)
    cfgb = synth.tiny_cfg_b()
    run_surface_b("tiny_surface_b", cfgb, synth.make_params_b(cfgb, 105), 105, 2, 16 * 33)
    cfgc = synth.tiny_cfg_b(norm_type="cumulative_laplace_norm")  # recipes .../baseline_{s,l}.toml:63 select this norm
    run_surface_b("tiny_surface_b_cumnorm", cfgc, synth.make_params_b(cfgc, 107), 107, 2, 16 * 33,
                  module="model_low_freq_count_time")
    zoo = torch.load(REF + "/model_zoo/intel_ndns/spike_fsb/baseline_s/checkpoints/best/pytorch_model.bin",
                     map_location="cpu")
    run_surface_b("zoo_s_1s", synth.CFG_ZOO_S, {k: v.numpy() for k, v in zoo.items()}, 106, 2, 16000,
                  store_weights=True)
    # config 1 of BASELINE.json: single 1 s clip, baseline_m, CPU forward (protocol P2)
    run_surface_a("cfg1_baseline_m_1s", synth.CFG_M, 20220815, 1, 16000, False)
