"""Drives the UNMODIFIED reference (vendored, git-ignored copy under baseline/_ref/) on CPU -- TEST / BENCH
INFRASTRUCTURE ONLY: imported by tests/, tests/golden/make_golden.py, `__graft_entry__` and the CPU legs of bench.py.

`vendor()` copies the few reference files the path needs (SURVEY.md 8c file list) from /root/reference into
baseline/_ref/ when the reference tree is present (the build container); baseline/_ref is git-ignored but not
gpurun-ignored, so the copy travels to the GPU box with the working tree and `bench.py --impl reference` can time
the reference's own stock `forward` code path there.  Nothing under spiking_fullsubnet_b200/ imports this module.

Import recipe: SURVEY.md Appendix C -- empty stub modules for librosa / soundfile / matplotlib (imported at module
top by audiozen/acoustics/audio_feature.py:4-7, never used on the path) and for onnxruntime / pesq / pystoi /
accelerate (audiozen/metric.py, reached through model_low_freq.py:12).
"""
from __future__ import annotations

import os
import shutil
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SRC = os.environ.get("GSN_REFERENCE_ROOT", "/root/reference")
REF_DST = os.path.join(ROOT, "baseline", "_ref")
RECIPE_B = "recipes/intel_ndns/spiking_fullsubnet_freeze_phase"

FILES = [
    "audiozen/__init__.py", "audiozen/constant.py", "audiozen/metric.py", "audiozen/utils.py",
    "audiozen/acoustics/__init__.py", "audiozen/acoustics/audio_feature.py",
    "audiozen/models/__init__.py",
    "audiozen/models/spiking_fullsubnet/__init__.py",
    "audiozen/models/spiking_fullsubnet/efficient_spiking_neuron.py",
    "audiozen/models/spiking_fullsubnet/modeling_spiking_fullsubnet.py",
    "audiozen/models/cirm_gsn/__init__.py",
    "audiozen/models/cirm_gsn/efficient_spiking_neuron.py",
    "audiozen/models/cirm_gsn/modeling_cirm_gsn.py",
    RECIPE_B + "/model_low_freq.py", RECIPE_B + "/efficient_spiking_neuron.py",
    "LICENSE",
]


def vendor(src=REF_SRC, dst=REF_DST):
    """Copy the reference files of the path (verbatim) into the git-ignored baseline/_ref/.  Returns the number of
    files copied (0 when the reference tree is absent, e.g. on the GPU box, where the shipped copy is used)."""
    if not os.path.isdir(src):
        return 0
    n = 0
    for rel in FILES:
        s = os.path.join(src, rel)
        if not os.path.exists(s):
            continue
        d = os.path.join(dst, rel)
        os.makedirs(os.path.dirname(d), exist_ok=True)
        shutil.copyfile(s, d)
        n += 1
    return n


def available(root=None):
    root = root or (REF_DST if os.path.isdir(os.path.join(REF_DST, "audiozen")) else REF_SRC)
    return os.path.isfile(os.path.join(root, "audiozen/models/spiking_fullsubnet/modeling_spiking_fullsubnet.py"))


def reference_root():
    """baseline/_ref when vendored, else the read-only reference tree (build container only)."""
    if os.path.isdir(os.path.join(REF_DST, "audiozen")):
        return REF_DST
    return REF_SRC


_loaded = {}


def load():
    """Import the reference modules; returns a namespace with MSF, ESN, CGN, MLF (surface B, may be None)."""
    if _loaded:
        return _loaded["ns"]
    root = reference_root()
    if not available(root):
        raise RuntimeError(f"reference not available under {root} (run __graft_entry__.build() in the build container)")
    for n in ["librosa", "soundfile", "matplotlib", "matplotlib.pyplot", "onnxruntime", "pesq", "pystoi", "accelerate"]:
        sys.modules.setdefault(n, types.ModuleType(n))
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    if not hasattr(sys.modules["pesq"], "pesq"):
        sys.modules["pesq"].pesq = None
    if not hasattr(sys.modules["pystoi"], "stoi"):
        sys.modules["pystoi"].stoi = None
    if not hasattr(sys.modules["accelerate"], "__version__"):
        sys.modules["accelerate"].__version__ = "stub"
    sys.path.insert(0, root)
    sys.path.insert(0, os.path.join(root, RECIPE_B))  # surface B imports its neuron by bare module name
    from audiozen.models.spiking_fullsubnet import modeling_spiking_fullsubnet as MSF
    from audiozen.models.spiking_fullsubnet import efficient_spiking_neuron as ESN
    from audiozen.models.cirm_gsn import modeling_cirm_gsn as CGN
    try:
        import model_low_freq as MLF
    except Exception:  # noqa: BLE001  (surface B needs a few more third-party names; optional for the bench)
        MLF = None
    ns = types.SimpleNamespace(MSF=MSF, ESN=ESN, CGN=CGN, MLF=MLF, root=root)
    _loaded["ns"] = ns
    return ns


def build_surface_a(cfg, params):
    """The reference's SpikingFullSubNet(**cfg) with `params` (numpy state_dict) loaded, eval mode."""
    import numpy as np
    import torch
    ns = load()
    model = ns.MSF.SpikingFullSubNet(**cfg)
    model.load_state_dict({k: torch.from_numpy(np.array(v)) for k, v in params.items()}, strict=True)
    return model.eval()


def build_surface_b(cfg, params):
    import numpy as np
    import torch
    ns = load()
    if ns.MLF is None:
        raise RuntimeError("surface B reference (model_low_freq.py) could not be imported")
    model = ns.MLF.Separator(**cfg)
    model.load_state_dict({k: torch.from_numpy(np.array(v)) for k, v in params.items()}, strict=True)
    return model.eval()


def network_a(model, mag, cfg):
    """The hot path of surface A exactly as SpikingFullSubNet.forward runs it between the STFT and the deep filter
    (modeling_spiking_fullsubnet.py:434-447): magnitude [B,F,T] -> list of coefficient tensors."""
    import torch
    with torch.no_grad():
        cm = (mag.unsqueeze(1) ** cfg["fdrc"])[..., :-1, :]
        fb_in = cm[..., : cfg["fb_input_size"], :].squeeze(1)
        fb_out, fb_all = model.fb_model(fb_in)
        fb_out = fb_out.unsqueeze(1).repeat(1, 1, (cfg["n_fft"] // 2 + 1) // cfg["fb_input_size"], 1)
        coefs, sb_all = model.sb_model(cm, fb_out)
    return coefs, fb_all, sb_all


def network_b(model, mag, cfg):
    """Surface B between the STFT and the deep filter (model_low_freq.py:574-586)."""
    import torch
    with torch.no_grad():
        cm = (mag.unsqueeze(1) ** cfg["fdrc"])[..., :-1, :]
        fb_in = model.norm(cm[..., : cfg["fb_freqs"], :]).squeeze(1)
        fb_out, fb_all = model.fb_model(fb_in)
        coefs, sb_all = model.sb_model(cm, fb_out.unsqueeze(1).repeat(1, 1, cfg["num_freqs"] // cfg["fb_freqs"], 1))
    return coefs, fb_all, sb_all
