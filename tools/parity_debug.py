"""Where does the CUDA path first leave the reference on a long fixture?  (development aid, GPU box)

    python tools/parity_debug.py cfgS_2x4s

Layer by layer (teacher-forced between layers: layer l is fed the REFERENCE's spikes of layer l-1): the first frame with a
flipped spike, the oracle's membrane potential there, max|c - c_oracle| before it, and the error of the input
projection against a float64 product."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import gsn_oracle as O  # noqa: E402
from spiking_fullsubnet_b200 import Separator, SpikingFullSubNet, ops  # noqa: E402
from tests.helpers import load_long, unpack  # noqa: E402

DEV = "cuda:0"


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def main(name, backend="auto"):
    g = load_long(name)
    cfg = g["cfg"]
    cls = SpikingFullSubNet if g["surface"] == "A" else Separator
    m = cls(**cfg)
    m.load_state_dict({k: torch.from_numpy(np.array(v)) for k, v in g["params"].items()}, strict=True)
    m = m.eval().to(DEV)
    with torch.no_grad():
        _, fb_all, sb_all = m.coefficients(_t(g["mag"]))
    if g["surface"] == "A":
        _, ofb, osb = O.spiking_fullsubnet_network(g["mag"], g["params"], cfg)
    else:
        _, ofb, osb = O.separator_network(g["mag"], g["params"], cfg)
    shared = cfg.get("shared_weights", False)
    models = [("fb", "fb_model.", fb_all, ofb, cfg["fb_hidden_size"])]
    models += [(f"sb{i}", f"sb_model.sb_models.{i}.", sb_all[i], osb[i], cfg["sb_hidden_size"])
               for i in range(len(sb_all))]
    for tag, prefix, all_out, o_all, H in models:
        x = all_out[0].cpu().numpy()
        print(f"{tag}: layer-0 input max|x - x_oracle| = {np.abs(x - o_all[0]).max():.3e} (max|x| {np.abs(o_all[0]).max():.3f})")
        inp = o_all[0]  # oracle input: isolates the recurrence from the front end
        for l in range(2):
            q = f"{prefix}sequence_model.layers.{l}.cell."
            p = g["params"]
            w_ih, w_hh, bias = p[q + "weight_ih"], p[q + "weight_hh"], p[q + "bias_ih"]
            a = b = None
            bn = None
            if q + "batchnorm.weight" in p:
                bn = {k: p[q + "batchnorm." + k] for k in ("weight", "bias", "running_mean", "running_var")}
                inv = 1.0 / np.sqrt(bn["running_var"] + np.float32(1e-5))
                al = (inv * bn["weight"]).astype(np.float32)
                a, b = _t(al), _t((bn["bias"] - bn["running_mean"] * al).astype(np.float32))
            href = unpack(g[f"{tag}_h{l}"], H)
            # oracle c trace, free-running on the same input (equals the reference while spikes agree)
            pp = {"layers.0.cell.weight_ih": w_ih, "layers.0.cell.weight_hh": w_hh, "layers.0.cell.bias_ih": bias}
            if bn:
                pp.update({"layers.0.cell.batchnorm." + k: v for k, v in bn.items()})
            _, otr, ocs = O.gsn_stack_forward(inp, pp, "", 1, shared, return_c=True)
            assert np.array_equal(otr[1], href), f"{tag} layer {l}: oracle differs from the reference fixture"
            xg = _t(inp)
            xproj = ops.linear(xg, _t(w_ih), spikes=l > 0)
            x64 = inp.astype(np.float64) @ w_ih.astype(np.float64).T
            xerr = np.abs(xproj.cpu().numpy() - x64).max()
            x32 = (inp @ w_ih.T)
            h, c, _ = ops.layer_recurrence(xproj, _t(w_hh), _t(bias), a, b, shared=shared, want_c=True, backend=backend)
            h, c = h.cpu().numpy(), c.cpu().numpy()
            diff = h != href
            T = h.shape[0]
            if diff.any():
                first = int(np.argmax(diff.reshape(T, -1).any(axis=1)))
                rr, jj = np.argwhere(diff[first])[0]
                pre = np.abs(c[:first] - ocs[0][:first]).max() if first else 0.0
                print(f"  {tag} L{l}: {int(diff.sum())} flips, first at frame {first} row {rr} neuron {jj}: "
                      f"c_gpu {c[first, rr, jj]:+.3e} c_oracle {ocs[0][first, rr, jj]:+.3e}; max|dc| before = {pre:.3e}; "
                      f"xproj err vs f64 {xerr:.2e} (numpy f32: {np.abs(x32 - x64).max():.2e})")
                # error growth: max |dc| per frame up to the flip
                per = [float(np.abs(c[t] - ocs[0][t]).max()) for t in range(0, first + 1, max(1, first // 8))]
                print("     max|dc| over frames:", " ".join(f"{v:.1e}" for v in per))
            else:
                print(f"  {tag} L{l}: 0 flips; max|dc| = {np.abs(c - ocs[0]).max():.3e}; min|c_oracle| = "
                      f"{np.abs(ocs[0]).min():.2e}; xproj err vs f64 {xerr:.2e} (numpy f32: {np.abs(x32 - x64).max():.2e})")
            inp = href


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "cfgS_2x4s", sys.argv[2] if len(sys.argv) > 2 else "auto")
