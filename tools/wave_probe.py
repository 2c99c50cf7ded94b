import sys, numpy as np, torch
sys.path.insert(0, ".")
from oracle import synth
from spiking_fullsubnet_b200 import SpikingFullSubNet
torch.manual_seed(0)
for size, B, T in (("S", 40, 60), ("L", 12, 80)):
    cfg = synth.CONFIGS[size]
    m = SpikingFullSubNet(**cfg)
    m.load_state_dict({k: torch.from_numpy(np.array(v)) for k, v in synth.make_params(cfg, 5).items()})
    m = m.eval().cuda()
    mag = torch.from_numpy(synth.make_mag(B, 257, T, 11)).cuda()
    with torch.no_grad():
        m.enable_streaming(True)
        p1, fb1, sb1 = m.network(mag)
        torch.cuda.synchronize()
        print(size, B, "waves", m.stream_waves)
        # reference: each utterance group alone (batch-composition independence) through the same pipeline
        b = m.stream_waves[0]
        p2 = [torch.cat([m.network(mag[lo:lo + b])[0][k] for lo in range(0, B, b)], dim=1) for k in range(len(p1))]
        for a, c in zip(p1, p2):
            print("  equal:", torch.equal(a, c), tuple(a.shape))
        print("  traces", [tuple(x.shape) for x in fb1[:]], [tuple(x.shape) for x in sb1[0][:]])
