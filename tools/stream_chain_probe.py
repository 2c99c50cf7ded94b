"""Development tool (GPU): where a two-stage counter chain loses time -- L0 publishing alone, L1 polling counters that
are already complete, and the chained pair; us per frame.  GSN_POLL_NS sets the poll back-off."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spiking_fullsubnet_b200 import ops  # noqa: E402

DEV = "cuda"
R, K, H, T = (int(a) for a in sys.argv[1:5]) if len(sys.argv) > 4 else (256, 38, 160, 501)


def t_(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def timed(fn, reps=4):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best * 1e3 / T


rs = np.random.RandomState(1)
s = 1 / np.sqrt(H)
x = t_(rs.standard_normal((T, R, K)).astype(np.float32))
w_ih0 = t_(rs.uniform(-s, s, (H, K)).astype(np.float32))
W = [(t_(rs.uniform(-s, s, (H, H)).astype(np.float32)), t_(rs.uniform(-s, s, 2 * H).astype(np.float32)),
      t_(rs.uniform(0.6, 1.2, H).astype(np.float32)), t_(rs.normal(0, 0.1, H).astype(np.float32))) for _ in range(2)]
w_ih1 = t_(rs.uniform(-s, s, (H, H)).astype(np.float32))
xproj = ops.linear(x, w_ih0)
bits0 = ops.spike_bits_buffer((T, R), H, DEV)
bits1 = ops.spike_bits_buffer((T, R), H, DEV)
ctas = ops.stream_ctas(R, H)
fused = ops.stream_ctas(R, H, H, True) > 0
cnt = ops.frame_counters(T, DEV, 2)


def l0(out_cnt=None):
    ops.recurrence_stream(W[0][0], W[0][1], W[0][2], W[0][3], xproj=xproj, out_bits=bits0, out_cnt=out_cnt)


def l1(in_cnt=None, out_cnt=None):
    ops.recurrence_stream(W[1][0], W[1][1], W[1][2], W[1][3], in_bits=bits0, w_ih=w_ih1, out_bits=bits1, in_cnt=in_cnt,
                          in_target=ctas, out_cnt=out_cnt)


print(f"R={R} K={K} H={H} T={T} poll_ns={os.environ.get('GSN_POLL_NS', 'default')}")
print(f"  L0 alone                 {timed(lambda: l0()):.3f} us/f")


def l0_pub():
    cnt[0].zero_()
    l0(cnt[0])


print(f"  L0 publishing            {timed(l0_pub):.3f} us/f")
if fused:
    print(f"  L1 alone                 {timed(lambda: l1()):.3f} us/f")
    full = torch.full_like(cnt[0], ctas)
    print(f"  L1 polling (complete)    {timed(lambda: l1(full)):.3f} us/f")
    s0, s1 = torch.cuda.Stream(), torch.cuda.Stream()

    def chained(pub1=False):
        cnt.zero_()
        cur = torch.cuda.current_stream()
        s0.wait_stream(cur)
        s1.wait_stream(cur)
        with torch.cuda.stream(s0):
            l0(cnt[0])
        with torch.cuda.stream(s1):
            l1(cnt[0], cnt[1] if pub1 else None)
        cur.wait_stream(s0)
        cur.wait_stream(s1)

    print(f"  chained L0 -> L1         {timed(chained):.3f} us/f")
    print(f"  chained, L1 publishing   {timed(lambda: chained(True)):.3f} us/f")
