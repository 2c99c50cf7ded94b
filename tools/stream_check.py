"""Development tool (GPU): the streaming recurrence kernel against the chunk-launch tcgen05 kernel (bit-identical
expected), standalone and chained through frame counters on two streams; prints us/frame."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spiking_fullsubnet_b200 import ops  # noqa: E402

DEV = "cuda"


def t_(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def timed(fn, reps=3):
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return out, best


def run(R, K, H, T, bn=True):
    rs = np.random.RandomState(R + H)
    s = 1 / np.sqrt(H)
    x = t_(rs.standard_normal((T, R, K)).astype(np.float32))
    w_ih0 = t_(rs.uniform(-s, s, (H, K)).astype(np.float32))
    W = [(t_(rs.uniform(-s, s, (H, H)).astype(np.float32)), t_(rs.uniform(-s, s, 2 * H).astype(np.float32)),
          t_(rs.uniform(0.6, 1.2, H).astype(np.float32)) if bn else None,
          t_(rs.normal(0, 0.1, H).astype(np.float32)) if bn else None) for _ in range(2)]
    w_ih1 = t_(rs.uniform(-s, s, (H, H)).astype(np.float32))
    xproj = ops.linear(x, w_ih0)
    # reference path: chunk-launch kernel
    bits0 = ops.spike_bits_buffer((T, R), H, DEV)
    (h0, c0, _), ms_old = timed(lambda: ops.layer_recurrence(xproj, W[0][0], W[0][1], W[0][2], W[0][3], want_c=True,
                                                             backend="tcgen05", out_bits=bits0))
    xp1 = ops.linear(h0, w_ih1, spikes=True, bits=bits0)
    bits1 = ops.spike_bits_buffer((T, R), H, DEV)
    h1, c1, _ = ops.layer_recurrence(xp1, W[1][0], W[1][1], W[1][2], W[1][3], want_c=True, backend="tcgen05",
                                     out_bits=bits1)
    torch.cuda.synchronize()
    # streaming kernel, standalone
    cbuf = torch.empty_like(c0)
    sb0, ms_new = timed(lambda: ops.recurrence_stream(W[0][0], W[0][1], W[0][2], W[0][3], xproj=xproj, out_c=cbuf))
    ok0 = torch.equal(sb0, bits0) and torch.equal(cbuf, c0)
    fused = ops.stream_ctas(R, H, H, True) > 0
    msg = f"R={R} K={K} H={H} T={T}: L0 old {ms_old * 1e3 / T:.2f} us/f, stream {ms_new * 1e3 / T:.2f} us/f, identical={ok0}"
    if fused:
        c1buf = torch.empty_like(c1)
        sb1, ms_f = timed(lambda: ops.recurrence_stream(W[1][0], W[1][1], W[1][2], W[1][3], in_bits=bits0, w_ih=w_ih1,
                                                        out_c=c1buf))
        ok1 = torch.equal(sb1, bits1) and torch.equal(c1buf, c1)
        msg += f"; L1 fused {ms_f * 1e3 / T:.2f} us/f identical={ok1}"
        # chained through counters on two streams: L1 consumes L0's bits frame by frame
        cnt = ops.frame_counters(T, DEV)
        s0, s1 = torch.cuda.Stream(), torch.cuda.Stream()
        ob0 = torch.zeros_like(bits0)
        ob1 = torch.zeros_like(bits1)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        s0.wait_stream(torch.cuda.current_stream())
        s1.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s1):  # consumer first: it must wait for the producer's counters
            ops.recurrence_stream(W[1][0], W[1][1], W[1][2], W[1][3], in_bits=ob0, w_ih=w_ih1, out_bits=ob1,
                                  in_cnt=cnt[0], in_target=ops.stream_ctas(R, H))
        with torch.cuda.stream(s0):
            ops.recurrence_stream(W[0][0], W[0][1], W[0][2], W[0][3], xproj=xproj, out_bits=ob0, out_cnt=cnt[0])
        torch.cuda.current_stream().wait_stream(s0)
        torch.cuda.current_stream().wait_stream(s1)
        e1.record()
        torch.cuda.synchronize()
        okc = torch.equal(ob0, bits0) and torch.equal(ob1, bits1)
        msg += f"; chained L0->L1 {e0.elapsed_time(e1) * 1e3 / T:.2f} us/f identical={okc} cnt_ok={bool((cnt[0] == ops.stream_ctas(R, H)).all())}"
    print(msg, flush=True)


if __name__ == "__main__":
    shapes = [(32, 64, 240, 501), (256, 38, 160, 501), (96, 94, 160, 501), (64, 158, 160, 501), (37, 20, 100, 40),
              (130, 12, 320, 60), (70, 33, 268, 33), (1536, 38, 256, 64)]
    if len(sys.argv) > 1:
        shapes = [tuple(map(int, sys.argv[1:5]))]
    for sh in shapes:
        try:
            run(*sh)
        except Exception as e:  # noqa: BLE001
            print(sh, "FAILED:", repr(e)[:300], flush=True)
