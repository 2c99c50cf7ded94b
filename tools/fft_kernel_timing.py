"""Standalone timing of the fused FFT kernels at the headline shape (S, batch 32 x 4 s): gsn_stft_compress and
gsn_deepfilter_irfft + gsn_overlap_add against the kernel sequences they replace.  Run under ncu for the details."""
import sys, torch
sys.path.insert(0, ".")
from spiking_fullsubnet_b200 import ops

dev = "cuda:0"
B, L, hop = 32, 64000, 128
y = torch.randn(B, L, device=dev)
win = torch.hann_window(512, device=dev)
flush = torch.empty(64 << 20, device=dev)


def timed(fn, n=10):
    """Mean GPU duration of everything fn() launches (torch profiler: CUDA-event brackets would time the Python launch
    overhead of these 10 - 50 us kernels), L2 flushed in front of every call."""
    from torch.profiler import profile, ProfilerActivity
    fn(); torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(n):
            flush.fill_(1.0)
            fn()
        torch.cuda.synchronize()
    tot = sum(e.time_range.end - e.time_range.start for e in prof.events()
              if e.device_type == torch.autograd.DeviceType.CUDA and "FillFunctor" not in e.name)
    return tot / n


spec, cm = ops.stft_compress(y, win, hop, 256, 0.5)
T = spec.shape[2]
bands = [(8, 4, 3), (3, 32, 1), (2, 64, 1)]  # (N, ctr, df) of the S recipe: 256 bins
projs = [torch.randn(T, B * n, 2 * c * d, device=dev) for n, c, d in bands]
Ns, ctrs, dfs = ([b[i] for b in bands] for i in range(3))
print("stft_compress            %7.1f us" % timed(lambda: ops.stft_compress(y, win, hop, 256, 0.5)))
print("stft (no cm)             %7.1f us" % timed(lambda: ops.stft_compress(y, win, hop)))
print("frame+cuFFT+compress     %7.1f us" % timed(lambda: ops.compress_mag(torch.fft.rfft(ops.frame_signal(y, win, hop), dim=-1).transpose(1, 2), 256, 0.5)))
print("irfft_frames             %7.1f us" % timed(lambda: ops.irfft_frames(spec)))
print("deepfilter_irfft         %7.1f us" % timed(lambda: ops.deepfilter_irfft(projs, spec, Ns, ctrs, dfs)))
print("deepfilter_irfft no mag  %7.1f us" % timed(lambda: ops.deepfilter_irfft(projs, spec, Ns, ctrs, dfs, want_mag=False)))
fr = ops.irfft_frames(spec)
print("overlap_add              %7.1f us" % timed(lambda: ops.overlap_add(fr, win, hop, L)))
print("cuFFT irfft              %7.1f us" % timed(lambda: torch.fft.irfft(spec.transpose(1, 2), n=512, dim=-1)))
