// Spectral front / back end of forward() around the recurrent network, on the COMPLEX STFT as torch.stft returns it
// (interleaved re/im), so that no |stft| tensor, real / imaginary copies, `repeat` or torch.complex pass is needed
// (SURVEY 8f, rows f1 / f2):
//   k_compress_spec   : MSF:434-436 from the complex spectrum: |X|^fdrc of the first f_keep bins, transposed to the
//                       time-major layout of the network ("b f t -> t b f", MSF:108);
//   k_deepfilter_spec : MSF:315-346 for one band straight from its proj output layout (MSF:160-167), complex in,
//                       complex out; k_copy_bins passes the un-filtered bins (Nyquist) through (MSF:461-468);
//   k_overlap_add     : the synthesis half of audiozen/acoustics/audio_feature.py:297-347 (torch.istft, center=True):
//                       synthesis window, overlap-add of the 4 frames covering a sample, division by the
//                       overlap-added squared window, removal of the n_fft/2 centre padding -- one pass over the
//                       inverse-FFT frames instead of window-multiply + F.fold + divide + slice.
// All three are HBM-bound streaming kernels (8 / 16 / ~5 bytes per output element).
#include "gsn_common.cuh"

namespace gsn {

// spec [B,F,T] complex64 -> cm [T,B,Fk]: 32x32 smem tile transpose so both sides are coalesced
__global__ void __launch_bounds__(256) k_compress_spec(const float2* __restrict__ spec, float* __restrict__ cm, int B,
                                                       int F, int Fk, int T, float fdrc, int mode) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int t0 = blockIdx.x * 32, f0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int f = f0 + ty + 8 * i, t = t0 + tx;
    float v = 0.f;
    if (f < Fk && t < T) {
      const float2 z = spec[((size_t)b * F + f) * T + t];
      v = hypotf(z.x, z.y);  // torch.abs of a complex tensor
      v = mode == 0 ? sqrtf(v) : (mode == 1 ? v : powf(v, fdrc));
    }
    tile[ty + 8 * i][tx] = v;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int t = t0 + ty + 8 * i, f = f0 + tx;
    if (f < Fk && t < T) cm[((size_t)t * B + b) * Fk + f] = tile[tx][ty + 8 * i];
  }
}

// spec [B,T,F] complex64 (the layout cuFFT writes: torch.stft's result transposed back) -> cm [T,B,Fk]: coalesced on both
// sides without a transpose
__global__ void __launch_bounds__(256) k_compress_spec_tf(const float2* __restrict__ spec, float* __restrict__ cm, int B,
                                                          int F, int Fk, int T, float fdrc, int mode) {
  // one block per (t, b) row of cm: the row index is decoded once per block, not per element (the per-element
  // division / modulo cost more than the |.|^fdrc itself)
  const int b = blockIdx.x % B, t = blockIdx.x / B;
  const float2* src = spec + ((size_t)b * T + t) * F;
  float* dst = cm + (size_t)blockIdx.x * Fk;
  for (int f = threadIdx.x; f < Fk; f += blockDim.x) {
    const float2 z = src[f];
    float v = hypotf(z.x, z.y);
    v = mode == 0 ? sqrtf(v) : (mode == 1 ? v : powf(v, fdrc));
    dst[f] = v;
  }
}

// one thread per (b, s, n, fc, t); proj [T, B*N, P] with P = (c, fc, df, s) fastest-last (MSF:160-167)
__global__ void __launch_bounds__(256) k_deepfilter_spec(const float* __restrict__ proj, const float2* __restrict__ spec,
                                                         float2* __restrict__ out, float* __restrict__ mag, int T, int B,
                                                         int N, int ctr, int df, int S, int lo, int F, int F_out,
                                                         int layout) {
  const size_t total = (size_t)B * S * N * ctr * T;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int t = idx % T;
  size_t q = idx / T;
  const int fc = q % ctr; q /= ctr;
  const int n = q % N; q /= N;
  const int s = q % S;
  const int b = q / S;
  const int P = 2 * ctr * df * S;
  const float* pr = proj + ((size_t)t * B * N + (size_t)b * N + n) * P;
  const int f = lo + n * ctr + fc;
  const float2* x = spec + ((size_t)b * F + f) * T;
  float yr = 0.f, yi = 0.f;
  for (int d = 0; d < df; ++d) {
    const int tt = t - (df - 1) + d;
    if (tt < 0) continue;
    // layout 0: (c fc df s), MSF:160-167;  layout 1: (c df s fc), CGN:230
    const float cr = layout == 0 ? pr[((0 * ctr + fc) * df + d) * S + s] : pr[((0 * df + d) * S + s) * ctr + fc];
    const float ci = layout == 0 ? pr[((1 * ctr + fc) * df + d) * S + s] : pr[((1 * df + d) * S + s) * ctr + fc];
    const float2 z = x[tt];
    yr += z.x * cr - z.y * ci;
    yi += z.x * ci + z.y * cr;
  }
  const size_t o = (((size_t)b * S + s) * F_out + f) * T + t;
  out[o] = make_float2(yr, yi);
  if (mag != nullptr) mag[o] = hypotf(yr, yi);  // torch.abs of a complex tensor (enh_mag, MSF:472)
}

// time-major spectra: spec [B,T,F], out [B,S,T,F_out]; one thread per (b, s, t, band bin), bins fastest
__global__ void __launch_bounds__(256) k_deepfilter_spec_tf(const float* __restrict__ proj, const float2* __restrict__ spec,
                                                            float2* __restrict__ out, float* __restrict__ mag, int T,
                                                            int B, int N, int ctr, int df, int S, int lo, int F, int F_out,
                                                            int layout) {
  const int W = N * ctr;
  const size_t total = (size_t)B * S * T * W;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int k = idx % W;
  size_t q = idx / W;
  const int t = q % T; q /= T;
  const int s = q % S;
  const int b = q / S;
  const int n = k / ctr, fc = k - n * ctr;
  const int P = 2 * ctr * df * S;
  const float* pr = proj + ((size_t)t * B * N + (size_t)b * N + n) * P;
  const int f = lo + k;
  float yr = 0.f, yi = 0.f;
  for (int d = 0; d < df; ++d) {
    const int tt = t - (df - 1) + d;
    if (tt < 0) continue;
    const float cr = layout == 0 ? pr[((0 * ctr + fc) * df + d) * S + s] : pr[((0 * df + d) * S + s) * ctr + fc];
    const float ci = layout == 0 ? pr[((1 * ctr + fc) * df + d) * S + s] : pr[((1 * df + d) * S + s) * ctr + fc];
    const float2 z = spec[((size_t)b * T + tt) * F + f];
    yr += z.x * cr - z.y * ci;
    yi += z.x * ci + z.y * cr;
  }
  const size_t o = (((size_t)b * S + s) * T + t) * F_out + f;
  out[o] = make_float2(yr, yi);
  if (mag != nullptr) mag[o] = hypotf(yr, yi);
}

// out[b, s, t, f] = spec[b, t, f] for f in [f_lo, F) (time-major spectra)
__global__ void __launch_bounds__(256) k_copy_bins_tf(const float2* __restrict__ spec, float2* __restrict__ out,
                                                      float* __restrict__ mag, int T, int B, int S, int f_lo, int F,
                                                      int F_out) {
  const int W = F - f_lo;
  const size_t total = (size_t)B * S * T * W;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int f = f_lo + (int)(idx % W);
  size_t q = idx / W;
  const int t = q % T; q /= T;
  const int s = q % S;
  const int b = q / S;
  const float2 z = spec[((size_t)b * T + t) * F + f];
  const size_t o = (((size_t)b * S + s) * T + t) * F_out + f;
  out[o] = z;
  if (mag != nullptr) mag[o] = hypotf(z.x, z.y);
}

// out[b, s, f, t] = spec[b, f, t] for f in [f_lo, F)
__global__ void __launch_bounds__(256) k_copy_bins(const float2* __restrict__ spec, float2* __restrict__ out,
                                                   float* __restrict__ mag, int T, int B, int S, int f_lo, int F, int F_out) {
  const size_t total = (size_t)B * S * (F - f_lo) * T;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int t = idx % T;
  size_t q = idx / T;
  const int f = f_lo + (int)(q % (F - f_lo)); q /= (F - f_lo);
  const int s = q % S;
  const int b = q / S;
  const float2 z = spec[((size_t)b * F + f) * T + t];
  const size_t o = (((size_t)b * S + s) * F_out + f) * T + t;
  out[o] = z;
  if (mag != nullptr) mag[o] = hypotf(z.x, z.y);
}

// frames [B, T, n_fft] (inverse real FFT of every frame, unwindowed) -> y [B, length]
__global__ void __launch_bounds__(256) k_overlap_add(const float* __restrict__ frames, const float* __restrict__ window,
                                                     float* __restrict__ y, int B, int T, int n_fft, int hop, int length) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;  // (no 64-bit division per sample)
  if (s >= length) return;
  const size_t idx = (size_t)b * length + s;
  const int j = s + n_fft / 2;  // position in the centre-padded signal
  if (j >= n_fft + hop * (T - 1)) {  // past the last frame: torch.istft pads with zeros
    y[idx] = 0.f;
    return;
  }
  int t_hi = j / hop;
  if (t_hi > T - 1) t_hi = T - 1;
  int t_lo = (j - n_fft + hop) / hop;  // smallest t with t*hop + n_fft > j
  if (j - n_fft + 1 <= 0) t_lo = 0;
  float acc = 0.f, env = 0.f;
  const float* fb = frames + (size_t)b * T * n_fft;
  for (int t = t_lo; t <= t_hi; ++t) {
    const int pos = j - t * hop;
    const float w = window[pos];
    acc = fmaf(fb[(size_t)t * n_fft + pos], w, acc);
    env = fmaf(w, w, env);
  }
  y[idx] = acc / env;
}

// the same, four consecutive samples per thread (hop, n_fft / 2 and length multiples of 4: every load and the store are
// 16-byte accesses; the four samples share their covering frames)
__global__ void __launch_bounds__(256) k_overlap_add4(const float* __restrict__ frames, const float* __restrict__ window,
                                                      float* __restrict__ y, int B, int T, int n_fft, int hop, int length) {
  const int s = (blockIdx.x * blockDim.x + threadIdx.x) * 4, b = blockIdx.y;
  if (s >= length) return;
  float4* out = reinterpret_cast<float4*>(y + (size_t)b * length + s);
  const int j = s + n_fft / 2;  // position in the centre-padded signal
  if (j >= n_fft + hop * (T - 1)) {  // past the last frame (the signal's end is a multiple of 4 as well)
    *out = make_float4(0.f, 0.f, 0.f, 0.f);
    return;
  }
  int t_hi = j / hop;
  if (t_hi > T - 1) t_hi = T - 1;
  int t_lo = (j - n_fft + hop) / hop;  // smallest t with t*hop + n_fft > j
  if (j - n_fft + 1 <= 0) t_lo = 0;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f), env = acc;
  const float* fb = frames + (size_t)b * T * n_fft;
  for (int t = t_lo; t <= t_hi; ++t) {
    const int pos = j - t * hop;
    const float4 w = *reinterpret_cast<const float4*>(window + pos);
    const float4 f = *reinterpret_cast<const float4*>(fb + (size_t)t * n_fft + pos);
    acc.x = fmaf(f.x, w.x, acc.x); acc.y = fmaf(f.y, w.y, acc.y); acc.z = fmaf(f.z, w.z, acc.z); acc.w = fmaf(f.w, w.w, acc.w);
    env.x = fmaf(w.x, w.x, env.x); env.y = fmaf(w.y, w.y, env.y); env.z = fmaf(w.z, w.z, env.z); env.w = fmaf(w.w, w.w, env.w);
  }
  *out = make_float4(acc.x / env.x, acc.y / env.y, acc.z / env.z, acc.w / env.w);
}

// y [B, L] -> frames [B, T, n_fft]: the analysis half of torch.stft(center=True, pad_mode="constant") in front of the
// real FFT (audio_feature.py:236-294): zero padding of n_fft/2 samples on both sides, framing at `hop`, analysis window.
// One float4 per thread; hop and n_fft are multiples of 4 and y is 16-byte aligned per row when L % 4 == 0, else scalar.
__global__ void __launch_bounds__(256) k_frame_signal(const float* __restrict__ y, const float* __restrict__ window,
                                                      float* __restrict__ frames, int B, int L, int T, int n_fft, int hop) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;  // over B * T * n_fft / 4
  const int q4 = n_fft / 4;
  if (idx >= (size_t)B * T * q4) return;
  const int k = (int)(idx % q4) * 4;
  const size_t bt = idx / q4;
  const int t = (int)(bt % T), b = (int)(bt / T);
  const int s0 = t * hop + k - n_fft / 2;  // first of my four samples in the unpadded signal
  const float* yb = y + (size_t)b * L;
  float4 v;
  if (s0 >= 0 && s0 + 3 < L && ((L | hop | (n_fft / 2)) & 3) == 0) {
    v = *reinterpret_cast<const float4*>(yb + s0);
  } else {
    v.x = (s0 + 0 >= 0 && s0 + 0 < L) ? yb[s0 + 0] : 0.f;
    v.y = (s0 + 1 >= 0 && s0 + 1 < L) ? yb[s0 + 1] : 0.f;
    v.z = (s0 + 2 >= 0 && s0 + 2 < L) ? yb[s0 + 2] : 0.f;
    v.w = (s0 + 3 >= 0 && s0 + 3 < L) ? yb[s0 + 3] : 0.f;
  }
  const float4 w = *reinterpret_cast<const float4*>(window + k);
  v.x = __fmul_rn(v.x, w.x); v.y = __fmul_rn(v.y, w.y); v.z = __fmul_rn(v.z, w.z); v.w = __fmul_rn(v.w, w.w);
  *reinterpret_cast<float4*>(frames + idx * 4) = v;
}

}  // namespace gsn

extern "C" int gsn_frame_signal(const float* y, const float* window, float* frames, int B, int L, int T, int n_fft,
                                int hop, gsn_stream_t stream) {
  GSN_REQUIRE(y && window && frames, "gsn_frame_signal: null pointer");
  GSN_REQUIRE(B > 0 && L > 0 && T > 0 && n_fft > 0 && n_fft % 8 == 0 && hop > 0 && hop <= n_fft,
              "gsn_frame_signal: bad shape B=%d L=%d T=%d n_fft=%d hop=%d", B, L, T, n_fft, hop);
  GSN_REQUIRE(T == 1 + L / hop, "gsn_frame_signal: T must be 1 + L / hop (center=True), got T=%d L=%d hop=%d", T, L, hop);
  GSN_REQUIRE((reinterpret_cast<uintptr_t>(y) & 15) == 0 && (reinterpret_cast<uintptr_t>(window) & 15) == 0 &&
                  (reinterpret_cast<uintptr_t>(frames) & 15) == 0, "gsn_frame_signal: 16-byte alignment");
  const unsigned long long n4 = (unsigned long long)B * T * (n_fft / 4);
  const unsigned long long blocks = (n4 + 255) / 256;
  GSN_REQUIRE(blocks < 2147483647ULL, "gsn_frame_signal: too large");
  gsn::k_frame_signal<<<(unsigned)blocks, 256, 0, gsn::as_stream(stream)>>>(y, window, frames, B, L, T, n_fft, hop);
  GSN_LAUNCH_CHECK("k_frame_signal");
  return GSN_OK;
}

extern "C" int gsn_compress_spec(const float* spec_ri, float* cm, int B, int F, int f_keep, int T, float fdrc,
                                 int time_major, gsn_stream_t stream) {
  GSN_REQUIRE(spec_ri && cm, "gsn_compress_spec: null pointer");
  GSN_REQUIRE(B > 0 && F > 0 && T > 0 && f_keep > 0 && f_keep <= F,
              "gsn_compress_spec: bad shape B=%d F=%d f_keep=%d T=%d", B, F, f_keep, T);
  GSN_REQUIRE(B <= 65535, "gsn_compress_spec: B=%d > 65535", B);
  GSN_REQUIRE((reinterpret_cast<uintptr_t>(spec_ri) & 7) == 0, "gsn_compress_spec: spec must be 8-byte aligned");
  const int mode = fdrc == 0.5f ? 0 : (fdrc == 1.0f ? 1 : 2);
  if (time_major) {
    const size_t rows = (size_t)T * B;
    GSN_REQUIRE(rows < 2147483647ULL, "gsn_compress_spec: too many frames");
    gsn::k_compress_spec_tf<<<(unsigned)rows, f_keep >= 256 ? 256 : (f_keep + 31) / 32 * 32, 0, gsn::as_stream(stream)>>>(
        reinterpret_cast<const float2*>(spec_ri), cm, B, F, f_keep, T, fdrc, mode);
    GSN_LAUNCH_CHECK("k_compress_spec_tf");
    return GSN_OK;
  }
  dim3 grid((T + 31) / 32, (f_keep + 31) / 32, B);
  gsn::k_compress_spec<<<grid, 256, 0, gsn::as_stream(stream)>>>(reinterpret_cast<const float2*>(spec_ri), cm, B, F,
                                                                 f_keep, T, fdrc, mode);
  GSN_LAUNCH_CHECK("k_compress_spec");
  return GSN_OK;
}

extern "C" int gsn_deepfilter_spec(const float* proj, const float* spec_ri, float* out_ri, float* mag_out, int T, int B,
                                   int N, int ctr, int df, int S, int lo, int F, int F_out, int layout, int time_major,
                                   gsn_stream_t stream) {
  GSN_REQUIRE(proj && spec_ri && out_ri, "gsn_deepfilter_spec: null pointer");
  GSN_REQUIRE(T > 0 && B > 0 && N > 0 && ctr > 0 && df > 0 && S > 0, "gsn_deepfilter_spec: bad shape");
  GSN_REQUIRE(lo + N * ctr <= F && lo + N * ctr <= F_out, "gsn_deepfilter_spec: band out of range");
  GSN_REQUIRE(layout == 0 || layout == 1, "gsn_deepfilter_spec: layout %d", layout);
  GSN_REQUIRE(((reinterpret_cast<uintptr_t>(spec_ri) | reinterpret_cast<uintptr_t>(out_ri)) & 7) == 0,
              "gsn_deepfilter_spec: complex tensors must be 8-byte aligned");
  const size_t total = (size_t)B * S * N * ctr * T;
  const size_t blocks = (total + 255) / 256;
  GSN_REQUIRE(blocks < 2147483647ULL, "gsn_deepfilter_spec: too large");
  if (time_major) {
    gsn::k_deepfilter_spec_tf<<<(unsigned)blocks, 256, 0, gsn::as_stream(stream)>>>(
        proj, reinterpret_cast<const float2*>(spec_ri), reinterpret_cast<float2*>(out_ri), mag_out, T, B, N, ctr, df, S, lo,
        F, F_out, layout);
    GSN_LAUNCH_CHECK("k_deepfilter_spec_tf");
    return GSN_OK;
  }
  gsn::k_deepfilter_spec<<<(unsigned)blocks, 256, 0, gsn::as_stream(stream)>>>(
      proj, reinterpret_cast<const float2*>(spec_ri), reinterpret_cast<float2*>(out_ri), mag_out, T, B, N, ctr, df, S, lo, F,
      F_out, layout);
  GSN_LAUNCH_CHECK("k_deepfilter_spec");
  return GSN_OK;
}

extern "C" int gsn_spec_passthrough(const float* spec_ri, float* out_ri, float* mag_out, int T, int B, int S, int f_lo,
                                    int F, int F_out, int time_major, gsn_stream_t stream) {
  GSN_REQUIRE(spec_ri && out_ri, "gsn_spec_passthrough: null pointer");
  GSN_REQUIRE(T > 0 && B > 0 && S > 0 && f_lo >= 0 && f_lo <= F && F <= F_out, "gsn_spec_passthrough: bad shape");
  if (f_lo == F) return GSN_OK;
  const size_t total = (size_t)B * S * (F - f_lo) * T;
  const size_t blocks = (total + 255) / 256;
  GSN_REQUIRE(blocks < 2147483647ULL, "gsn_spec_passthrough: too large");
  if (time_major) {
    gsn::k_copy_bins_tf<<<(unsigned)blocks, 256, 0, gsn::as_stream(stream)>>>(
        reinterpret_cast<const float2*>(spec_ri), reinterpret_cast<float2*>(out_ri), mag_out, T, B, S, f_lo, F, F_out);
    GSN_LAUNCH_CHECK("k_copy_bins_tf");
    return GSN_OK;
  }
  gsn::k_copy_bins<<<(unsigned)blocks, 256, 0, gsn::as_stream(stream)>>>(
      reinterpret_cast<const float2*>(spec_ri), reinterpret_cast<float2*>(out_ri), mag_out, T, B, S, f_lo, F, F_out);
  GSN_LAUNCH_CHECK("k_copy_bins");
  return GSN_OK;
}

extern "C" int gsn_overlap_add(const float* frames, const float* window, float* y, int B, int T, int n_fft, int hop,
                               int length, gsn_stream_t stream) {
  GSN_REQUIRE(frames && window && y, "gsn_overlap_add: null pointer");
  GSN_REQUIRE(B > 0 && T > 0 && n_fft > 0 && hop > 0 && hop <= n_fft && length > 0, "gsn_overlap_add: bad shape");
  GSN_REQUIRE(B <= 65535, "gsn_overlap_add: B=%d > 65535", B);
  const bool vec4 = ((hop | (n_fft / 2) | length) & 3) == 0 &&
                    ((reinterpret_cast<uintptr_t>(frames) | reinterpret_cast<uintptr_t>(window) |
                      reinterpret_cast<uintptr_t>(y)) & 15) == 0;
  if (vec4) {
    gsn::k_overlap_add4<<<dim3((length / 4 + 255) / 256, B), 256, 0, gsn::as_stream(stream)>>>(frames, window, y, B, T, n_fft,
                                                                                            hop, length);
    GSN_LAUNCH_CHECK("k_overlap_add4");
    return GSN_OK;
  }
  gsn::k_overlap_add<<<dim3((length + 255) / 256, B), 256, 0, gsn::as_stream(stream)>>>(frames, window, y, B, T, n_fft, hop,
                                                                                     length);
  GSN_LAUNCH_CHECK("k_overlap_add");
  return GSN_OK;
}
