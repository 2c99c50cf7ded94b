// The 512-point real FFTs of forward() fused with their neighbours, for the recipes' STFT
// (n_fft = win_length = 512, hann window, center=True, pad_mode="constant"; audio_feature.py:236-347):
//   k_stft512     : zero padding + framing + analysis window + real FFT + |X|^fdrc ("b f t -> t b f", MSF:434-436,
//                   MSF:108) in ONE pass over the waveform: wave in, complex spectrum [B,T,257] and the network's
//                   compressed magnitude [T,B,f_keep] out (was: k_frame_signal + cuFFT R2C + k_compress_spec_tf, with the
//                   windowed frames and the spectrum making a round trip through HBM in between);
//   k_irfft512<DF>: inverse real FFT of every frame, 1/n normalised, unwindowed (what gsn_overlap_add reads).
//                   DF = false: of a given spectrum (was: torch's defensive clone + cuFFT C2R + a scaling kernel);
//                   DF = true : of the DEEP-FILTERED spectrum, computed on the fly from the bands' proj outputs
//                   (MSF:315-346, 449-472) -- the enhanced spectrum never exists in HBM, only its magnitude
//                   (enh_mag, MSF:472) and the time-domain frames are written.
// A frame is 256 complex points (even / odd samples packed into one complex sequence) transformed as 16 x 16: SIXTEEN
// threads per frame, each a 16-point transform in registers (two radix-4 stages), the 256-point twiddles (kept in
// registers, computed once per thread), a transpose through padded shared memory, a second 16-point transform, then the
// usual split / merge step of a real FFT with the partner bin fetched from shared memory.  Both frames of a warp are
// private to it, so the only synchronisation is __syncwarp; global loads and stores are 128-byte rows.  A block walks
// over the frames with a grid stride; its 512-point twiddle table and the window are built / staged once.
#include "gsn_common.cuh"
#include "gsn_fft_tables.cuh"

namespace gsn {

constexpr int FFT_N = 512, FFT_M = 256, FFT_THREADS = 128, FFT_FPB = FFT_THREADS / 16, FFT_SM = 16 * 17;

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ float2 cmulc(float2 a, float2 b) {  // a * conj(b)
  return make_float2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y);
}

template <bool INV>
__device__ __forceinline__ void dft4(float2& a0, float2& a1, float2& a2, float2& a3) {
  const float2 t0 = make_float2(a0.x + a2.x, a0.y + a2.y), t1 = make_float2(a0.x - a2.x, a0.y - a2.y);
  const float2 t2 = make_float2(a1.x + a3.x, a1.y + a3.y), d = make_float2(a1.x - a3.x, a1.y - a3.y);
  const float2 t3 = INV ? make_float2(-d.y, d.x) : make_float2(d.y, -d.x);  // +i d : -i d
  a0 = make_float2(t0.x + t2.x, t0.y + t2.y);
  a1 = make_float2(t1.x + t3.x, t1.y + t3.y);
  a2 = make_float2(t0.x - t2.x, t0.y - t2.y);
  a3 = make_float2(t1.x - t3.x, t1.y - t3.y);
}

// exp(-2 pi i m / 16) for the products m = b * c the 4 x 4 decomposition needs (compile-time after unrolling)
__device__ __forceinline__ float2 w16(int m) {
  constexpr float C1 = 0.92387953251128674f, S1 = 0.38268343236508977f, H = 0.70710678118654752f;
  switch (m) {
    case 0: return make_float2(1.f, 0.f);
    case 1: return make_float2(C1, -S1);
    case 2: return make_float2(H, -H);
    case 3: return make_float2(S1, -C1);
    case 4: return make_float2(0.f, -1.f);
    case 6: return make_float2(-H, -H);
    default: return make_float2(-C1, S1);  // m = 9
  }
}

// position of output k of dft16 (input v[n] at position n)
__device__ __forceinline__ constexpr int pos16(int k) { return 4 * (k & 3) + (k >> 2); }

// 16-point DFT in registers: input x[n] at v[n], output X[k] at v[pos16(k)].  INV: unnormalised inverse.
template <bool INV>
__device__ __forceinline__ void dft16(float2 (&v)[16]) {
#pragma unroll
  for (int b = 0; b < 4; ++b) dft4<INV>(v[b], v[4 + b], v[8 + b], v[12 + b]);  // over a of x[4a + b]: y[b][c] at v[4c + b]
#pragma unroll
  for (int c = 1; c < 4; ++c)
#pragma unroll
    for (int b = 1; b < 4; ++b) v[4 * c + b] = INV ? cmulc(v[4 * c + b], w16(b * c)) : cmul(v[4 * c + b], w16(b * c));
#pragma unroll
  for (int c = 0; c < 4; ++c) dft4<INV>(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);  // over b: X[c + 4d] at v[4c + d]
}

// 256-point DFT of one frame by its 16 threads: thread j holds x[16 n1 + j] at v[n1] on entry and X[j + 16 k2] at
// v[pos16(k2)] on return.  w256[16 k1 + j] = exp(-2 pi i j k1 / 256).  s: the frame's FFT_SM scratch entries.
template <bool INV>
__device__ __forceinline__ void fft256(float2 (&v)[16], const float2* __restrict__ w256, float2* s, int j) {
  dft16<INV>(v);
#pragma unroll
  for (int k1 = 0; k1 < 16; ++k1) {
    float2 y = v[pos16(k1)];
    if (k1 > 0) y = INV ? cmulc(y, w256[16 * k1 + j]) : cmul(y, w256[16 * k1 + j]);
    s[k1 * 17 + j] = y;
  }
  __syncwarp();
#pragma unroll
  for (int n2 = 0; n2 < 16; ++n2) v[n2] = s[j * 17 + n2];
  __syncwarp();
  dft16<INV>(v);
}

// tw[k] = exp(-2 pi i k / 512), k = 0 .. 256;  w256[16 k1 + j] = exp(-2 pi i j k1 / 256)  (from the float64-rounded table)
__device__ __forceinline__ void fft_tables(float2* tw, float2* w256) {
  for (int i = threadIdx.x; i <= FFT_M; i += blockDim.x) tw[i] = kW512[i];
  for (int i = threadIdx.x; i < 256; i += blockDim.x) w256[i] = kW512[(2 * (i >> 4) * (i & 15)) & 511];
}

// |z| as torch.abs of a complex number gives it (hypotf), without hypotf's scaling on the common path: when x^2 + y^2
// neither overflows nor loses bits to underflow, sqrt(fma(x, x, y*y)) is within one ulp of it
__device__ __noinline__ float cabs_slow(float x, float y) { return hypotf(x, y); }  // (one copy: code size)
__device__ __forceinline__ float cabs_fast(float2 z) {
  const float v = fmaf(z.x, z.x, z.y * z.y);
  if (v > 1e-30f && v < 1e30f) return sqrtf(v);
  return cabs_slow(z.x, z.y);
}
__device__ __forceinline__ float compress_abs(float2 z, float fdrc, int mode) {
  const float v = cabs_fast(z);
  return mode == 0 ? sqrtf(v) : (mode == 1 ? v : powf(v, fdrc));
}

// y [B,L] -> spec [B,T,257] complex, cm [T,B,Fk] (NULL: not wanted)
template <bool PAIRS>
__global__ void __launch_bounds__(FFT_THREADS, 8) k_stft512(const float* __restrict__ y, const float* __restrict__ window,
                                                         float2* __restrict__ spec, float* __restrict__ cm, int B, int L,
                                                         int T, int hop, int Fk, float fdrc, int mode) {
  __shared__ float2 tw[FFT_M + 1], w256[256];
  __shared__ float2 win[FFT_M];
  __shared__ float2 sm[FFT_FPB][FFT_SM];
  const int f = threadIdx.x >> 4, j = threadIdx.x & 15;
  float2* s = sm[f];
  fft_tables(tw, w256);
  for (int i = threadIdx.x; i < FFT_M; i += blockDim.x) win[i] = make_float2(window[2 * i], window[2 * i + 1]);
  __syncthreads();
  const int nframes = B * T;
  constexpr int F = FFT_M + 1;
  for (int base = blockIdx.x * FFT_FPB; base < nframes; base += gridDim.x * FFT_FPB) {
    const int fr = base + f;
    const bool valid = fr < nframes;
    const int b = valid ? fr / T : 0, t = valid ? fr - b * T : 0;
    const float* yb = y + (size_t)b * L;
    const int s0 = t * hop - FFT_N / 2;  // first sample of the frame in the unpadded signal
    float2 v[16];
#pragma unroll
    for (int n1 = 0; n1 < 16; ++n1) {
      const int n = 16 * n1 + j, p = s0 + 2 * n;
      float2 x = make_float2(0.f, 0.f);
      if (valid) {
        if (PAIRS) {  // L, hop even and y 8-byte aligned: both samples of a pair are in range or neither is
          if (p >= 0 && p < L) x = *reinterpret_cast<const float2*>(yb + p);
        } else {
          if (p >= 0 && p < L) x.x = yb[p];
          if (p + 1 >= 0 && p + 1 < L) x.y = yb[p + 1];
        }
      }
      const float2 w = win[n];
      v[n1] = make_float2(__fmul_rn(x.x, w.x), __fmul_rn(x.y, w.y));  // even samples real, odd imaginary
    }
    fft256<false>(v, w256, s, j);
#pragma unroll
    for (int k2 = 0; k2 < 16; ++k2) s[j + 16 * k2] = v[pos16(k2)];  // Z in natural order: the partner bins
    __syncwarp();
    if (valid) {
      float2* srow = spec + (size_t)fr * F;
      float* crow = cm != nullptr ? cm + ((size_t)t * B + b) * Fk : nullptr;
#pragma unroll 4  // (rolled: fully unrolled, this step made the kernel 4 096 instructions -- instruction-cache misses)
      for (int k2 = 0; k2 < 16; ++k2) {
        const int k = j + 16 * k2;
        const float2 zk = s[k], zm = s[(FFT_M - k) & (FFT_M - 1)];
        // E = (Z[k] + conj Z[M-k]) / 2 (even samples' DFT), O = -i (Z[k] - conj Z[M-k]) / 2 (odd samples' DFT)
        const float2 e = make_float2(0.5f * (zk.x + zm.x), 0.5f * (zk.y - zm.y));
        const float2 o = make_float2(0.5f * (zk.y + zm.y), -0.5f * (zk.x - zm.x));
        const float2 wo = cmul(tw[k], o);
        const float2 x = make_float2(e.x + wo.x, e.y + wo.y);
        srow[k] = x;
        if (crow != nullptr && k < Fk) crow[k] = compress_abs(x, fdrc, mode);
        if (k == 0) {  // Nyquist bin: E[0] - O[0]
          const float2 xn = make_float2(e.x - o.x, 0.f);
          srow[FFT_M] = xn;
          if (crow != nullptr && FFT_M < Fk) crow[FFT_M] = compress_abs(xn, fdrc, mode);
        }
      }
    }
    __syncwarp();  // the next frame's transpose overwrites `s`
  }
}

struct DfBands {
  const float* proj[4];  // [T, B*N, P], P = 2 * ctr * df (one speaker)
  int N[4], ctr[4], df[4], lo[4];
  int shift[4];  // log2(ctr) when ctr is a power of two, else -1
  int nb;       // bands in use
  int f_pass;   // bins >= f_pass pass through unfiltered (MSF:461-468)
  int layout;   // 0 = (c fc df s), MSF:160-167;  1 = (c df s fc), CGN:230
};

// spec [B,T,257] -> frames [B*T,512].  DF: the spectrum handed to the inverse transform is the deep-filtered one.
template <bool DF>
__global__ void __launch_bounds__(FFT_THREADS, 8) k_irfft512(const float2* __restrict__ spec, float* __restrict__ frames,
                                                          float* __restrict__ mag, float2* __restrict__ enh, int B, int T,
                                                          const __grid_constant__ DfBands bands) {
  __shared__ float2 tw[FFT_M + 1], w256[256];
  __shared__ float2 sm[FFT_FPB][FFT_SM];
  const int f = threadIdx.x >> 4, j = threadIdx.x & 15;
  float2* s = sm[f];
  fft_tables(tw, w256);
  __syncthreads();
  const int nframes = B * T;
  constexpr int F = FFT_M + 1;
  for (int base = blockIdx.x * FFT_FPB; base < nframes; base += gridDim.x * FFT_FPB) {
    const int fr = base + f;
    const bool valid = fr < nframes;
    if (valid) {
      const float2* srow = spec + (size_t)fr * F;
      float* mrow = DF && mag != nullptr ? mag + (size_t)fr * F : nullptr;
      float2* erow = DF && enh != nullptr ? enh + (size_t)fr * F : nullptr;
      int k_pass = 0;  // first bin that is copied, not filtered
      if (DF) {
        const int b = fr / T, t = fr - b * T;
        k_pass = bands.f_pass;
        for (int i = 0; i < bands.nb; ++i) {  // uniform over the frame's threads
          const int ctr = bands.ctr[i], df = bands.df[i], N = bands.N[i], lo = bands.lo[i], W = N * ctr;
          const int shift = bands.shift[i], layout = bands.layout;
          const float* __restrict__ prow = bands.proj[i] + ((size_t)t * B + b) * (size_t)N * (size_t)(2 * ctr * df);
          const float2* __restrict__ xs0 = spec + ((ptrdiff_t)fr - (df - 1)) * F + lo;  // tap d reads frame t - (df-1) + d
          const int d0 = t < df - 1 ? df - 1 - t : 0;  // taps that reach in front of the first frame are skipped
          // one bin: (n, fc) of band bin q, its df taps
          auto bin = [&](int q) {
            int n, fc;
            if (shift >= 0) { n = q >> shift; fc = q & (ctr - 1); } else { n = q / ctr; fc = q - n * ctr; }
            const float* pr = prow + (size_t)n * (2 * ctr * df);
            float yr = 0.f, yi = 0.f;
            for (int d = d0; d < df; ++d) {
              const float cr = layout == 0 ? pr[fc * df + d] : pr[d * ctr + fc];
              const float ci = layout == 0 ? pr[(ctr + fc) * df + d] : pr[(df + d) * ctr + fc];
              const float2 z = xs0[(ptrdiff_t)d * F + q];
              yr += z.x * cr - z.y * ci;
              yi += z.x * ci + z.y * cr;
            }
            return make_float2(yr, yi);
          };
          auto put = [&](int q, float2 x) {
            const int k = lo + q;
            if (mrow != nullptr) mrow[k] = cabs_fast(x);  // enh_mag, MSF:472
            if (erow != nullptr) erow[k] = x;
            if (k == 0 || k == FFT_M) x.y = 0.f;  // a C2R transform ignores the imaginary parts of DC and Nyquist
            s[k] = x;
          };
          if (df == 1 && shift >= 0) {
            // most bins of every recipe: one tap, four bins per thread at a time so that their loads are in flight together
            for (int q0 = j; q0 < W; q0 += 64) {
              float2 x[4];
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                const int q = q0 + 16 * u;
                x[u] = make_float2(0.f, 0.f);
                if (q < W) {
                  const float* pr = prow + (size_t)(q >> shift) * (2 * ctr);
                  const int fc = q & (ctr - 1);
                  const float cr = pr[fc], ci = pr[ctr + fc];
                  const float2 z = xs0[q];
                  x[u] = make_float2(z.x * cr - z.y * ci, z.x * ci + z.y * cr);
                }
              }
#pragma unroll
              for (int u = 0; u < 4; ++u)
                if (q0 + 16 * u < W) put(q0 + 16 * u, x[u]);
            }
          } else {
#pragma unroll 1
            for (int q = j; q < W; q += 16) put(q, bin(q));
          }
        }
      }
      for (int k0 = k_pass + j; k0 < F; k0 += 64) {
        float2 x[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) x[u] = k0 + 16 * u < F ? srow[k0 + 16 * u] : make_float2(0.f, 0.f);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int k = k0 + 16 * u;
          if (k < F) {
            if (mrow != nullptr) mrow[k] = cabs_fast(x[u]);
            if (erow != nullptr) erow[k] = x[u];
            if (k == 0 || k == FFT_M) x[u].y = 0.f;
            s[k] = x[u];
          }
        }
      }
    }
    __syncwarp();
    float2 v[16];
#pragma unroll
    for (int n1 = 0; n1 < 16; ++n1) {
      const int k = 16 * n1 + j;
      const float2 xk = s[k], xm = s[FFT_M - k];
      // E = (X[k] + conj X[M-k]) / 2, O = (X[k] - conj X[M-k]) / 2 * exp(+2 pi i k / 512), Z = E + i O
      const float2 e = make_float2(0.5f * (xk.x + xm.x), 0.5f * (xk.y - xm.y));
      const float2 d = make_float2(0.5f * (xk.x - xm.x), 0.5f * (xk.y + xm.y));
      const float2 o = cmulc(d, tw[k]);
      v[n1] = make_float2(e.x - o.y, e.y + o.x);
    }
    __syncwarp();  // the transpose overwrites the staged spectrum
    fft256<true>(v, w256, s, j);
    if (valid) {
      float2* frow = reinterpret_cast<float2*>(frames + (size_t)fr * FFT_N);
#pragma unroll
      for (int k2 = 0; k2 < 16; ++k2) {
        const float2 z = v[pos16(k2)];
        frow[j + 16 * k2] = make_float2(z.x * (1.0f / FFT_M), z.y * (1.0f / FFT_M));  // x[2n], x[2n+1], n = j + 16 k2
      }
    }
    __syncwarp();
  }
}

static int fft_grid(int nframes) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int want = (nframes + FFT_FPB - 1) / FFT_FPB, cap = sms * 8;  // a few resident blocks per SM, grid-stride loop
  return want < cap ? want : cap;
}

}  // namespace gsn

extern "C" int gsn_stft_compress(const float* y, const float* window, float* spec_ri, float* cm, int B, int L, int T,
                                 int n_fft, int hop, int f_keep, float fdrc, gsn_stream_t stream) {
  GSN_REQUIRE(y && window && spec_ri, "gsn_stft_compress: null pointer");
  GSN_REQUIRE(n_fft == gsn::FFT_N, "gsn_stft_compress: n_fft = %d (only the recipes' 512 is built)", n_fft);
  GSN_REQUIRE(B > 0 && L > 0 && T > 0 && hop > 0 && hop <= n_fft, "gsn_stft_compress: bad shape B=%d L=%d T=%d hop=%d", B, L,
              T, hop);
  GSN_REQUIRE(T == 1 + L / hop, "gsn_stft_compress: T must be 1 + L / hop (center=True), got T=%d L=%d hop=%d", T, L, hop);
  GSN_REQUIRE((size_t)B * T < 2147483647ULL, "gsn_stft_compress: too many frames");
  GSN_REQUIRE(cm == nullptr || (f_keep > 0 && f_keep <= n_fft / 2 + 1), "gsn_stft_compress: f_keep = %d", f_keep);
  GSN_REQUIRE((reinterpret_cast<uintptr_t>(window) & 7) == 0 && (reinterpret_cast<uintptr_t>(spec_ri) & 7) == 0,
              "gsn_stft_compress: window and spectrum must be 8-byte aligned");
  const int mode = fdrc == 0.5f ? 0 : (fdrc == 1.0f ? 1 : 2);
  const int pairs = ((L | hop) & 1) == 0 && (reinterpret_cast<uintptr_t>(y) & 7) == 0;
  if (pairs)
    gsn::k_stft512<true><<<gsn::fft_grid(B * T), gsn::FFT_THREADS, 0, gsn::as_stream(stream)>>>(
        y, window, reinterpret_cast<float2*>(spec_ri), cm, B, L, T, hop, f_keep, fdrc, mode);
  else
    gsn::k_stft512<false><<<gsn::fft_grid(B * T), gsn::FFT_THREADS, 0, gsn::as_stream(stream)>>>(
        y, window, reinterpret_cast<float2*>(spec_ri), cm, B, L, T, hop, f_keep, fdrc, mode);
  GSN_LAUNCH_CHECK("k_stft512");
  return GSN_OK;
}

extern "C" int gsn_irfft_frames(const float* spec_ri, float* frames, int B, int T, int n_fft, gsn_stream_t stream) {
  GSN_REQUIRE(spec_ri && frames, "gsn_irfft_frames: null pointer");
  GSN_REQUIRE(n_fft == gsn::FFT_N, "gsn_irfft_frames: n_fft = %d (only the recipes' 512 is built)", n_fft);
  GSN_REQUIRE(B > 0 && T > 0 && (size_t)B * T < 2147483647ULL, "gsn_irfft_frames: bad shape B=%d T=%d", B, T);
  GSN_REQUIRE(((reinterpret_cast<uintptr_t>(spec_ri) | reinterpret_cast<uintptr_t>(frames)) & 7) == 0,
              "gsn_irfft_frames: 8-byte alignment");
  gsn::DfBands none = {};
  gsn::k_irfft512<false><<<gsn::fft_grid(B * T), gsn::FFT_THREADS, 0, gsn::as_stream(stream)>>>(
      reinterpret_cast<const float2*>(spec_ri), frames, nullptr, nullptr, B, T, none);
  GSN_LAUNCH_CHECK("k_irfft512");
  return GSN_OK;
}

extern "C" int gsn_deepfilter_irfft(const float* const* projs, const int* N, const int* ctr, const int* df, int n_bands,
                                    int layout, const float* spec_ri, float* frames, float* mag_out, float* enh_ri, int B,
                                    int T, int n_fft, gsn_stream_t stream) {
  GSN_REQUIRE(projs && N && ctr && df && spec_ri && frames, "gsn_deepfilter_irfft: null pointer");
  GSN_REQUIRE(n_fft == gsn::FFT_N, "gsn_deepfilter_irfft: n_fft = %d (only the recipes' 512 is built)", n_fft);
  GSN_REQUIRE(n_bands >= 1 && n_bands <= 4, "gsn_deepfilter_irfft: %d bands (1 .. 4)", n_bands);
  GSN_REQUIRE(layout == 0 || layout == 1, "gsn_deepfilter_irfft: layout %d", layout);
  GSN_REQUIRE(B > 0 && T > 0 && (size_t)B * T < 2147483647ULL, "gsn_deepfilter_irfft: bad shape B=%d T=%d", B, T);
  GSN_REQUIRE(((reinterpret_cast<uintptr_t>(spec_ri) | reinterpret_cast<uintptr_t>(frames) |
                reinterpret_cast<uintptr_t>(enh_ri)) & 7) == 0, "gsn_deepfilter_irfft: 8-byte alignment");
  gsn::DfBands bands = {};
  int lo = 0;
  for (int i = 0; i < n_bands; ++i) {
    GSN_REQUIRE(projs[i] && N[i] > 0 && ctr[i] > 0 && df[i] > 0, "gsn_deepfilter_irfft: band %d: bad shape", i);
    bands.proj[i] = projs[i];
    bands.N[i] = N[i];
    bands.ctr[i] = ctr[i];
    bands.df[i] = df[i];
    bands.lo[i] = lo;
    bands.shift[i] = -1;
    for (int sh = 0; sh < 20; ++sh)
      if ((1 << sh) == ctr[i]) bands.shift[i] = sh;
    lo += N[i] * ctr[i];
  }
  GSN_REQUIRE(lo <= n_fft / 2 + 1, "gsn_deepfilter_irfft: the bands cover %d bins of %d", lo, n_fft / 2 + 1);
  bands.nb = n_bands;
  bands.f_pass = lo;
  bands.layout = layout;
  gsn::k_irfft512<true><<<gsn::fft_grid(B * T), gsn::FFT_THREADS, 0, gsn::as_stream(stream)>>>(
      reinterpret_cast<const float2*>(spec_ri), frames, mag_out, reinterpret_cast<float2*>(enh_ri), B, T, bands);
  GSN_LAUNCH_CHECK("k_irfft512<DF>");
  return GSN_OK;
}
