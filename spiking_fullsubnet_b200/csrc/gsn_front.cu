// Front-end / back-end glue kernels of the GSN hot path: small, coalesced, HBM-bound.
//   k_compress_mag      : MSF:434-436 (+ transpose to time-major, MSF:108)
//   k_subband_features  : MSF:241-312 gather (+ reflect pad, + tiled full-band output MSF:443) fused with the
//                         pre-LayerNorm MSF:111-112
//   k_deepfilter_band   : MSF:315-346 applied straight from the proj output layout MSF:160-167
#include <stdlib.h>

#include "gsn_common.cuh"

namespace gsn {

// mag [B,F,T] -> cm [T,B,Fk]: 32x32 smem tile transpose so both sides are coalesced.
__global__ void __launch_bounds__(256) k_compress_mag(const float* __restrict__ mag,
                                                      float* __restrict__ cm, int B, int F, int Fk,
                                                      int T, float fdrc, int mode) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int t0 = blockIdx.x * 32, f0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int f = f0 + ty + 8 * i, t = t0 + tx;
    float v = 0.f;
    if (f < Fk && t < T) {
      v = mag[((size_t)b * F + f) * T + t];
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

// one warp per (t, row); lanes stride over the K features (K <= 1024 -> <= 32 per lane).  MAXPL = features per lane
// the instantiation can hold: the small ones (K <= 64 / 256) need a third of the registers, so that several of their
// CTAs fit beside a resident 512-thread recurrence CTA (wavefront schedule); same lane partition and reduction
// order in all of them, i.e. bit-identical results.
constexpr int kMaxPerLane = 32;
template <int MAXPL>
__global__ void __launch_bounds__(256) k_subband_features(
    const float* __restrict__ cm, int f_cm, const float* __restrict__ fb, int f_fb,
    float* __restrict__ x, int T, int B, int N, int lo, int ctr, int nbr,
    const float* __restrict__ ln_w, const float* __restrict__ ln_b, float eps, TraceBuf* tb) {
  const int tslot = trace_begin(tb, 3, T, B * N, ctr + 2 * nbr + (fb ? ctr : 0));
  trace_end(tb, tslot);  // entry stamp only (warps exit independently)
  const int lane = threadIdx.x & 31;
  const int R = B * N;
  const int k_noisy = ctr + 2 * nbr;
  const int K = k_noisy + (fb ? ctr : 0);
  // grid-stride over the (t, row) pairs: the launcher may cap the grid so that this kernel cannot flood the SMs
  // the latency-critical recurrence chunks are waiting for (wavefront schedule)
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; warp < (long long)T * R;
       warp += nwarps) {
    const int t = (int)(warp / R), r = (int)(warp - (long long)t * R);
    const int b = r / N, n = r - b * N;
    const float* cm_row = cm + ((size_t)t * B + b) * f_cm;
    const float* fb_row = fb ? fb + ((size_t)t * B + b) * f_fb : nullptr;
    const int base = lo + n * ctr;
    float v[MAXPL];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < MAXPL; ++i) {
      const int j = lane + 32 * i;
      v[i] = 0.f;
      if (j < K) {
        if (j < k_noisy) {
          int q = base - nbr + j;
          q = q < 0 ? -q : q;
          q = q > f_cm - 1 ? 2 * (f_cm - 1) - q : q;
          v[i] = cm_row[q];
        } else {
          v[i] = fb_row[(base + j - k_noisy) % f_fb];
        }
        sum += v[i];
      }
      if (32 * (i + 1) >= K) break;
    }
    float* out = x + (size_t)warp * K;
    if (ln_w == nullptr) {
#pragma unroll
      for (int i = 0; i < MAXPL; ++i) {
        const int j = lane + 32 * i;
        if (j < K) out[j] = v[i];
        if (32 * (i + 1) >= K) break;
      }
      continue;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float mean = sum / (float)K;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < MAXPL; ++i) {
      const int j = lane + 32 * i;
      if (j < K) {
        const float d = v[i] - mean;
        sq += d * d;
      }
      if (32 * (i + 1) >= K) break;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    const float rstd = 1.0f / sqrtf(sq / (float)K + eps);
#pragma unroll
    for (int i = 0; i < MAXPL; ++i) {
      const int j = lane + 32 * i;
      if (j < K) out[j] = (v[i] - mean) * rstd * ln_w[j] + ln_b[j];
      if (32 * (i + 1) >= K) break;
    }
  }
}

// rowsum[t, r] = sum over the K gathered features of row r at frame t -- k_subband_features without the LayerNorm and
// WITHOUT writing x: all that surface B's laplace norms need of the gathered input (model_low_freq.py:146-171: its mean
// over an utterance; model_low_freq_count_time.py:173-204: the running mean of a row) before the streaming front end
// gathers it again on the fly.  One warp per (t, row), lanes stride over the features, shuffle reduction.
__global__ void __launch_bounds__(256) k_subband_rowsums(const float* __restrict__ cm, int f_cm,
                                                         const float* __restrict__ fb, int f_fb,
                                                         float* __restrict__ rowsum, int T, int B, int N, int lo, int ctr,
                                                         int nbr) {
  const int lane = threadIdx.x & 31;
  const int k_noisy = ctr + 2 * nbr;
  const int K = k_noisy + (fb ? ctr : 0);
  // one warp per (t, b) FRAME: its N rows one after the other (the frame's spectrum and full-band rows stay in L1), 32-bit
  // index arithmetic, no per-element modulo
  const unsigned nwarps = (gridDim.x * blockDim.x) >> 5;
  for (unsigned fr = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; fr < (unsigned)(T * B); fr += nwarps) {
    const float* cm_row = cm + (size_t)fr * f_cm;
    const float* fb_row = fb ? fb + (size_t)fr * f_fb : nullptr;
    for (int n = 0; n < N; ++n) {
      const int base = lo + n * ctr;
      const int fb0 = fb ? base % f_fb : 0;
      float sum = 0.f;
      for (int j = lane; j < K; j += 32) {
        if (j < k_noisy) {
          int q = base - nbr + j;
          q = q < 0 ? -q : q;
          q = q > f_cm - 1 ? 2 * (f_cm - 1) - q : q;
          sum += cm_row[q];
        } else {
          int q = fb0 + j - k_noisy;
          while (q >= f_fb) q -= f_fb;
          sum += fb_row[q];
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
      if (lane == 0) rowsum[(size_t)fr * N + n] = sum;  // row r = b*N + n of frame t: (t*B + b)*N + n
    }
  }
}

// thread per (b, s, n, fc, t): t fastest so spectrogram reads/writes are coalesced
__global__ void __launch_bounds__(256) k_deepfilter_band(
    const float* __restrict__ proj, const float* __restrict__ sre, const float* __restrict__ sim,
    float* __restrict__ ore, float* __restrict__ oim, int T, int B, int N, int ctr, int df, int S,
    int lo, int F, int F_out) {
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
  const float* xr = sre + ((size_t)b * F + f) * T;
  const float* xi = sim + ((size_t)b * F + f) * T;
  float yr = 0.f, yi = 0.f;
  for (int d = 0; d < df; ++d) {
    const int tt = t - (df - 1) + d;
    if (tt < 0) continue;
    const float cr = pr[((0 * ctr + fc) * df + d) * S + s];
    const float ci = pr[((1 * ctr + fc) * df + d) * S + s];
    const float a = xr[tt], bb = xi[tt];
    yr += a * cr - bb * ci;
    yi += a * ci + bb * cr;
  }
  const size_t o = (((size_t)b * S + s) * F_out + f) * T + t;
  ore[o] = yr;
  oim[o] = yi;
}

}  // namespace gsn

extern "C" int gsn_compress_mag(const float* mag, float* cm, int B, int F, int f_keep, int T,
                                float fdrc, gsn_stream_t stream) {
  GSN_REQUIRE(mag && cm, "gsn_compress_mag: null pointer");
  GSN_REQUIRE(B > 0 && F > 0 && T > 0 && f_keep > 0 && f_keep <= F,
              "gsn_compress_mag: bad shape B=%d F=%d f_keep=%d T=%d", B, F, f_keep, T);
  GSN_REQUIRE(B <= 65535, "gsn_compress_mag: B=%d > 65535", B);
  const int mode = fdrc == 0.5f ? 0 : (fdrc == 1.0f ? 1 : 2);
  dim3 grid((T + 31) / 32, (f_keep + 31) / 32, B);
  gsn::k_compress_mag<<<grid, 256, 0, gsn::as_stream(stream)>>>(mag, cm, B, F, f_keep, T, fdrc, mode);
  GSN_LAUNCH_CHECK("k_compress_mag");
  return GSN_OK;
}

extern "C" int gsn_subband_features(const float* cm, int f_cm, const float* fb, int f_fb, float* x,
                                    int T, int B, int N, int lo, int ctr, int nbr,
                                    const float* ln_weight, const float* ln_bias, float ln_eps,
                                    gsn_stream_t stream) {
  GSN_REQUIRE(cm && x, "gsn_subband_features: null pointer");
  GSN_REQUIRE(T > 0 && B > 0 && N > 0 && ctr > 0 && nbr >= 0 && lo >= 0,
              "gsn_subband_features: bad shape");
  const int K = ctr + 2 * nbr + (fb ? ctr : 0);
  GSN_REQUIRE(K <= 32 * gsn::kMaxPerLane, "gsn_subband_features: K=%d > %d", K, 32 * gsn::kMaxPerLane);
  GSN_REQUIRE(lo + N * ctr <= f_cm, "gsn_subband_features: band [%d,%d) leaves the %d-bin spectrum",
              lo, lo + N * ctr, f_cm);
  GSN_REQUIRE(nbr < f_cm, "gsn_subband_features: nbr=%d too large", nbr);
  // interior bands must see real neighbours; only the spectrum edges reflect (MSF:290-302)
  GSN_REQUIRE(lo == 0 || lo - nbr >= 0, "gsn_subband_features: lower neighbourhood out of range");
  GSN_REQUIRE(lo + N * ctr == f_cm || lo + N * ctr + nbr <= f_cm,
              "gsn_subband_features: upper neighbourhood out of range");
  GSN_REQUIRE(!fb || f_fb > 0, "gsn_subband_features: f_fb");
  GSN_REQUIRE((ln_weight == nullptr) == (ln_bias == nullptr), "gsn_subband_features: ln params");
  const long long warps = (long long)T * B * N;
  long long blocks = (warps + 7) / 8;
  GSN_REQUIRE(blocks < 2147483647LL, "gsn_subband_features: too many rows");
  const int cap = gsn::launch_option(GSN_OPT_F32_MAX_CTAS);
  if (cap > 0 && blocks > cap) blocks = cap;
  auto launch = [&](auto kern) {
    kern<<<(unsigned)blocks, 256, 0, gsn::as_stream(stream)>>>(cm, f_cm, fb, f_fb, x, T, B, N, lo, ctr, nbr, ln_weight,
                                                              ln_bias, ln_eps, gsn::trace_buffer());
  };
  if (K <= 64) launch(gsn::k_subband_features<2>);
  else if (K <= 256) launch(gsn::k_subband_features<8>);
  else launch(gsn::k_subband_features<gsn::kMaxPerLane>);
  GSN_LAUNCH_CHECK("k_subband_features");
  return GSN_OK;
}

extern "C" int gsn_subband_rowsums(const float* cm, int f_cm, const float* fb, int f_fb, float* rowsum, int T, int B,
                                   int N, int lo, int ctr, int nbr, gsn_stream_t stream) {
  GSN_REQUIRE(cm && rowsum, "gsn_subband_rowsums: null pointer");
  GSN_REQUIRE(T > 0 && B > 0 && N > 0 && ctr > 0 && nbr >= 0 && lo >= 0, "gsn_subband_rowsums: bad shape");
  GSN_REQUIRE(lo + N * ctr <= f_cm, "gsn_subband_rowsums: band [%d,%d) leaves the %d-bin spectrum", lo, lo + N * ctr,
              f_cm);
  GSN_REQUIRE(nbr < f_cm, "gsn_subband_rowsums: nbr=%d too large", nbr);
  GSN_REQUIRE(lo == 0 || lo - nbr >= 0, "gsn_subband_rowsums: lower neighbourhood out of range");
  GSN_REQUIRE(lo + N * ctr == f_cm || lo + N * ctr + nbr <= f_cm, "gsn_subband_rowsums: upper neighbourhood out of range");
  GSN_REQUIRE(!fb || f_fb > 0, "gsn_subband_rowsums: f_fb");
  GSN_REQUIRE((long long)T * B < 2147483647LL, "gsn_subband_rowsums: too many frames");
  long long blocks = ((long long)T * B + 7) / 8;  // a warp per (t, b) frame
  if (blocks > 148 * 16) blocks = 148 * 16;  // grid-stride
  gsn::k_subband_rowsums<<<(unsigned)blocks, 256, 0, gsn::as_stream(stream)>>>(cm, f_cm, fb, f_fb, rowsum, T, B, N, lo, ctr,
                                                                               nbr);
  GSN_LAUNCH_CHECK("k_subband_rowsums");
  return GSN_OK;
}

extern "C" int gsn_deepfilter_band(const float* proj, const float* spec_re, const float* spec_im,
                                   float* out_re, float* out_im, int T, int B, int N, int ctr, int df,
                                   int S, int lo, int F, int F_out, gsn_stream_t stream) {
  GSN_REQUIRE(proj && spec_re && spec_im && out_re && out_im, "gsn_deepfilter_band: null pointer");
  GSN_REQUIRE(T > 0 && B > 0 && N > 0 && ctr > 0 && df > 0 && S > 0, "gsn_deepfilter_band: bad shape");
  GSN_REQUIRE(lo + N * ctr <= F && lo + N * ctr <= F_out, "gsn_deepfilter_band: band out of range");
  const size_t total = (size_t)B * S * N * ctr * T;
  const size_t blocks = (total + 255) / 256;
  GSN_REQUIRE(blocks < 2147483647ULL, "gsn_deepfilter_band: too large");
  gsn::k_deepfilter_band<<<(unsigned)blocks, 256, 0, gsn::as_stream(stream)>>>(
      proj, spec_re, spec_im, out_re, out_im, T, B, N, ctr, df, S, lo, F, F_out);
  GSN_LAUNCH_CHECK("k_deepfilter_band");
  return GSN_OK;
}
