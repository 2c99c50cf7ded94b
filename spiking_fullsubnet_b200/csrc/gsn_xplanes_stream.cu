// Streaming front end of one sequence model for the FUSED layer-0 recurrence: sub-band gather (MSF:241-312, SURVEY
// App. B) + pre-LayerNorm (MSF:111-112) + the split of the normalised input into three exact bf16 planes
// (x = hi + mid + lo, truncation), written to global memory directly in the tcgen05 B-operand image the layer-0
// recurrence (gsn_recurrence_stream, input mode "planes") consumes: one contiguous block
//     xop[t][tile][plane 0..2 = lo, mid, hi][NT rows x Kmma, K-major core matrices (8 rows x 16 bytes)]
// per frame and row tile, fetched by that kernel with ONE bulk copy per frame.  The input-to-hidden product
// x_t . W_ih^T (ESN:141) itself runs on the tensor cores of the recurrence CTAs (W_ih planes in their tensor memory,
// 8 of the 9 plane pairs, off the critical path), so the layer-0 xproj never exists in HBM.
//
// CUDA cores only (no tensor memory): 16 worker warps + one publisher warp per CTA, persistent.  A work unit is 8 rows
// of one frame; units are dealt frame-major over all worker warps of the grid, so a frame is finished by the whole grid
// at once.  Per unit: (1) the raw features of the 8 rows are loaded with lanes running over the features (coalesced)
// into a per-warp scratch, after an acquire poll of the full-band model's frame counter (the sub-band input of frame t
// needs the full-band output of frame t, MSF:441-447); (2) each lane reads back (row, 8-feature chunks) as 16-byte
// words, does the LayerNorm (two shuffles per reduction), the split, and three 16-byte stores per chunk (8 lanes fill
// one 128-byte core matrix).  The publisher issues one gpu-scope release per round of units and adds the rows done to
// out_cnt[t] (frame complete at R).
#include <stdlib.h>

#include "gsn_common.cuh"
#include "gsn_tc.cuh"

namespace gsn {

int preload_recurrence_stream();  // gsn_recurrence_stream.cu
int preload_stage_stream();       // gsn_stage_stream.cu

struct XpParams {
  const float* cm;      // [T, B, f_cm] compressed magnitude
  const float* fb;      // [T, B, f_fb] full-band output or null
  const float* ln_w;    // [K] or null
  const float* ln_b;
  const float* row_div; // or null: divisor of every feature of a row (surface B's laplace norms, model_low_freq.py:146-204)
  int div_mode;         // 1: row_div[b] per utterance (offline norm);  2: row_div[t * R + r] (cumulative norm)
  float* x_out;         // [T, R, K] normalised input (all_layer_outputs[0]) or null
  uint8_t* xop;         // operand images, see above
  const unsigned int* in_cnt;  // [T] or null
  unsigned int in_target;
  unsigned int* out_cnt;       // [T] or null: += rows
  const unsigned int* bp_cnt;  // [T] or null: the consumer's frame counters (back-pressure of the ring)
  unsigned int bp_target;
  int ring;                    // frames the operand-image buffer holds (slot = t % ring); >= T: no reuse
  int T, B, N, lo, ctr, nbr, f_cm, f_fb, K, Kmma, R, nt, pitch;
  float eps;
  unsigned int poll_ns;
  TraceBuf* trace;
};

constexpr int kXpWorkers = 16;
constexpr int kXpThreads = (kXpWorkers + 1) * 32;
constexpr int kXpRing = 4;

__device__ __forceinline__ uint32_t xp_ld_cg(const void* p) {
  uint32_t v;
  asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ bool xp_poll_frames(const unsigned int* cnt, unsigned int target, int T, int& ready, int need,
                                               int lane, unsigned int ns) {
  unsigned long long t0 = 0;
  for (unsigned int spins = 0;; ++spins) {
    const int t = ready + lane;
    bool ok = false;
    if (t < T) {
      unsigned int v;
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(cnt + t) : "memory");
      ok = v >= target;
    }
    const unsigned int m = __ballot_sync(0xffffffffu, ok);
    ready += m == 0xffffffffu ? 32 : __ffs(~m) - 1;
    if (ready > need) return true;
    if ((spins & 0x3FFu) == 0x3FFu) {  // wall-clock bound: the producer kernel may start late (lazy module loading)
      const unsigned long long now = tc::wait_clock_ns();
      if (t0 == 0) t0 = now;
      else if (now - t0 > tc::kWaitTimeoutNs) return false;
    }
    __nanosleep(ns);
  }
}

template <int J>
__global__ void __launch_bounds__(kXpThreads, 1) k_xplanes_stream(const XpParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int tslot = trace_begin(p.trace, 7, p.T, p.R, p.K);
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
  const int K = p.K, Kmma = p.Kmma, R = p.R, NT = p.nt;
  float* s_lnw = reinterpret_cast<float*>(smem);
  float* s_lnb = s_lnw + Kmma;
  float* s_scr = s_lnb + Kmma;  // [workers][8][pitch]
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_scr + (size_t)kXpWorkers * 8 * p.pitch);
  uint64_t* bar_pub = bars;            // [4] round stored by the worker warps
  uint64_t* bar_pfree = bars + kXpRing;  // [4] round published
  if (tid == 0) {
    for (int i = 0; i < kXpRing; ++i) {
      tc::mbar_init(&bar_pub[i], kXpWorkers);
      tc::mbar_init(&bar_pfree[i], 1);
    }
    tc::fence_mbar_init();
  }
  for (int k = tid; k < Kmma; k += kXpThreads) {
    s_lnw[k] = (p.ln_w != nullptr && k < K) ? p.ln_w[k] : 1.f;
    s_lnb[k] = (p.ln_w != nullptr && k < K) ? p.ln_b[k] : 0.f;
  }
  __syncthreads();

  const int G = (R + 7) / 8;                     // units per frame
  const long long total = (long long)p.T * G;
  const int per_round = gridDim.x * kXpWorkers;  // units per round of the whole grid
  const long long rounds = (total + per_round - 1) / per_round;
  const int ntiles = (R + NT - 1) / NT;
  const uint32_t SBO = 16u * Kmma;
  const size_t plane_bytes = (size_t)NT * Kmma * 2;
  const bool do_pub = p.out_cnt != nullptr;

  if (warp == kXpWorkers) {
    // =============================== publisher warp ===============================
    if (lane == 0 && do_pub) {
      for (long long k = 0; k < rounds; ++k) {
        const int ps = (int)(k % kXpRing);
        if (!tc::mbar_wait_cta(&bar_pub[ps], (uint32_t)((k / kXpRing) & 1))) __trap();
        asm volatile("fence.acq_rel.gpu;" ::: "memory");  // the workers' stores, observed through the mbarrier
        long long u = k * per_round + (long long)blockIdx.x * kXpWorkers;
        long long ue = u + kXpWorkers < total ? u + kXpWorkers : total;
        while (u < ue) {  // units [u, ue) of this CTA: rows done per touched frame
          const int t = (int)(u / G), g0 = (int)(u - (long long)t * G);
          const long long fe = (long long)(t + 1) * G < ue ? (long long)(t + 1) * G : ue;
          const int g1 = (int)(fe - (long long)t * G);  // groups [g0, g1) of frame t
          const int rows = (g1 * 8 < R ? g1 * 8 : R) - g0 * 8;
          asm volatile("red.relaxed.gpu.global.add.u32 [%0], %1;" ::"l"(p.out_cnt + t), "r"((unsigned int)rows) : "memory");
          u = fe;
        }
        tc::mbar_arrive(&bar_pfree[ps]);
      }
    }
    __syncwarp();
  } else {
    // =============================== worker warps ===============================
    const int rg = lane & 7, cg = lane >> 3;  // row of the 8-row unit / chunk group
    const int k_noisy = p.ctr + 2 * p.nbr;
    const int k8n = Kmma / 8;
    const bool use_ln = p.ln_w != nullptr;
    const float inv_k = 1.0f / (float)K;
    float* scr = s_scr + (size_t)warp * 8 * p.pitch;
    // per-lane constants of feature jx = lane + 32 i: from the noisy band or the full-band output, inside K, and
    // (jx - k_noisy) mod f_fb
    bool g_noisy[J], g_valid[J];
    int g_off[J];
    const float* g_base[J];
#pragma unroll
    for (int i = 0; i < J; ++i) {
      const int jx = lane + 32 * i;
      g_noisy[i] = jx < k_noisy;
      g_valid[i] = jx < K;
      g_off[i] = (!g_noisy[i] && p.f_fb > 0) ? (jx - k_noisy) % p.f_fb : 0;
      g_base[i] = (g_noisy[i] || p.fb == nullptr) ? p.cm : p.fb;
    }
    const int lo_mod = p.f_fb > 0 ? p.lo % p.f_fb : 0;
    int ready = 0;     // frames [0, ready) of the input are known complete
    int ready_bp = 0;  // frames [0, ready_bp) have been consumed downstream (their ring slots are free)
    for (long long k = 0; k < rounds; ++k) {
      const long long u = k * per_round + (long long)blockIdx.x * kXpWorkers + warp;
      if (u < total) {
        const int t = (int)(u / G), g = (int)(u - (long long)t * G);
        const int r0 = g * 8;
        if (p.in_cnt != nullptr && ready <= t && !xp_poll_frames(p.in_cnt, p.in_target, p.T, ready, t, lane, p.poll_ns)) __trap();
        // ring reuse: slot t % ring still holds frame t - ring until the consumer has finished that frame
        if (p.bp_cnt != nullptr && t >= p.ring && ready_bp <= t - p.ring &&
            !xp_poll_frames(p.bp_cnt, p.bp_target, p.T, ready_bp, t - p.ring, lane, p.poll_ns)) __trap();
        {
          // (1) raw features, lanes over the features; rows walk (b, ns) without a division per row
          int rb = r0 / p.N, rn = r0 - rb * p.N;
          int bmod = p.f_fb > 0 ? (p.lo + rn * p.ctr) % p.f_fb : 0;  // (lo + ns*ctr) mod f_fb; ctr <= f_fb
          float raw[8][J];
#pragma unroll
          for (int u8 = 0; u8 < 8; ++u8) {
            const bool rvu = r0 + u8 < R;
            const int tb = t * p.B + (rvu ? rb : 0);
            const int row_cm = tb * p.f_cm, row_fb = tb * p.f_fb;
            const int q0 = p.lo + rn * p.ctr - p.nbr + lane;
#pragma unroll
            for (int i = 0; i < J; ++i) {
              int qq = q0 + 32 * i;  // reflect padding at both ends of the spectrum (MSF:262)
              qq = qq < 0 ? -qq : qq;
              qq = min(qq, 2 * (p.f_cm - 1) - qq);
              int fi = bmod + g_off[i];
              fi = fi >= p.f_fb ? fi - p.f_fb : fi;
              const int off = g_noisy[i] ? row_cm + qq : row_fb + fi;
              raw[u8][i] = (rvu && g_valid[i]) ? __uint_as_float(xp_ld_cg(g_base[i] + off)) : 0.f;
            }
            ++rn;
            bmod += p.ctr;
            bmod = bmod >= p.f_fb ? bmod - p.f_fb : bmod;
            if (rn == p.N) { rn = 0; ++rb; bmod = lo_mod; }
          }
#pragma unroll
          for (int u8 = 0; u8 < 8; ++u8)
#pragma unroll
            for (int i = 0; i < J; ++i)
              if (lane + 32 * i < Kmma) scr[u8 * p.pitch + lane + 32 * i] = raw[u8][i];
        }
        __syncwarp();
        // (2) lane = (row rg, chunks cg, cg+4, ... of 8 features)
        const int r = r0 + rg;
        const bool rv = r < R;
        float v[J][8];
#pragma unroll
        for (int jc = 0; jc < J; ++jc) {
          const int ch = cg + 4 * jc;
          float4 x0 = make_float4(0.f, 0.f, 0.f, 0.f), x1 = x0;
          if (ch < k8n) {
            x0 = *reinterpret_cast<const float4*>(scr + rg * p.pitch + ch * 8);
            x1 = *reinterpret_cast<const float4*>(scr + rg * p.pitch + ch * 8 + 4);
          }
          v[jc][0] = x0.x; v[jc][1] = x0.y; v[jc][2] = x0.z; v[jc][3] = x0.w;
          v[jc][4] = x1.x; v[jc][5] = x1.y; v[jc][6] = x1.z; v[jc][7] = x1.w;
        }
        __syncwarp();  // the scratch may be overwritten by the next unit
        if (use_ln) {  // two-pass moments over the K features of the row (MSF:111-112, torch.nn.LayerNorm)
          float sum = 0.f;
#pragma unroll
          for (int jc = 0; jc < J; ++jc)
#pragma unroll
            for (int e = 0; e < 8; ++e) sum += v[jc][e];
          sum += __shfl_xor_sync(0xffffffffu, sum, 8);
          sum += __shfl_xor_sync(0xffffffffu, sum, 16);
          const float mean = sum * inv_k;
          float sq = 0.f;
#pragma unroll
          for (int jc = 0; jc < J; ++jc)
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const float d = ((cg + 4 * jc) * 8 + e < K) ? v[jc][e] - mean : 0.f;
              sq = fmaf(d, d, sq);
            }
          sq += __shfl_xor_sync(0xffffffffu, sq, 8);
          sq += __shfl_xor_sync(0xffffffffu, sq, 16);
          const float rstd = rsqrtf(fmaf(sq, inv_k, p.eps));
#pragma unroll
          for (int jc = 0; jc < J; ++jc) {
            const int ch = cg + 4 * jc;
            if (ch < k8n) {
              const float4 w0 = *reinterpret_cast<const float4*>(s_lnw + ch * 8);
              const float4 w1 = *reinterpret_cast<const float4*>(s_lnw + ch * 8 + 4);
              const float4 b0 = *reinterpret_cast<const float4*>(s_lnb + ch * 8);
              const float4 b1 = *reinterpret_cast<const float4*>(s_lnb + ch * 8 + 4);
              const float ww[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
              const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
              for (int e = 0; e < 8; ++e)
                v[jc][e] = (rv && ch * 8 + e < K) ? fmaf((v[jc][e] - mean) * rstd, ww[e], bb[e]) : 0.f;
            }
          }
        }
        if (p.row_div != nullptr) {  // x / (mu + eps), the division the reference performs (not a reciprocal multiply)
          const float dv = rv ? (p.div_mode == 1 ? p.row_div[r / p.N] : p.row_div[(size_t)t * R + r]) : 1.f;
#pragma unroll
          for (int jc = 0; jc < J; ++jc)
#pragma unroll
            for (int e = 0; e < 8; ++e) v[jc][e] = __fdiv_rn(v[jc][e], dv);
        }
        float* xo = (p.x_out != nullptr && rv) ? p.x_out + ((size_t)t * R + r) * K : nullptr;
        const int tile = r / NT, n = r - tile * NT;  // row tile of the recurrence and the row inside it
        uint8_t* blk = p.xop + ((size_t)(t % p.ring) * ntiles + tile) * 3 * plane_bytes + (size_t)(n >> 3) * SBO + (n & 7) * 16;
#pragma unroll
        for (int jc = 0; jc < J; ++jc) {
          const int ch = cg + 4 * jc;
          if (ch < k8n) {
            if (xo != nullptr)
#pragma unroll
              for (int e = 0; e < 8; ++e)
                if (ch * 8 + e < K) xo[ch * 8 + e] = v[jc][e];
            // truncation split of 8 values into three bf16 planes, two values per 32-bit word (PRMT packs the halves)
            uint32_t wh[4], wm[4], wl[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const uint32_t a0 = __float_as_uint(v[jc][2 * q]), a1 = __float_as_uint(v[jc][2 * q + 1]);
              wh[q] = __byte_perm(a0, a1, 0x7632);
              const float e0 = v[jc][2 * q] - __uint_as_float(a0 & 0xFFFF0000u);
              const float e1 = v[jc][2 * q + 1] - __uint_as_float(a1 & 0xFFFF0000u);
              const uint32_t c0 = __float_as_uint(e0), c1 = __float_as_uint(e1);
              wm[q] = __byte_perm(c0, c1, 0x7632);
              const float f0 = e0 - __uint_as_float(c0 & 0xFFFF0000u);
              const float f1 = e1 - __uint_as_float(c1 & 0xFFFF0000u);
              wl[q] = __byte_perm(__float_as_uint(f0), __float_as_uint(f1), 0x7632);
            }
            uint8_t* d0 = blk + (size_t)ch * 128;  // rows past R of the unit are written as zeros
            *reinterpret_cast<uint4*>(d0) = make_uint4(wl[0], wl[1], wl[2], wl[3]);
            *reinterpret_cast<uint4*>(d0 + plane_bytes) = make_uint4(wm[0], wm[1], wm[2], wm[3]);
            *reinterpret_cast<uint4*>(d0 + 2 * plane_bytes) = make_uint4(wh[0], wh[1], wh[2], wh[3]);
          }
        }
      }
      if (do_pub) {
        __syncwarp();
        if (lane == 0) {
          const int ps = (int)(k % kXpRing);
          if (k >= kXpRing && !tc::mbar_wait_cta(&bar_pfree[ps], (uint32_t)(((k / kXpRing) - 1) & 1))) __trap();
          tc::mbar_arrive(&bar_pub[ps]);
        }
      }
    }
  }
  trace_end(p.trace, tslot);
}

template <int J>
static int launch_xplanes(XpParams p, int ctas, cudaStream_t st) {
  p.pitch = p.Kmma + 4;  // Kmma % 16 == 0, so pitch % 8 == 4: conflict-free 16-byte reads by 8 consecutive rows
  static const unsigned int poll_ns = getenv("GSN_POLL_NS") ? (unsigned int)atoi(getenv("GSN_POLL_NS")) : 100u;
  p.poll_ns = poll_ns;
  const size_t smem = ((size_t)2 * p.Kmma + (size_t)kXpWorkers * 8 * p.pitch) * 4 + 2 * kXpRing * 8 + 64;
  if (smem > tc::kMaxDynamicSmem) return fail(GSN_ENOSUP, "gsn_xplanes_stream: K=%d does not fit shared memory", p.K);
  GSN_CUDA(cudaFuncSetAttribute(k_xplanes_stream<J>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long units = (long long)p.T * ((p.R + 7) / 8);
  long long n = ctas < 1 ? 1 : ctas;
  if (n * kXpWorkers > units) n = (units + kXpWorkers - 1) / kXpWorkers;
  k_xplanes_stream<J><<<(unsigned)n, kXpThreads, smem, st>>>(p);
  GSN_LAUNCH_CHECK("k_xplanes_stream");
  return GSN_OK;
}

int preload_xplanes_stream() {
  cudaFuncAttributes a;
  GSN_CUDA(cudaFuncGetAttributes(&a, k_xplanes_stream<2>));
  GSN_CUDA(cudaFuncGetAttributes(&a, k_xplanes_stream<3>));
  GSN_CUDA(cudaFuncGetAttributes(&a, k_xplanes_stream<5>));
  GSN_CUDA(cudaFuncGetAttributes(&a, k_xplanes_stream<8>));
  return GSN_OK;
}

}  // namespace gsn

// Kernels of the streaming pipeline spin on counters their producers advance.  With lazy module loading the FIRST
// launch of a kernel loads it, which synchronises with running kernels -- a consumer already spinning would then wait
// for a producer that cannot be loaded.  Call once per process and device before the first pipeline launch.
extern "C" int gsn_stream_preload(void) {
  int rc = gsn::preload_recurrence_stream();
  if (rc == GSN_OK) rc = gsn::preload_stage_stream();
  if (rc == GSN_OK) rc = gsn::preload_xplanes_stream();
  return rc;
}

extern "C" size_t gsn_xplanes_bytes(int T, int R, int K, int nt) {
  if (T <= 0 || R <= 0 || K <= 0 || (nt != 16 && nt != 32 && nt != 64)) return 0;
  const size_t Kmma = (size_t)(K + 15) / 16 * 16;
  return (size_t)T * ((R + nt - 1) / nt) * 3 * nt * Kmma * 2;
}

extern "C" int gsn_xplanes_stream(const float* cm, int f_cm, const float* fb, int f_fb, const float* ln_weight,
                                  const float* ln_bias, float ln_eps, const float* row_div, int div_mode, float* x_out,
                                  void* xop, int ring,
                                  const unsigned int* in_cnt, unsigned int in_target, unsigned int* out_cnt,
                                  const unsigned int* bp_cnt, unsigned int bp_target, int T, int B, int N, int lo,
                                  int ctr, int nbr, int nt, int ctas, gsn_stream_t stream) {
  using namespace gsn;
  GSN_REQUIRE(cm && xop, "gsn_xplanes_stream: null pointer");
  GSN_REQUIRE(T > 0 && B > 0 && N > 0 && ctr > 0 && nbr >= 0 && lo >= 0, "gsn_xplanes_stream: bad shape");
  GSN_REQUIRE(nt == 16 || nt == 32 || nt == 64, "gsn_xplanes_stream: nt=%d must be 16, 32 or 64", nt);
  const int K = ctr + 2 * nbr + (fb ? ctr : 0);
  GSN_REQUIRE(K <= 256, "gsn_xplanes_stream: K=%d not supported (K <= 256)", K);
  GSN_REQUIRE((reinterpret_cast<uintptr_t>(xop) & 127) == 0, "gsn_xplanes_stream: xop must be 128-byte aligned");
  GSN_REQUIRE(lo + N * ctr <= f_cm, "gsn_xplanes_stream: band leaves the spectrum");
  GSN_REQUIRE(lo == 0 || lo - nbr >= 0, "gsn_xplanes_stream: lower neighbourhood out of range");
  GSN_REQUIRE(lo + N * ctr == f_cm || lo + N * ctr + nbr <= f_cm, "gsn_xplanes_stream: upper neighbourhood out of range");
  GSN_REQUIRE(!fb || (f_fb > 0 && ctr <= f_fb), "gsn_xplanes_stream: f_fb=%d must be >= ctr=%d", f_fb, ctr);
  GSN_REQUIRE((long long)T * B * (f_cm > f_fb ? f_cm : f_fb) < (1ll << 31), "gsn_xplanes_stream: inputs too large");
  GSN_REQUIRE((ln_weight == nullptr) == (ln_bias == nullptr), "gsn_xplanes_stream: ln params");
  GSN_REQUIRE(row_div == nullptr || div_mode == 1 || div_mode == 2, "gsn_xplanes_stream: div_mode %d", div_mode);
  if (ring <= 0 || ring > T) ring = T;
  GSN_REQUIRE(ring == T || bp_cnt != nullptr, "gsn_xplanes_stream: a ring shorter than T needs the consumer's counters");
  XpParams p{};
  p.cm = cm; p.fb = fb; p.ln_w = ln_weight; p.ln_b = ln_bias; p.x_out = x_out; p.xop = static_cast<uint8_t*>(xop);
  p.row_div = row_div; p.div_mode = div_mode;
  p.in_cnt = in_cnt; p.in_target = in_target; p.out_cnt = out_cnt;
  p.bp_cnt = ring < T ? bp_cnt : nullptr; p.bp_target = bp_target; p.ring = ring;
  p.T = T; p.B = B; p.N = N; p.lo = lo; p.ctr = ctr; p.nbr = nbr; p.f_cm = f_cm; p.f_fb = f_fb;
  p.K = K; p.Kmma = (K + 15) / 16 * 16; p.R = B * N; p.nt = nt; p.eps = ln_eps; p.trace = trace_buffer();
  cudaStream_t st = as_stream(stream);
  // the ring must hold every frame the grid can have in flight (two rounds of units)
  const int G = (p.R + 7) / 8;
  const long long n_ctas = ctas < 1 ? 1 : ctas;
  GSN_REQUIRE(ring == T || (long long)ring * G >= 2 * n_ctas * kXpWorkers + G,
              "gsn_xplanes_stream: ring=%d frames is too short for %lld CTAs", ring, n_ctas);
  if (p.Kmma <= 64) return launch_xplanes<2>(p, ctas, st);
  if (p.Kmma <= 96) return launch_xplanes<3>(p, ctas, st);
  if (p.Kmma <= 160) return launch_xplanes<5>(p, ctas, st);
  return launch_xplanes<8>(p, ctas, st);
}
