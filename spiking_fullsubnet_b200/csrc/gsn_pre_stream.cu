// Streaming front end of one sequence model on tcgen05: sub-band gather (MSF:241-312) + pre-LayerNorm
// (MSF:111-112) + the layer-0 input-to-hidden product x_t . W_ih^T (ESN:141) in ONE persistent kernel that follows the
// full-band model frame by frame through per-frame counters (the sub-band input of frame t needs the full-band
// output of frame t, MSF:441-447) and feeds the layer-0 streaming recurrence the same way.
//
// The operand x is REAL valued, so the fp32 product is emulated exactly enough on the tensor cores: both operands are
// split by truncation into three bf16 planes (w = w_hi + w_mid + w_lo, x = x_hi + x_mid + x_lo, exact), every
// bf16 x bf16 product is exact in fp32, and 8 of the 9 plane pairs are accumulated in fp32 in tensor memory,
// smallest magnitude first (only lo x lo, <= 2^-32 of |w||x|, is dropped) -- an fp32-faithful dot product whose error is
// below that of an fp32 FMA chain.  A = the three weight planes, stationary in tensor memory (as in
// gsn_linear_tc.cu); B = the three planes of 64 (or 32) gathered + normalised rows in shared memory, double buffered.
#include "gsn_common.cuh"
#include "gsn_tc.cuh"

namespace gsn {

struct PreParams {
  const float* cm;      // [T, B, f_cm] compressed magnitude
  const float* fb;      // [T, B, f_fb] full-band output or null
  const float* ln_w;    // [K] or null
  const float* ln_b;
  const float* w_ih;    // [H, K]
  float* x_out;         // [T, R, K] normalised input (all_layer_outputs[0]) or null
  float* xproj;         // [T, R, H]
  const unsigned int* in_cnt;  // [T] or null (counts of the full-band proj kernel)
  unsigned int in_target;
  unsigned int* out_cnt;       // [T] or null: += rows per (tile, slice); frame complete at R * slices
  int T, B, N, lo, ctr, nbr, f_cm, f_fb, K, Kmma, H;
  float eps;
  int wpitch;
  TraceBuf* trace;
};

constexpr int kPreThreads = 512;
constexpr int kPreMaxPL = 9;  // K <= 288

__device__ __forceinline__ uint32_t pre_ld_cg(const float* p) {
  uint32_t v;
  asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// 8 plane pairs x KS k steps, smallest terms first; w plane pw at tmem_a + (pw*KS + ks)*8, x plane px at
// desc_b0 + px*PB + ks*16 (PB = plane bytes >> 4).  Plane index: 0 = lo, 1 = mid, 2 = hi.
template <int KS, int PB>
__device__ __forceinline__ void mma_pairs_unrolled(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b0, uint32_t idesc) {
  constexpr int PW[8] = {0, 1, 0, 2, 1, 1, 2, 2};
  constexpr int PX[8] = {1, 0, 2, 0, 1, 2, 1, 2};
  tc::mma_ts_c<0>(tmem_d, tmem_a + (uint32_t)((PW[0] * KS) * 8), desc_b0 + (uint64_t)(PX[0] * PB), idesc);
#pragma unroll
  for (int i = 1; i < 8 * KS; ++i) {
    const int term = i / KS, ks = i % KS;
    tc::mma_ts_c<1>(tmem_d, tmem_a + (uint32_t)((PW[term] * KS + ks) * 8),
                    desc_b0 + (uint64_t)(PX[term] * PB + ks * 16), idesc);
  }
}
template <int NT>
__device__ __forceinline__ bool mma_pairs(int ksteps, uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b0, uint32_t idesc) {
  switch (ksteps) {
#define GSN_KS_CASE(n) case n: mma_pairs_unrolled<n, NT * n * 2>(tmem_d, tmem_a, desc_b0, idesc); return true;
    GSN_KS_CASE(1) GSN_KS_CASE(2) GSN_KS_CASE(3) GSN_KS_CASE(4) GSN_KS_CASE(5) GSN_KS_CASE(6) GSN_KS_CASE(7)
    GSN_KS_CASE(8) GSN_KS_CASE(9) GSN_KS_CASE(10) GSN_KS_CASE(11) GSN_KS_CASE(12) GSN_KS_CASE(13) GSN_KS_CASE(14)
    GSN_KS_CASE(15) GSN_KS_CASE(16) GSN_KS_CASE(17) GSN_KS_CASE(18)
#undef GSN_KS_CASE
    default: return false;
  }
}

__device__ __forceinline__ void pre_split3(float w, uint32_t& hi, uint32_t& mid, uint32_t& lo) {
  const uint32_t wb = __float_as_uint(w);
  hi = wb >> 16;
  const float r1 = w - __uint_as_float(wb & 0xFFFF0000u);
  const uint32_t r1b = __float_as_uint(r1);
  mid = r1b >> 16;
  const float r2 = r1 - __uint_as_float(r1b & 0xFFFF0000u);
  lo = __float_as_uint(r2) >> 16;
}

static inline __host__ __device__ int pre_wpitch(int K) { return ((K / 4) & 1) ? K : K + 4; }

template <int NT>
__global__ void __launch_bounds__(kPreThreads, 1) k_pre_stream(const PreParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int tslot = trace_begin(p.trace, 6, p.T, p.B * p.N, p.K);
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
  const int q = warp & 3, g = warp >> 2;  // TMEM lane quarter / row group (4 groups of NT/4 rows)
  const int slice = blockIdx.x, P = gridDim.y;
  const int K = p.K, Kmma = p.Kmma, H = p.H, R = p.B * p.N;
  const long long M = (long long)p.T * R;
  const long long ntiles_all = (M + NT - 1) / NT;
  const int j = slice * 128 + q * 32 + lane;  // output feature of this thread (TMEM lane)
  const bool jv = j < H;

  const uint32_t SBO = 16u * Kmma;
  const size_t plane_bytes = (size_t)NT * Kmma * 2;           // multiple of 512
  const size_t buf_bytes = 3 * plane_bytes;
  const size_t stage_bytes = p.wpitch > 0 ? (size_t)128 * p.wpitch * 4 : 0;
  const size_t bar_off = ((2 * buf_bytes > stage_bytes ? 2 * buf_bytes : stage_bytes) + 127) / 128 * 128;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + bar_off);  // [2] MMA completion, [1] weight staging
  uint64_t* bar_w = bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 3);
  float* wst = reinterpret_cast<float*>(smem);
  const float* wrow = p.w_ih + (size_t)(jv ? j : 0) * K;

  if (tid == 0) {
    tc::mbar_init(&bar[0], 1);
    tc::mbar_init(&bar[1], 1);
    tc::mbar_init(bar_w, 1);
    tc::fence_mbar_init();
    if (p.wpitch > 0) {
      const int nrows = H - slice * 128 < 128 ? H - slice * 128 : 128;
      tc::mbar_arrive_expect_tx(bar_w, (uint32_t)nrows * (uint32_t)K * 4u);
    }
  }
  if (warp == 0) tc::tmem_alloc<512>(tmem_slot);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  if (p.wpitch > 0 && g == 0 && jv) tc::bulk_g2s(wst + (size_t)(q * 32 + lane) * p.wpitch, wrow, (uint32_t)K * 4u, bar_w);
  const uint32_t tmem = *tmem_slot;
  const uint32_t lane_base = (uint32_t)(q * 32) << 16;
  const uint32_t tmem_d0 = tmem, tmem_d1 = tmem + NT;
  const uint32_t tmem_a = tmem + 2 * NT;
  const uint32_t plane_cols = Kmma / 2;
  {
    const bool staged = p.wpitch > 0;
    if (staged && !tc::mbar_wait_cta(bar_w, 0)) __trap();
    const float* srow = wst + (size_t)(q * 32 + lane) * p.wpitch;
    for (int c0 = 8 * g; c0 < (int)plane_cols; c0 += 32) {
      float wv[16];
      if (staged) {
#pragma unroll
        for (int v4 = 0; v4 < 4; ++v4) {
          const int k = 2 * c0 + 4 * v4;
          const float4 x = (jv && k < K) ? *reinterpret_cast<const float4*>(srow + k) : make_float4(0.f, 0.f, 0.f, 0.f);
          wv[4 * v4 + 0] = x.x; wv[4 * v4 + 1] = x.y; wv[4 * v4 + 2] = x.z; wv[4 * v4 + 3] = x.w;
        }
      } else {
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          const int k = 2 * c0 + e;
          wv[e] = (jv && k < K) ? __ldg(wrow + k) : 0.f;
        }
      }
      uint32_t vh[8], vm[8], vl[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        uint32_t h2[2], m2[2], l2[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) pre_split3(wv[2 * u + e], h2[e], m2[e], l2[e]);
        vh[u] = h2[0] | (h2[1] << 16);
        vm[u] = m2[0] | (m2[1] << 16);
        vl[u] = l2[0] | (l2[1] << 16);
      }
      tc::tmem_st8(tmem_a + lane_base + 0 * plane_cols + c0, vl);
      tc::tmem_st8(tmem_a + lane_base + 1 * plane_cols + c0, vm);
      tc::tmem_st8(tmem_a + lane_base + 2 * plane_cols + c0, vh);
    }
    tc::tmem_wait_st();
    __syncthreads();  // the staging bytes become the operand buffers
  }

  const int k_noisy = p.ctr + 2 * p.nbr;
  const int npl = (Kmma + 31) / 32;  // operand columns per lane (incl. zero padding up to Kmma)
  // ---- gather + LayerNorm + split of one row tile -> the three operand planes of buffer `dst` ----
  auto convert = [&](long long tile, uint8_t* dst) {
    const long long r0 = tile * NT;
    if (p.in_cnt) {  // the full-band output of every frame this tile touches must be complete
      if (tid == 0) {
        long long rl = r0 + NT - 1;
        if (rl >= M) rl = M - 1;
        const int t_lo = (int)(r0 / R), t_hi = (int)(rl / R);
        for (int t = t_lo; t <= t_hi; ++t) {
          unsigned int polls = 0;
          while (true) {
            unsigned int v;
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p.in_cnt + t) : "memory");
            if (v >= p.in_target) break;
            if (++polls > (1u << 23)) __trap();
            __nanosleep(40);
          }
        }
      }
      __syncthreads();
    }
    constexpr int RPW = NT / 16;  // rows per warp
#pragma unroll 1
    for (int rr = 0; rr < RPW; ++rr) {
      const int n = warp * RPW + rr;           // row of the tile
      const long long m = r0 + n;
      const bool rv = m < M;
      const int t = rv ? (int)(m / R) : 0;
      const int r = rv ? (int)(m - (long long)t * R) : 0;
      const int b = r / p.N, ns = r - b * p.N;
      const float* cm_row = p.cm + ((size_t)t * p.B + b) * p.f_cm;
      const float* fb_row = p.fb ? p.fb + ((size_t)t * p.B + b) * p.f_fb : nullptr;
      const int base = p.lo + ns * p.ctr;
      float v[kPreMaxPL];
      float sum = 0.f;
#pragma unroll
      for (int i = 0; i < kPreMaxPL; ++i) {
        const int jx = lane + 32 * i;
        v[i] = 0.f;
        if (rv && jx < K) {
          if (jx < k_noisy) {
            int qq = base - p.nbr + jx;
            qq = qq < 0 ? -qq : qq;
            qq = qq > p.f_cm - 1 ? 2 * (p.f_cm - 1) - qq : qq;
            v[i] = __ldg(cm_row + qq);
          } else {
            v[i] = __uint_as_float(pre_ld_cg(fb_row + (base + jx - k_noisy) % p.f_fb));
          }
          sum += v[i];
        }
        if (32 * (i + 1) >= K) break;
      }
      if (p.ln_w != nullptr) {  // same reduction order as k_subband_features: bit-identical x
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        const float mean = sum / (float)K;
        float sq = 0.f;
#pragma unroll
        for (int i = 0; i < kPreMaxPL; ++i) {
          const int jx = lane + 32 * i;
          if (jx < K) {
            const float d = v[i] - mean;
            sq += d * d;
          }
          if (32 * (i + 1) >= K) break;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
        const float rstd = 1.0f / sqrtf(sq / (float)K + p.eps);
#pragma unroll
        for (int i = 0; i < kPreMaxPL; ++i) {
          const int jx = lane + 32 * i;
          if (jx < K) v[i] = (v[i] - mean) * rstd * p.ln_w[jx] + p.ln_b[jx];
          if (32 * (i + 1) >= K) break;
        }
      }
      float* xo = (p.x_out && rv && slice == 0) ? p.x_out + (size_t)m * K : nullptr;
      const uint32_t rowoff = (uint32_t)((n >> 3) * SBO + (n & 7) * 16);
#pragma unroll
      for (int i = 0; i < kPreMaxPL; ++i) {
        if (i >= npl) break;
        const int jx = lane + 32 * i;
        if (jx < Kmma) {
          const float xv = (rv && jx < K) ? v[i] : 0.f;
          if (xo && jx < K) xo[jx] = xv;
          uint32_t hi, mid, lo;
          pre_split3(xv, hi, mid, lo);
          const uint32_t off = rowoff + (uint32_t)((jx >> 3) * 128 + (jx & 7) * 2);
          *reinterpret_cast<uint16_t*>(dst + off) = (uint16_t)lo;
          *reinterpret_cast<uint16_t*>(dst + plane_bytes + off) = (uint16_t)mid;
          *reinterpret_cast<uint16_t*>(dst + 2 * plane_bytes + off) = (uint16_t)hi;
        }
      }
    }
  };

  const uint32_t idesc = tc::make_idesc_f16(128, NT, true);
  const int ksteps = Kmma / 16;
  auto issue = [&](int buf) {  // warp 0 only
    tc::tc_fence_after();
    if (tc::elect_one()) {
      const uint64_t desc_b0 = tc::make_smem_desc(tc::smem_u32(smem + (size_t)buf * buf_bytes), 128, SBO);
      if (!mma_pairs<NT>(ksteps, buf ? tmem_d1 : tmem_d0, tmem_a, desc_b0, idesc)) __trap();
      tc::mma_commit(&bar[buf]);
    }
    __syncwarp();
  };

  const long long first = blockIdx.y;
  const long long my_tiles = first < ntiles_all ? (ntiles_all - first + P - 1) / P : 0;
  if (my_tiles > 0) {
    convert(first, smem);
    tc::fence_proxy_async_smem();
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) issue(0);
  }
  for (long long i = 0; i < my_tiles; ++i) {
    const int cur = (int)(i & 1), nxt = cur ^ 1;
    const long long tile = first + i * P;
    if (i + 1 < my_tiles) convert(tile + P, smem + (size_t)nxt * buf_bytes);  // overlaps MMA(i)
    tc::fence_proxy_async_smem();
    tc::tc_fence_before();
    __syncthreads();
    if (!tc::mbar_wait_cta(&bar[cur], (uint32_t)((i >> 1) & 1))) __trap();
    tc::tc_fence_after();
    // accumulators of tile i -> registers before the MMAs of tile i+1 are issued
    uint32_t zr[NT / 32][8];
#pragma unroll
    for (int c8 = 0; c8 < NT / 32; ++c8)
      tc::tmem_ld8((cur ? tmem_d1 : tmem_d0) + lane_base + g * (NT / 4) + c8 * 8, zr[c8]);
    tc::tmem_wait_ld();
    if (i + 1 < my_tiles) {
      tc::tc_fence_before();
      __syncthreads();
      if (warp == 0) issue(nxt);
    }
    // stores of tile i: thread = feature j, rows [g*NT/4, +NT/4) of the tile; coalesced over features
    const long long r0 = tile * NT + g * (NT / 4);
    const long long left = M - r0;
    if (jv && left > 0) {
      float* po = p.xproj + r0 * H + j;
#pragma unroll
      for (int u = 0; u < NT / 4; ++u)
        if (u < left) po[(size_t)u * H] = __uint_as_float(zr[u >> 3][u & 7]);
    }
    if (p.out_cnt) {
      __syncthreads();  // every thread's stores of this tile are issued
      if (tid == 0) {
        __threadfence();
        const long long ra = tile * NT;
        long long rb = ra + NT;
        if (rb > M) rb = M;
        for (long long t = ra / R; t * R < rb; ++t) {
          const long long a = t * R > ra ? t * R : ra, bnd = (t + 1) * R < rb ? (t + 1) * R : rb;
          asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p.out_cnt + t), "r"((unsigned int)(bnd - a))
                       : "memory");
        }
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc<512>(tmem);
  trace_end(p.trace, tslot);
}

template <int NT>
static int launch_pre(PreParams p, int ctas_per_slice, cudaStream_t st) {
  const size_t buf2 = 2 * 3 * (size_t)NT * p.Kmma * 2;
  p.wpitch = 0;
  size_t stage = 0;
  if (p.K % 4 == 0 && (reinterpret_cast<uintptr_t>(p.w_ih) & 15) == 0) {
    p.wpitch = pre_wpitch(p.K);
    stage = (size_t)128 * p.wpitch * sizeof(float);
    if (stage + 192 > tc::kMaxDynamicSmem) { p.wpitch = 0; stage = 0; }
  }
  size_t smem = ((buf2 > stage ? buf2 : stage) + 127) / 128 * 128 + 64;
  if (smem > tc::kMaxDynamicSmem) return fail(GSN_ENOSUP, "gsn_pre_stream: K=%d does not fit shared memory", p.K);
  if (smem < tc::kTmemExclusiveSmem) smem = tc::kTmemExclusiveSmem;
  GSN_CUDA(cudaFuncSetAttribute(k_pre_stream<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int slices = (p.H + 127) / 128;
  const long long ntiles = ((long long)p.T * p.B * p.N + NT - 1) / NT;
  long long P = ctas_per_slice < 1 ? 1 : ctas_per_slice;
  if (P > ntiles) P = ntiles;
  dim3 grid((unsigned)slices, (unsigned)P);
  k_pre_stream<NT><<<grid, kPreThreads, smem, st>>>(p);
  GSN_LAUNCH_CHECK("k_pre_stream");
  return GSN_OK;
}

}  // namespace gsn

extern "C" int gsn_pre_stream_supported(int K, int H) {
  const int Kmma = (K + 15) / 16 * 16;
  return (K >= 1 && K <= 32 * gsn::kPreMaxPL && H >= 1 && 3 * (Kmma / 2) + 2 * 32 <= 512) ? 1 : 0;
}

extern "C" int gsn_pre_stream(const float* cm, int f_cm, const float* fb, int f_fb, const float* ln_weight,
                              const float* ln_bias, float ln_eps, const float* w_ih, float* x_out, float* xproj,
                              const unsigned int* in_cnt, unsigned int in_target, unsigned int* out_cnt, int T, int B,
                              int N, int lo, int ctr, int nbr, int H, int ctas_per_slice, gsn_stream_t stream) {
  using namespace gsn;
  GSN_REQUIRE(cm && w_ih && xproj, "gsn_pre_stream: null pointer");
  GSN_REQUIRE(T > 0 && B > 0 && N > 0 && ctr > 0 && nbr >= 0 && lo >= 0 && H > 0, "gsn_pre_stream: bad shape");
  const int K = ctr + 2 * nbr + (fb ? ctr : 0);
  GSN_REQUIRE(gsn_pre_stream_supported(K, H), "gsn_pre_stream: K=%d H=%d not supported", K, H);
  GSN_REQUIRE(lo + N * ctr <= f_cm, "gsn_pre_stream: band leaves the spectrum");
  GSN_REQUIRE(lo == 0 || lo - nbr >= 0, "gsn_pre_stream: lower neighbourhood out of range");
  GSN_REQUIRE(lo + N * ctr == f_cm || lo + N * ctr + nbr <= f_cm, "gsn_pre_stream: upper neighbourhood out of range");
  GSN_REQUIRE(!fb || f_fb > 0, "gsn_pre_stream: f_fb");
  GSN_REQUIRE((ln_weight == nullptr) == (ln_bias == nullptr), "gsn_pre_stream: ln params");
  PreParams p{};
  p.cm = cm; p.fb = fb; p.ln_w = ln_weight; p.ln_b = ln_bias; p.w_ih = w_ih; p.x_out = x_out; p.xproj = xproj;
  p.in_cnt = in_cnt; p.in_target = in_target; p.out_cnt = out_cnt;
  p.T = T; p.B = B; p.N = N; p.lo = lo; p.ctr = ctr; p.nbr = nbr; p.f_cm = f_cm; p.f_fb = f_fb; p.K = K;
  p.Kmma = (K + 15) / 16 * 16; p.H = H; p.eps = ln_eps; p.trace = trace_buffer();
  if (3 * (p.Kmma / 2) + 2 * 64 <= 512) return launch_pre<64>(p, ctas_per_slice, as_stream(stream));
  return launch_pre<32>(p, ctas_per_slice, as_stream(stream));
}
