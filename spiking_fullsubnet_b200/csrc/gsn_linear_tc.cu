// tcgen05 linear for SPIKE inputs:  out[M,N] = a[M,K] @ w[N,K]^T + bias,  a in {0,1} (any bf16-exact values).
// Used for the input-to-hidden product of layers >= 1 (ESN:141 with the previous layer's spikes as input) and
// for proj (MSF:118): both have an exactly-bf16 left operand, so with the fp32 weights split into three exact
// bf16 planes (w = hi + mid + lo) every product is exact and the result is an fp32 sum of exact terms -- the
// same argument as for the recurrence (gsn_recurrence_tc.cu).
//
// Same "weights stationary in TMEM, swap-AB" layout as the recurrence: CTA (slice, p) owns the 128 output
// features [128 slice, +128) -- their weight rows live in tensor memory as the A operand for the whole kernel --
// and walks over row tiles p, p+P, p+2P, ... (persistent).  Per tile: NT rows of `a` are converted to a bf16
// K-major B operand in shared memory (double buffered), D[128 features x NT rows] accumulates 3*K/16 MMAs in
// one of two TMEM accumulators, and the epilogue (bias, optional activation, coalesced stores over features)
// of tile i runs while the tensor pipe works on tile i+1 and the loads of tile i+2 are in flight.
#include "gsn_common.cuh"
#include "gsn_tc.cuh"

namespace gsn {

constexpr uint32_t kLinTmemCols = 512;

__device__ __forceinline__ float lin_act(float v, int act) {
  switch (act) {
    case 1: return tanhf(v);
    case 2: return 1.0f / (1.0f + expf(-v));
    case 3: return fmaxf(v, 0.f);
    default: return v;
  }
}

__device__ __forceinline__ void lin_split3(float w, uint32_t& hi, uint32_t& mid, uint32_t& lo) {
  const uint32_t wb = __float_as_uint(w);
  hi = wb >> 16;
  const float r1 = w - __uint_as_float(wb & 0xFFFF0000u);
  const uint32_t r1b = __float_as_uint(r1);
  mid = r1b >> 16;
  const float r2 = r1 - __uint_as_float(r1b & 0xFFFF0000u);
  lo = __float_as_uint(r2) >> 16;
}

// Row pitch (floats) of the weight staging area: 16-byte aligned rows, pitch/4 odd (conflict-free 16-byte reads
// by 8 consecutive lanes = 8 consecutive rows).
static inline __host__ __device__ int lin_wpitch(int K) { return ((K / 4) & 1) ? K : K + 4; }

// BITS: `a_bits` holds the spike trace bit-packed ([M, W] uint32, W = ceil(K/32), neuron k = bit k%32 of word
// k/32, as the recurrence kernels emit it): 1 bit instead of 4 bytes read per spike.
template <int NT, bool BITS>
__global__ void __launch_bounds__(256, 1)
    k_linear_tc(const float* __restrict__ a, const uint32_t* __restrict__ a_bits, const float* __restrict__ w,
                const float* __restrict__ bias, float* __restrict__ out, float* __restrict__ out_act, int act,
                long long M, int K, int N, int Kmma, int wpitch, TraceBuf* tb, const unsigned int* in_cnt,
                unsigned int in_target, unsigned int* out_cnt, int Rf) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int tslot = trace_begin(tb, 4, (int)M, K, N);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q = warp & 3, g = warp >> 2;  // lane quarter / row half
  const int slice = blockIdx.x;
  const int P = gridDim.y;
  const long long ntiles_all = (M + NT - 1) / NT;
  const int j = slice * 128 + q * 32 + lane;  // output feature of this thread (TMEM lane)
  const bool jv = j < N;

  const uint32_t SBO = 16u * Kmma;
  const size_t b_bytes = ((size_t)NT * Kmma * 2 + 127) / 128 * 128;
  uint8_t* const sB0 = smem;
  uint8_t* const sB1 = smem + b_bytes;
  // the weight staging area of the prologue shares the bytes of the two B buffers (and may extend past them)
  const size_t stage_bytes = wpitch > 0 ? (size_t)128 * wpitch * 4 : 0;
  const size_t bar_off = ((2 * b_bytes > stage_bytes ? 2 * b_bytes : stage_bytes) + 127) / 128 * 128;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + bar_off);  // [2] MMA completion, [1] weight staging
  uint64_t* bar_w = bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 3);
  float* wst = reinterpret_cast<float*>(smem);
  const float* wrow = w + (size_t)(jv ? j : 0) * K;

  if (tid == 0) {
    tc::mbar_init(&bar[0], 1);
    tc::mbar_init(&bar[1], 1);
    tc::mbar_init(bar_w, 1);
    tc::fence_mbar_init();
    if (wpitch > 0) {
      const int nrows = N - slice * 128 < 128 ? N - slice * 128 : 128;
      tc::mbar_arrive_expect_tx(bar_w, (uint32_t)nrows * (uint32_t)K * 4u);
    }
  }
  if (warp == 0) tc::tmem_alloc<kLinTmemCols>(tmem_slot);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  // one 1-D bulk copy (TMA unit) per weight row of this CTA's 128 output features: coalesced, asynchronous
  if (wpitch > 0 && g == 0 && jv)
    tc::bulk_g2s(wst + (size_t)(q * 32 + lane) * wpitch, wrow, (uint32_t)K * 4u, bar_w);
  const uint32_t tmem = *tmem_slot;
  const uint32_t lane_base = (uint32_t)(q * 32) << 16;
  const uint32_t tmem_d0 = tmem, tmem_d1 = tmem + NT;
  const uint32_t tmem_a = tmem + 2 * NT;
  const uint32_t plane_cols = Kmma / 2;

  // weights of my feature -> three exact bf16 planes in TMEM (the two row-half warps split the K range)
  {
    const bool staged = wpitch > 0;
    if (staged && !tc::mbar_wait_cta(bar_w, 0)) __trap();
    const float* srow = wst + (size_t)(q * 32 + lane) * wpitch;
    for (int c0 = 8 * g; c0 < (int)plane_cols; c0 += 16) {
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
        for (int e = 0; e < 2; ++e) lin_split3(wv[2 * u + e], h2[e], m2[e], l2[e]);
        vh[u] = h2[0] | (h2[1] << 16);
        vm[u] = m2[0] | (m2[1] << 16);
        vl[u] = l2[0] | (l2[1] << 16);
      }
      tc::tmem_st8(tmem_a + lane_base + 0 * plane_cols + c0, vl);
      tc::tmem_st8(tmem_a + lane_base + 1 * plane_cols + c0, vm);
      tc::tmem_st8(tmem_a + lane_base + 2 * plane_cols + c0, vh);
    }
    tc::tmem_wait_st();
    if (staged) __syncthreads();  // the staging bytes become the B buffers
  }
  const float bj = (bias && jv) ? bias[j] : 0.f;

  // B operand of one row tile: (row n, 8 consecutive k) -> 16 bytes; 8 consecutive threads fill one core matrix.
  // UNR tasks per thread are loaded before any is converted, so 2*UNR 16-byte loads are in flight per thread.
  const int k8n = Kmma / 8;
  const int ntask = NT * k8n;
  constexpr int UNR = 4;
  const int Wb = (K + 31) / 32;  // words per row of the bit-packed trace
  // streaming use (in_cnt / out_cnt, Rf = rows per frame): `a_bits` is produced frame by frame by a concurrently
  // running recurrence kernel -- wait for the frames a tile touches, read them past the L1, publish what was written
  auto convert = [&](long long tile, uint8_t* dst) {
    const long long r0 = tile * NT;
    if (in_cnt) {
      if (tid == 0) {
        long long rl = r0 + NT - 1;
        if (rl >= M) rl = M - 1;
        for (long long t = r0 / Rf; t <= rl / Rf; ++t) {
          unsigned int polls = 0;
          while (true) {
            unsigned int v;
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(in_cnt + t) : "memory");
            if (v >= in_target) break;
            if (++polls > (1u << 23)) __trap();
            __nanosleep(40);
          }
        }
      }
      __syncthreads();
    }
    if (BITS) {
      // task = (row n, one 32-bit word of its packed trace) -> up to four 16-byte chunks of the operand; all the
      // words of a thread are loaded before any is expanded (independent loads in flight, not a dependent chain)
      const int nw = (k8n + 3) / 4;
      constexpr int MAXW = (NT * 10 + 255) / 256;  // Kmma <= 320: at most 10 words per row
      uint32_t wd[MAXW];
#pragma unroll
      for (int it = 0; it < MAXW; ++it) {
        const int i = tid + 256 * it;
        const int nlo = i & 7, wi = (i >> 3) % nw, nhi = (i >> 3) / nw;
        const long long row = r0 + nhi * 8 + nlo;
        uint32_t wv_ = 0u;
        if (i < NT * nw && row < M) {
          if (in_cnt) asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(wv_) : "l"(a_bits + row * Wb + wi) : "memory");
          else wv_ = __ldg(a_bits + row * Wb + wi);
        }
        wd[it] = wv_;
      }
#pragma unroll
      for (int it = 0; it < MAXW; ++it) {
        const int i = tid + 256 * it;
        if (i >= NT * nw) break;
        const int nlo = i & 7, wi = (i >> 3) % nw, nhi = (i >> 3) / nw;
#pragma unroll
        for (int e4 = 0; e4 < 4; ++e4) {
          const int k8 = 4 * wi + e4;
          if (k8 >= k8n) break;
          const uint32_t b8 = (wd[it] >> (8 * e4)) & 0xFFu;
          uint32_t v[4];
#pragma unroll
          for (int e = 0; e < 4; ++e)
            v[e] = ((b8 >> (2 * e)) & 1u ? 0x3F80u : 0u) | ((b8 >> (2 * e + 1)) & 1u ? 0x3F800000u : 0u);
          *reinterpret_cast<uint4*>(dst + (uint32_t)(nhi * SBO + k8 * 128 + nlo * 16)) = make_uint4(v[0], v[1], v[2], v[3]);
        }
      }
      return;
    }
    for (int base = tid; base < ntask; base += 256 * UNR) {
      float4 x0[UNR], x1[UNR];
      uint32_t doff[UNR];
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        const int i = base + 256 * u;
        const int nlo = i & 7, k8 = (i >> 3) % k8n, nhi = (i >> 3) / k8n;
        const long long row = r0 + nhi * 8 + nlo;
        doff[u] = i < ntask ? (uint32_t)(nhi * SBO + k8 * 128 + nlo * 16) : 0xFFFFFFFFu;
        x0[u] = x1[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (i < ntask && row < M) {
          const float* src = a + row * K + k8 * 8;
          if (k8 * 8 + 8 <= K) {
            x0[u] = __ldg(reinterpret_cast<const float4*>(src));
            x1[u] = __ldg(reinterpret_cast<const float4*>(src) + 1);
          } else {
            float t8[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) t8[e] = (k8 * 8 + e < K) ? __ldg(src + e) : 0.f;
            x0[u] = make_float4(t8[0], t8[1], t8[2], t8[3]);
            x1[u] = make_float4(t8[4], t8[5], t8[6], t8[7]);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        if (doff[u] == 0xFFFFFFFFu) continue;
        const float xs[8] = {x0[u].x, x0[u].y, x0[u].z, x0[u].w, x1[u].x, x1[u].y, x1[u].z, x1[u].w};
        uint32_t v[4] = {0, 0, 0, 0};
#pragma unroll
        for (int e = 0; e < 8; ++e)
          v[e >> 1] |= (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(xs[e])) << (16 * (e & 1));
        *reinterpret_cast<uint4*>(dst + doff[u]) = make_uint4(v[0], v[1], v[2], v[3]);
      }
    }
  };

  const uint32_t idesc = tc::make_idesc_f16(128, NT, true);
  const int ksteps = Kmma / 16;
  auto issue = [&](int buf) {  // warp 0 only
    tc::tc_fence_after();
    if (tc::elect_one()) {
      const uint64_t desc_b0 = tc::make_smem_desc(tc::smem_u32(buf ? sB1 : sB0), 128, SBO);
      const uint32_t td = buf ? tmem_d1 : tmem_d0;
      tc::mma_planes<3>(ksteps, td, tmem_a, desc_b0, idesc);  // straight-line issue, lo plane first
      tc::mma_commit(buf ? &bar[1] : &bar[0]);
    }
    __syncwarp();
  };

  // tiles of this CTA: p, p+P, ...
  const long long first = blockIdx.y;
  const long long my_tiles = first < ntiles_all ? (ntiles_all - first + P - 1) / P : 0;
  bool alive = true;
  if (my_tiles > 0) {
    convert(first, sB0);
    tc::fence_proxy_async_smem();
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) issue(0);
  }
  // development aid: with a trace buffer set, thread 0 of CTA (0,0) accumulates per-phase cycles into its header
  const bool prof = tb != nullptr && tslot >= 0;
  unsigned int pc[5] = {0, 0, 0, 0, 0};
  for (long long i = 0; i < my_tiles; ++i) {
    const int cur = (int)(i & 1), nxt = cur ^ 1;
    const long long tile = first + i * P;
    const long long q0 = prof ? clock64() : 0;
    if (i + 1 < my_tiles) convert(tile + P, nxt ? sB1 : sB0);  // overlaps MMA(i)
    const long long q1 = prof ? clock64() : 0;
    tc::fence_proxy_async_smem();
    tc::tc_fence_before();
    __syncthreads();
    const long long q2 = prof ? clock64() : 0;
    if (!tc::mbar_wait_cta(cur ? &bar[1] : &bar[0], (uint32_t)((i >> 1) & 1))) { alive = false; break; }
    tc::tc_fence_after();
    const long long q3 = prof ? clock64() : 0;
    // accumulators of tile i -> registers BEFORE the next MMAs are issued: tcgen05.ld issued behind in-flight MMAs
    // waits for them (measured: 3650 cycles per tile for this epilogue when it ran under the MMAs of tile i+1)
    uint32_t zr[NT / 16][8];  // the destination registers are only read after tcgen05.wait::ld
#pragma unroll
    for (int c8 = 0; c8 < NT / 16; ++c8)
      tc::tmem_ld8((cur ? tmem_d1 : tmem_d0) + lane_base + g * (NT / 2) + c8 * 8, zr[c8]);
    tc::tmem_wait_ld();
    if (i + 1 < my_tiles) {
      tc::tc_fence_before();
      __syncthreads();
      if (warp == 0) issue(nxt);  // the tensor pipe works on tile i+1 while tile i is stored
    }
    const long long q4 = prof ? clock64() : 0;
    // stores of tile i: thread = feature j, rows [g*NT/2, +NT/2) of the tile; coalesced over features.
    // Everything row-invariant is hoisted: a full half tile is a straight run of stores (no per-row predicate).
    const long long r0 = tile * NT + g * (NT / 2);
    const long long left = M - r0;
    if (jv && left > 0) {
      float* po = out + r0 * N + j;
      if (left >= NT / 2) {
#pragma unroll
        for (int u = 0; u < NT / 2; ++u) po[(size_t)u * N] = __uint_as_float(zr[u >> 3][u & 7]) + bj;
      } else {
#pragma unroll
        for (int u = 0; u < NT / 2; ++u)
          if (u < (int)left) po[(size_t)u * N] = __uint_as_float(zr[u >> 3][u & 7]) + bj;
      }
      if (out_act) {
        float* pa = out_act + r0 * N + j;
#pragma unroll
        for (int u = 0; u < NT / 2; ++u)
          if (u < left) pa[(size_t)u * N] = lin_act(__uint_as_float(zr[u >> 3][u & 7]) + bj, act);
      }
    }
    if (out_cnt) {
      __syncthreads();  // every thread's stores of this tile are issued
      if (tid == 0) {
        __threadfence();
        const long long ra = tile * NT;
        long long rb = ra + NT;
        if (rb > M) rb = M;
        for (long long t = ra / Rf; t * Rf < rb; ++t) {
          const long long lo_ = t * Rf > ra ? t * Rf : ra, hi_ = (t + 1) * Rf < rb ? (t + 1) * Rf : rb;
          asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(out_cnt + t), "r"((unsigned int)(hi_ - lo_))
                       : "memory");
        }
      }
    }
    if (prof) {
      const long long q5 = clock64();
      pc[0] += (unsigned)(q1 - q0); pc[1] += (unsigned)(q2 - q1); pc[2] += (unsigned)(q3 - q2);
      pc[3] += (unsigned)(q4 - q3); pc[4] += (unsigned)(q5 - q4);
    }
  }
  if (prof) {  // convert, sync, MMA wait, MMA issue, epilogue, tiles (last traced linear wins)
    for (int u = 0; u < 5; ++u) tb->pad[u] = pc[u];
    tb->pad[5] = (unsigned)my_tiles;
  }
  if (!alive) __trap();
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc<kLinTmemCols>(tmem);
  trace_end(tb, tslot);
}

template <int NT, bool BITS>
static int launch_linear_tc(const float* a, const uint32_t* a_bits, const float* w, const float* bias, float* out,
                            float* out_act, int act, long long M, int K, int N, int sm_budget, cudaStream_t st,
                            const unsigned int* in_cnt = nullptr, unsigned int in_target = 0,
                            unsigned int* out_cnt = nullptr, int Rf = 1) {
  const int Kmma = (K + 15) / 16 * 16;
  const size_t b2 = 2 * (((size_t)NT * Kmma * 2 + 127) / 128 * 128);
  // weight rows staged through shared memory with bulk copies when they are 16-byte aligned
  int wpitch = 0;
  size_t stage = 0;
  if (K % 4 == 0 && (reinterpret_cast<uintptr_t>(w) & 15) == 0) {
    wpitch = lin_wpitch(K);
    stage = (size_t)128 * wpitch * sizeof(float);
    if (stage + 192 > tc::kMaxDynamicSmem) { wpitch = 0; stage = 0; }
  }
  size_t smem = ((b2 > stage ? b2 : stage) + 127) / 128 * 128 + 64;
  if (smem < tc::kTmemExclusiveSmem) smem = tc::kTmemExclusiveSmem;
  GSN_CUDA(cudaFuncSetAttribute(k_linear_tc<NT, BITS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int dev = 0, sms = 148;
  GSN_CUDA(cudaGetDevice(&dev));
  GSN_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  if (sm_budget > 0 && sm_budget < sms) sms = sm_budget;
  const int slices = (N + 127) / 128;
  const long long ntiles = (M + NT - 1) / NT;
  long long P = sms / slices;
  if (P < 1) P = 1;
  if (P > ntiles) P = ntiles;
  dim3 grid((unsigned)slices, (unsigned)P);
  k_linear_tc<NT, BITS><<<grid, 256, smem, st>>>(a, a_bits, w, bias, out, out_act, act, M, K, N, Kmma, wpitch,
                                                 trace_buffer(), in_cnt, in_target, out_cnt, Rf);
  GSN_LAUNCH_CHECK("k_linear_tc");
  return GSN_OK;
}

template <bool BITS>
static int dispatch_linear_tc(const float* a, const uint32_t* a_bits, const float* w, const float* bias, float* out,
                              float* out_act, int act, long long M, int K, int N, int sm_budget, cudaStream_t st,
                              const unsigned int* in_cnt = nullptr, unsigned int in_target = 0,
                              unsigned int* out_cnt = nullptr, int Rf = 1) {
  const int Kmma = (K + 15) / 16 * 16;
  if (3 * Kmma / 2 + 2 * 64 <= 512)
    return launch_linear_tc<64, BITS>(a, a_bits, w, bias, out, out_act, act, M, K, N, sm_budget, st, in_cnt,
                                      in_target, out_cnt, Rf);
  if (3 * Kmma / 2 + 2 * 16 <= 512)
    return launch_linear_tc<16, BITS>(a, a_bits, w, bias, out, out_act, act, M, K, N, sm_budget, st, in_cnt,
                                      in_target, out_cnt, Rf);
  return fail(GSN_ENOSUP, "gsn_linear_spikes: K=%d does not fit tensor memory (K <= 320)", K);
}

// fp32 {0,1} trace -> bit-packed trace (one warp per 32 neurons of a row)
__global__ void __launch_bounds__(256) k_pack_spikes(const float* __restrict__ h, uint32_t* __restrict__ bits,
                                                     long long rows, int H, int W) {
  const long long gw = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (gw >= rows * W) return;
  const long long r = gw / W;
  const int wd = (int)(gw % W), k = wd * 32 + lane;
  const uint32_t word = __ballot_sync(0xffffffffu, k < H && h[r * H + k] != 0.f);
  if (lane == 0) bits[gw] = word;
}

}  // namespace gsn

extern "C" int gsn_linear_spikes(const float* a, const float* w, const float* bias, float* out, float* out_act,
                                 int act, int64_t M, int K, int N, int sm_budget, gsn_stream_t stream) {
  GSN_REQUIRE(a && w && out, "gsn_linear_spikes: null pointer");
  GSN_REQUIRE(M > 0 && K > 0 && N > 0, "gsn_linear_spikes: bad shape M=%lld K=%d N=%d", (long long)M, K, N);
  GSN_REQUIRE(act >= 0 && act <= 3, "gsn_linear_spikes: unknown activation %d", act);
  GSN_REQUIRE(K % 4 == 0, "gsn_linear_spikes: K=%d must be a multiple of 4 (16-byte row alignment)", K);
  GSN_REQUIRE((reinterpret_cast<uintptr_t>(a) & 15) == 0, "gsn_linear_spikes: a must be 16-byte aligned");
  return gsn::dispatch_linear_tc<false>(a, nullptr, w, bias, out, out_act, act, M, K, N, sm_budget,
                                        gsn::as_stream(stream));
}

extern "C" int gsn_linear_spike_bits(const uint32_t* a_bits, const float* w, const float* bias, float* out,
                                     float* out_act, int act, int64_t M, int K, int N, int sm_budget,
                                     gsn_stream_t stream) {
  GSN_REQUIRE(a_bits && w && out, "gsn_linear_spike_bits: null pointer");
  GSN_REQUIRE(M > 0 && K > 0 && N > 0, "gsn_linear_spike_bits: bad shape M=%lld K=%d N=%d", (long long)M, K, N);
  GSN_REQUIRE(act >= 0 && act <= 3, "gsn_linear_spike_bits: unknown activation %d", act);
  return gsn::dispatch_linear_tc<true>(nullptr, a_bits, w, bias, out, out_act, act, M, K, N, sm_budget,
                                       gsn::as_stream(stream));
}

extern "C" int gsn_pack_spikes(const float* h, uint32_t* bits, int64_t rows, int H, gsn_stream_t stream) {
  GSN_REQUIRE(h && bits, "gsn_pack_spikes: null pointer");
  GSN_REQUIRE(rows > 0 && H > 0, "gsn_pack_spikes: bad shape rows=%lld H=%d", (long long)rows, H);
  const int W = (H + 31) / 32;
  const long long warps = rows * W;
  const unsigned blocks = (unsigned)((warps * 32 + 255) / 256);
  gsn::k_pack_spikes<<<blocks, 256, 0, gsn::as_stream(stream)>>>(h, bits, rows, H, W);
  GSN_LAUNCH_CHECK("k_pack_spikes");
  return GSN_OK;
}
