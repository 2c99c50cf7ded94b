// gsn_tc_selftest: one [128 x N] = A[128 x K] * B[N x K]^T product on tcgen05 with the exact operand
// layouts the recurrence kernel uses (K-major no-swizzle shared-memory operands; A optionally resident
// in TMEM).  tests/test_gpu_tc_selftest.py checks it against a CPU product; it pins the descriptor
// encodings independently of the recurrence logic.
#include "gsn_common.cuh"
#include "gsn_tc.cuh"

namespace gsn {

__global__ void __launch_bounds__(128)
    k_tc_probe(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ d,
               int* __restrict__ status, int N, int K, int a_in_tmem, int swap_lbo_sbo, int use_fp16) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t lbo = 128, sbo = 16u * K;
  uint8_t* sA = smem;
  uint8_t* sB = smem + 128 * K * 2;

  auto cvt = [&](float v) -> uint16_t {
    if (use_fp16) return __half_as_ushort(__float2half_rn(v));
    return __bfloat16_as_ushort(__float2bfloat16_rn(v));
  };
  for (int i = tid; i < 128 * K; i += 128) {
    const int r = i / K, k = i % K;
    *reinterpret_cast<uint16_t*>(sA + (r / 8) * sbo + (k / 8) * lbo + (r % 8) * 16 + (k % 8) * 2) = cvt(a[i]);
  }
  for (int i = tid; i < N * K; i += 128) {
    const int r = i / K, k = i % K;
    *reinterpret_cast<uint16_t*>(sB + (r / 8) * sbo + (k / 8) * lbo + (r % 8) * 16 + (k % 8) * 2) = cvt(b[i]);
  }
  if (tid == 0) {
    tc::mbar_init(&bar, 1);
    tc::fence_mbar_init();
  }
  if (warp == 0) tc::tmem_alloc<512>(&tmem_base_slot);
  tc::tc_fence_before();
  tc::fence_proxy_async_smem();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = tmem_base_slot;
  const uint32_t tmem_d = tmem;            // columns [0, N)
  const uint32_t tmem_a = tmem + 256;      // columns [256, 256 + K/2)
  const uint32_t lane_base = (uint32_t)(warp * 32) << 16;

  if (a_in_tmem) {
    // thread m owns TMEM lane m: pack two consecutive k per 32-bit column
    for (int c0 = 0; c0 < K / 2; c0 += 8) {
      uint32_t v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int k = 2 * (c0 + j);
        v[j] = (uint32_t)cvt(a[tid * K + k]) | ((uint32_t)cvt(a[tid * K + k + 1]) << 16);
      }
      tc::tmem_st8(tmem_a + lane_base + c0, v);
    }
    tc::tmem_wait_st();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
  }

  if (tid == 0) {
    const uint32_t idesc = tc::make_idesc_f16(128, N, !use_fp16);
    const uint32_t l = swap_lbo_sbo ? sbo : lbo, s = swap_lbo_sbo ? lbo : sbo;
    for (int ks = 0; ks < K / 16; ++ks) {
      const uint64_t db = tc::make_smem_desc(tc::smem_u32(sB) + ks * 2 * lbo, l, s);
      if (a_in_tmem) {
        tc::mma_ts(tmem_d, tmem_a + ks * 8, db, idesc, ks > 0);
      } else {
        const uint64_t da = tc::make_smem_desc(tc::smem_u32(sA) + ks * 2 * lbo, l, s);
        tc::mma_ss(tmem_d, da, db, idesc, ks > 0);
      }
    }
    tc::mma_commit(&bar);
  }
  const bool ok = tc::mbar_wait(&bar, 0);
  tc::tc_fence_after();
  if (!ok) {
    if (tid == 0) *status = 1;
  } else {
    for (int c0 = 0; c0 < N; c0 += 16) {
      uint32_t v[16];
      tc::tmem_ld16(tmem_d + lane_base + c0, v);
      tc::tmem_wait_ld();
#pragma unroll
      for (int j = 0; j < 16; ++j) d[tid * N + c0 + j] = __uint_as_float(v[j]);
    }
    if (tid == 0) *status = 0;
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc<512>(tmem);
}

// timing probe: `reps` x (K/16) back-to-back MMAs of shape 128 x N x 16, operands uninitialised.
__global__ void __launch_bounds__(128) k_tc_mma_timing(long long* out, int N, int K, int reps, int a_in_tmem) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (128 + N) * K / 2; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (tid == 0) {
    tc::mbar_init(&bar, 1);
    tc::fence_mbar_init();
  }
  if (warp == 0) tc::tmem_alloc<512>(&tmem_base_slot);
  tc::tc_fence_before();
  tc::fence_proxy_async_smem();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = tmem_base_slot;
  const uint32_t lbo = 128, sbo = 16u * K;
  long long t0 = 0, t1 = 0, t2 = 0;
  if (warp == 0) {
    uint32_t leader;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(leader));
    if (leader) {
      const uint32_t idesc = tc::make_idesc_f16(128, N, true);
      const uint64_t da0 = tc::make_smem_desc(tc::smem_u32(smem), lbo, sbo);
      const uint64_t db0 = tc::make_smem_desc(tc::smem_u32(smem + 128 * K * 2), lbo, sbo);
      t0 = clock64();
      for (int r = 0; r < reps; ++r)
        for (int ks = 0; ks < K / 16; ++ks) {
          if (a_in_tmem) tc::mma_ts(tmem, tmem + 256 + ks * 8, db0 + (uint64_t)(ks * 16), idesc, 1);
          else tc::mma_ss(tmem, da0 + (uint64_t)(ks * 16), db0 + (uint64_t)(ks * 16), idesc, 1);
        }
      tc::mma_commit(&bar);
      t1 = clock64();
    }
    __syncwarp();
  }
  tc::mbar_wait(&bar, 0);
  t2 = clock64();
  if (tid == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc<512>(tmem);
}

// timing probe 2: fully unrolled issue (operands = base + immediates), M = 128 or 64, A in tensor memory
template <int KS, int MM>
__global__ void __launch_bounds__(128) k_tc_mma_timing2(long long* out, int N) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  constexpr int K = KS * 16;
  for (int i = tid; i < N * K / 2; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (tid == 0) {
    tc::mbar_init(&bar, 1);
    tc::fence_mbar_init();
  }
  if (warp == 0) tc::tmem_alloc<512>(&tmem_base_slot);
  tc::tc_fence_before();
  tc::fence_proxy_async_smem();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = tmem_base_slot;
  long long t0 = 0, t1 = 0, t2 = 0;
  if (warp == 0) {
    if (tc::elect_one()) {
      const uint32_t idesc = tc::make_idesc_f16(MM, N, true);
      const uint64_t db0 = tc::make_smem_desc(tc::smem_u32(smem), 128, 16u * K);
      const uint32_t ta = tmem + 64;
      t0 = clock64();
#pragma unroll
      for (int pl = 0; pl < 3; ++pl)
#pragma unroll
        for (int ks = 0; ks < KS; ++ks)
          asm volatile("tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, 1;" ::"r"(tmem),
                       "r"(ta + (uint32_t)((pl * KS + ks) * 8)), "l"(db0 + (uint64_t)(ks * 16)), "r"(idesc)
                       : "memory");
      tc::mma_commit(&bar);
      t1 = clock64();
    }
    __syncwarp();
  }
  tc::mbar_wait(&bar, 0);
  t2 = clock64();
  if (tid == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc<512>(tmem);
}

// int8 probe: d[128,N] (int32) = a[128,K] (int8, resident in TMEM: 4 consecutive k per 32-bit column) x
// b[N,K] (uint8, K-major no-swizzle shared memory: 8 rows x 16 bytes core matrices), K = 32 per instruction.
__global__ void __launch_bounds__(128)
    k_tc_probe_i8(const int* __restrict__ a, const int* __restrict__ b, int* __restrict__ d, long long* timing,
                  int N, int K, int a_signed, int reps) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t lbo = 128, sbo = 8u * K;
  uint8_t* sB = smem;
  for (int i = tid; i < N * K; i += 128) {
    const int r = i / K, k = i % K;
    sB[(r / 8) * sbo + (k / 16) * lbo + (r % 8) * 16 + (k % 16)] = (uint8_t)b[i];
  }
  if (tid == 0) {
    tc::mbar_init(&bar, 1);
    tc::fence_mbar_init();
  }
  if (warp == 0) tc::tmem_alloc<512>(&tmem_base_slot);
  tc::tc_fence_before();
  tc::fence_proxy_async_smem();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = tmem_base_slot;
  const uint32_t tmem_d = tmem, tmem_a = tmem + 256;
  const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
  for (int c0 = 0; c0 < K / 4; c0 += 8) {
    uint32_t v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      v[j] = 0;
#pragma unroll
      for (int e = 0; e < 4; ++e) v[j] |= ((uint32_t)(a[tid * K + 4 * (c0 + j) + e] & 0xFF)) << (8 * e);
    }
    tc::tmem_st8(tmem_a + lane_base + c0, v);
  }
  tc::tmem_wait_st();
  tc::tc_fence_before();
  __syncthreads();
  long long t0 = 0, t1 = 0;
  if (warp == 0) {
    tc::tc_fence_after();
    if (tc::elect_one()) {
      const uint32_t idesc = tc::make_idesc_i8(128, N, a_signed != 0, false);
      const uint64_t db0 = tc::make_smem_desc(tc::smem_u32(sB), lbo, sbo);
      t0 = clock64();
      for (int r = 0; r < reps; ++r)
        for (int ks = 0; ks < K / 32; ++ks)
          tc::mma_i8_ts(tmem_d, tmem_a + ks * 8, db0 + (uint64_t)(ks * 16), idesc, (r | ks) > 0);
      tc::mma_commit(&bar);
      t1 = clock64();
    }
    __syncwarp();
  }
  const bool ok = tc::mbar_wait(&bar, 0);
  const long long t2 = clock64();
  tc::tc_fence_after();
  if (ok) {
    for (int c0 = 0; c0 < N; c0 += 16) {
      uint32_t v[16];
      tc::tmem_ld16(tmem_d + lane_base + c0, v);
      tc::tmem_wait_ld();
#pragma unroll
      for (int j = 0; j < 16; ++j) d[tid * N + c0 + j] = (int)v[j];
    }
  }
  if (tid == 0) { timing[0] = ok ? t1 - t0 : -1; timing[1] = t2 - t0; }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc<512>(tmem);
}

}  // namespace gsn

extern "C" GSN_API int gsn_tc_probe_i8(const int* a, const int* b, int* d, long long* timing, int N, int K,
                                       int a_signed, int reps, gsn_stream_t stream) {
  const size_t smem = (size_t)N * K;
  GSN_CUDA(cudaFuncSetAttribute(gsn::k_tc_probe_i8, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  gsn::k_tc_probe_i8<<<1, 128, smem, gsn::as_stream(stream)>>>(a, b, d, timing, N, K, a_signed, reps);
  GSN_LAUNCH_CHECK("k_tc_probe_i8");
  return GSN_OK;
}

extern "C" GSN_API int gsn_tc_mma_timing(long long* out, int N, int K, int reps, int a_in_tmem, gsn_stream_t stream) {
  const size_t smem = (size_t)(128 + N) * K * 2;
  GSN_REQUIRE(smem <= 200 * 1024, "gsn_tc_mma_timing: operands do not fit shared memory");
  GSN_CUDA(cudaFuncSetAttribute(gsn::k_tc_mma_timing, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  gsn::k_tc_mma_timing<<<1, 128, smem, gsn::as_stream(stream)>>>(out, N, K, reps, a_in_tmem);
  GSN_LAUNCH_CHECK("k_tc_mma_timing");
  return GSN_OK;
}

// undeclared dev symbol (tools/tc_mma_timing.py): variant 0 = K 160 M 128, 1 = K 160 M 64, 2 = K 240 M 128
extern "C" GSN_API int gsn_tc_mma_timing2(long long* out, int N, int variant, gsn_stream_t stream) {
  const size_t smem = (size_t)N * 240 * 2 + 1024;
  cudaStream_t st = gsn::as_stream(stream);
  if (variant == 0) gsn::k_tc_mma_timing2<10, 128><<<1, 128, smem, st>>>(out, N);
  else if (variant == 1) gsn::k_tc_mma_timing2<10, 64><<<1, 128, smem, st>>>(out, N);
  else gsn::k_tc_mma_timing2<15, 128><<<1, 128, smem, st>>>(out, N);
  GSN_LAUNCH_CHECK("k_tc_mma_timing2");
  return GSN_OK;
}

// d[128, N] = a[128, K] @ b[N, K]^T on one CTA with the operand layouts of the recurrence kernels (K-major no-swizzle
// shared memory; a_in_tmem != 0 keeps A resident in tensor memory); a and b must hold values exactly representable in
// bf16 (fp16 if use_fp16).  status[0] = 0 ok, 1 = timeout.  swap_lbo_sbo is a diagnostic knob (0 for a correct result).
extern "C" GSN_API int gsn_tc_selftest(const float* a, const float* b, float* d, int* status, int N, int K,
                               int a_in_tmem, int swap_lbo_sbo, int use_fp16, gsn_stream_t stream) {
  GSN_REQUIRE(a && b && d && status, "gsn_tc_selftest: null pointer");
  GSN_REQUIRE(N >= 16 && N <= 256 && N % 16 == 0, "gsn_tc_selftest: N=%d must be a multiple of 16 in [16,256]", N);
  GSN_REQUIRE(K >= 16 && K <= 512 && K % 16 == 0, "gsn_tc_selftest: K=%d must be a multiple of 16 in [16,512]", K);
  const size_t smem = (size_t)(128 + N) * K * 2;
  GSN_REQUIRE(smem <= 200 * 1024, "gsn_tc_selftest: operands do not fit shared memory");
  GSN_CUDA(cudaFuncSetAttribute(gsn::k_tc_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  gsn::k_tc_probe<<<1, 128, smem, gsn::as_stream(stream)>>>(a, b, d, status, N, K, a_in_tmem, swap_lbo_sbo,
                                                          use_fp16);
  GSN_LAUNCH_CHECK("k_tc_probe");
  return GSN_OK;
}
