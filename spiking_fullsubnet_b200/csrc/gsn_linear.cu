// Dense fp32 linear  out[M,N] = a[M,K] @ w[N,K]^T + bias  on CUDA cores (FFMA, k ascending).
// Used for the input-to-hidden product with a REAL-valued operand (layer 0; ESN:141) where plain bf16
// tensor-core math cannot meet the 1e-3 parity bar, and for proj (MSF:118).
// 128x64 CTA tile, 16-deep k slabs, 256 threads, 8x4 register tile per thread.
#include <stdlib.h>

#include "gsn_common.cuh"

namespace gsn {

constexpr int LBN = 64, LBK = 16;  // CTA tile LBM x 64 (LBM = 128 or 64 rows), 16-deep k slabs, 2 * LBM threads

__device__ __forceinline__ float apply_act(float v, int act) {
  switch (act) {
    case 1: return tanhf(v);
    case 2: return 1.0f / (1.0f + expf(-v));
    case 3: return fmaxf(v, 0.f);
    default: return v;
  }
}

// Software pipelined: the global loads of k-slab i+1 are in flight (registers) while slab i is multiplied
// out of shared memory.  <= 96 registers so that a CTA fits beside a resident 512-thread recurrence CTA
// (wavefront schedule: the projections of chunk k+1 run while the recurrences of chunk k occupy the SMs).
// LBM = 64 (128 threads, ~12 K registers per CTA): two CTAs fit beside a resident recurrence CTA, used for the small
// chunked launches of the wavefront schedule; LBM = 128 (256 threads) for large M.  Same k-ascending FMA chain per
// output element in both, i.e. identical results.
template <int LBM>
__global__ void __launch_bounds__(LBM == 128 ? 320 : 128, LBM == 128 ? 2 : 5)  // launched with 2*LBM threads; both cap ptxas at 96 registers
    k_linear_f32(const float* __restrict__ a, const float* __restrict__ w,
                 const float* __restrict__ bias, float* __restrict__ out,
                 float* __restrict__ out_act, int act, long long M, int K, int N, TraceBuf* tb) {
  const int tslot = trace_begin(tb, 1, (int)M, K, N);
  __shared__ __align__(16) float As[LBK][LBM + 4];
  __shared__ __align__(16) float Ws[LBK][LBN + 4];
  const int tid = threadIdx.x;
  // 1-D grid-stride over the (row tile, column tile) pairs, column tile fastest: the launcher may cap the grid (see
  // gsn_subband_features) without changing any result
  const int ny = (N + LBN - 1) / LBN;
  const long long ntiles = ((M + LBM - 1) / LBM) * ny;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
  const long long m0 = (tile / ny) * LBM;
  const int n0 = (int)(tile % ny) * LBN;
  const int tm = tid >> 4;   // 0..LBM/8-1 -> rows tm*8 .. +7
  const int tn = tid & 15;   // 0..15 -> cols tn*4 .. +3
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  // loader mapping: 16 consecutive threads cover the 16 k of one row (64 contiguous bytes); one base pointer per
  // operand + a row count instead of per-row pointers (registers: a CTA must fit beside a recurrence CTA)
  const int lk = tid & 15, lr = tid >> 4;  // lr 0..LR-1
  constexpr int LR = LBM / 8;
  const float* abase = a + (m0 + lr) * K;
  const float* wbase = w + (size_t)(n0 + lr) * K;
  const int arows = (int)(M - m0 < LBM ? M - m0 : LBM) - lr;  // rows lr + LR i with LR i < arows exist
  const int wrows = (N - n0 < LBN ? N - n0 : LBN) - lr;
  float ra[LBM / LR], rw[LBN / LR];
  auto fetch = [&](int k0) {
    const int k = k0 + lk;
    const bool kv = k < K;
#pragma unroll
    for (int i = 0; i < LBM / LR; ++i) ra[i] = (kv && LR * i < arows) ? __ldg(abase + LR * i * K + k) : 0.f;
#pragma unroll
    for (int i = 0; i < LBN / LR; ++i) rw[i] = (kv && LR * i < wrows) ? __ldg(wbase + LR * i * K + k) : 0.f;
  };
  fetch(0);
  for (int k0 = 0; k0 < K; k0 += LBK) {
#pragma unroll
    for (int i = 0; i < LBM / LR; ++i) As[lk][lr + LR * i] = ra[i];
#pragma unroll
    for (int i = 0; i < LBN / LR; ++i) Ws[lk][lr + LR * i] = rw[i];
    __syncthreads();
    if (k0 + LBK < K) fetch(k0 + LBK);
#pragma unroll
    for (int kk = 0; kk < LBK; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[kk][tm * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[kk][tm * 8 + 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Ws[kk][tn * 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bv[4] = {b0.x, b0.y, b0.z, b0.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const long long m = m0 + tm * 8 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tn * 4 + j;
      if (n >= N) continue;
      const float v = acc[i][j] + (bias ? bias[n] : 0.f);
      out[m * N + n] = v;
      if (out_act) out_act[m * N + n] = apply_act(v, act);
    }
  }
  }  // tile loop
  trace_end(tb, tslot);
}

}  // namespace gsn

extern "C" int gsn_linear_f32(const float* a, const float* w, const float* bias, float* out,
                              float* out_act, int act, int64_t M, int K, int N,
                              gsn_stream_t stream) {
  GSN_REQUIRE(a && w && out, "gsn_linear_f32: null pointer");
  GSN_REQUIRE(M > 0 && K > 0 && N > 0, "gsn_linear_f32: bad shape M=%lld K=%d N=%d", (long long)M, K, N);
  GSN_REQUIRE(act >= 0 && act <= 3, "gsn_linear_f32: unknown activation %d", act);
  const int gy = (N + gsn::LBN - 1) / gsn::LBN;
  // small launches (the frame chunks of the wavefront schedule) use the 64-row tile: more, lighter CTAs
  static const long long small_m = getenv("GSN_F32_SMALL_M") ? atoll(getenv("GSN_F32_SMALL_M")) : 32768;
  const int lbm = M <= small_m ? 64 : 128;
  long long nblocks = ((M + lbm - 1) / lbm) * gy;
  GSN_REQUIRE(nblocks < 2147483647LL, "gsn_linear_f32: grid too large");
  const int cap = gsn::launch_option(GSN_OPT_F32_MAX_CTAS);
  if (cap > 0 && nblocks > cap) nblocks = cap;
  dim3 grid((unsigned)nblocks);
  if (lbm == 64)
    gsn::k_linear_f32<64><<<grid, 128, 0, gsn::as_stream(stream)>>>(a, w, bias, out, out_act, act, M, K, N,
                                                                    gsn::trace_buffer());
  else
    gsn::k_linear_f32<128><<<grid, 256, 0, gsn::as_stream(stream)>>>(a, w, bias, out, out_act, act, M, K, N,
                                                                     gsn::trace_buffer());
  GSN_LAUNCH_CHECK("k_linear_f32");
  return GSN_OK;
}
