// Hand-written sm_100a primitives used by the tcgen05 recurrence: mbarrier, cluster/DSMEM, bulk async
// copy (TMA unit, 1-D), TMEM allocation, tcgen05.mma / ld / st / commit, and the UMMA descriptors.
// Descriptor bit layouts follow the PTX ISA "tcgen05 matrix descriptor" / "instruction descriptor"
// tables (cross-checked against the CuTe headers vendored in this image).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace gsn {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// arrive on a barrier that lives in CTA `rank` of this cluster (release at cluster scope)
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* bar, uint32_t rank) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(bar)), "r"(rank));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
// Bounded spin: returns false on timeout so that a broken pipeline cannot hang the GPU box.  The bound is wall-clock
// (kWaitTimeoutNs of %globaltimer, checked every 64 Ki failed polls): kernels of the streaming pipeline legitimately wait
// for producer kernels whose launch may be delayed by tens of milliseconds (lazy module loading on first use).
constexpr unsigned long long kWaitTimeoutNs = 4000000000ull;
__device__ __forceinline__ unsigned long long wait_clock_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t a = smem_u32(bar);
  unsigned long long t0 = 0;
  for (uint32_t i = 0;; ++i) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(a), "r"(parity)
        : "memory");
    if (ok) return true;
    if ((i & 0xFFFFu) == 0xFFFFu) {
      const unsigned long long now = wait_clock_ns();
      if (t0 == 0) t0 = now;
      else if (now - t0 > kWaitTimeoutNs) return false;
    }
  }
}

// Same bounded wait with the default (CTA-scope) acquire: for barriers completed by the async proxy -- bulk copies,
// st.async complete_tx from peer CTAs, tcgen05.commit -- whose data lands in THIS CTA's shared / tensor memory.
// (The cluster-scope acquire makes the compiler emit CCTL.IVALL, an L1 invalidation, after every successful wait.)
__device__ __forceinline__ bool mbar_wait_cta(uint64_t* bar, uint32_t parity) {
  const uint32_t a = smem_u32(bar);
  unsigned long long t0 = 0;
  for (uint32_t i = 0;; ++i) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(a), "r"(parity)
        : "memory");
    if (ok) return true;
    if ((i & 0xFFFFu) == 0xFFFFu) {
      const unsigned long long now = wait_clock_ns();
      if (t0 == 0) t0 = now;
      else if (now - t0 > kWaitTimeoutNs) return false;
    }
  }
}

// ---------------------------------------------------------------- cluster / DSMEM
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cluster_nctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t map_to_rank(const void* local_smem, uint32_t rank) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(local_smem)), "r"(rank));
  return remote;
}
__device__ __forceinline__ void st_cluster_u32(uint32_t cluster_addr, uint32_t v) {
  asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(cluster_addr), "r"(v) : "memory");
}

// asynchronous 4-byte store into the shared memory of CTA `rank` of the cluster; the bytes are counted
// on that CTA's mbarrier (complete_tx), so no release fence / separate arrive is needed by the sender
__device__ __forceinline__ void st_async_u32(uint32_t remote_addr, uint32_t v, uint32_t remote_bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.u32 [%0], %1, [%2];" ::"r"(remote_addr),
               "r"(v), "r"(remote_bar)
               : "memory");
}
// 16-byte variant, and mapa on an already converted shared-window address
__device__ __forceinline__ void st_async_v4(uint32_t remote_addr, const uint32_t (&v)[4], uint32_t remote_bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(
                   remote_addr),
               "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(remote_bar)
               : "memory");
}
__device__ __forceinline__ uint32_t map_shared_rank(uint32_t local_shared_addr, uint32_t rank) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local_shared_addr), "r"(rank));
  return remote;
}
__device__ __forceinline__ bool elect_one() {
  uint32_t leader;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(leader));
  return leader != 0;
}

// ---------------------------------------------------------------- bulk async copy (TMA unit, 1-D)
// global -> this CTA's shared memory, completion counted in bytes on `bar` (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// make generic-proxy shared-memory writes visible to the async proxy (tcgen05.mma operand reads)
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ---------------------------------------------------------------- TMEM
// Every TMEM kernel of this library allocates all 512 columns, and tcgen05.alloc BLOCKS while another resident
// CTA holds them.  Requesting more than half of the SM's shared memory makes two such CTAs mutually exclusive on
// an SM, so the hardware block scheduler queues the second one instead of parking it inside the allocator.
constexpr size_t kTmemExclusiveSmem = 116 * 1024;
constexpr size_t kMaxDynamicSmem = 227 * 1024;  // opt-in dynamic shared memory per CTA on sm_100
template <uint32_t COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result) {  // one full warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
               "n"(COLS)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <uint32_t COLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {  // the allocating warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// 32 lanes x 32-bit, 16 consecutive columns: thread i of warp w gets TMEM lane 32*(w%4)+i
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, uint32_t (&v)[4]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3])
               : "r"(taddr)
               : "memory");
}
template <int N>
__device__ __forceinline__ void tmem_ld(uint32_t taddr, uint32_t (&v)[N]) {
  static_assert(N == 4 || N == 8 || N == 16, "unsupported tcgen05.ld width");
  if constexpr (N == 4) tmem_ld4(taddr, v);
  else if constexpr (N == 8) tmem_ld8(taddr, v);
  else tmem_ld16(taddr, v);
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(v[0]),
               "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}

// ---------------------------------------------------------------- UMMA descriptors
// K-major, no swizzle ("interleave"): 8 rows x 16 bytes core matrices, each 128 contiguous bytes.
//   byte(r, k) = (r/8)*SBO + (k/8)*LBO + (r%8)*16 + (k%8)*2        (16-bit elements)
// LBO = stride between the two core matrices one K=16 instruction spans, SBO = stride between 8-row groups.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);        // [0,14)  start address >> 4
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;    // [16,30) leading byte offset >> 4
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;    // [32,46) stride byte offset >> 4
  d |= 1ull << 46;                                                 // [46,48) descriptor version = 1 (sm_100)
  return d;                                                        // [61,64) layout type 0 = no swizzle
}
// kind::f16 instruction descriptor: D=f32, A=B=(bf16|f16), both K-major, dense
__device__ __forceinline__ uint32_t make_idesc_f16(uint32_t M, uint32_t N, bool bf16) {
  const uint32_t fmt = bf16 ? 1u : 0u;
  return (1u << 4) | (fmt << 7) | (fmt << 10) | ((N >> 3) << 17) | ((M >> 4) << 24);
}
// D[tmem] (+)= A[smem desc] * B[smem desc]
__device__ __forceinline__ void mma_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                       uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem desc]
__device__ __forceinline__ void mma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                       uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// ---- straight-line MMA issue --------------------------------------------------------------------------------
// A rolled issue loop (runtime trip count, operands recomputed per iteration) costs ~42-50 cycles per tcgen05.mma
// on the issuing thread -- 5x the tensor pipe's own 8 cycles for a 128 x 16 x 16 tile (tools/tc_mma_timing.py:
// 8.5 / 13.6 / 25.4 cycles per MMA at N = 16 / 32 / 64 when every operand is base + immediate).  So the k loop is
// fully unrolled for a compile-time number of k steps and selected by a switch on the runtime value.
template <int ACC>
__device__ __forceinline__ void mma_ts_c(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "n"(ACC)
      : "memory");
}
// D = sum over PLANES planes (plane pl at tmem_a + pl * KS * 8 columns) and KS k steps of A[plane][ks] * B[ks];
// the first MMA overwrites D
template <int KS, int PLANES>
__device__ __forceinline__ void mma_planes_unrolled(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b0,
                                                    uint32_t idesc) {
  mma_ts_c<0>(tmem_d, tmem_a, desc_b0, idesc);
#pragma unroll
  for (int i = 1; i < KS * PLANES; ++i) {
    const int pl = i / KS, ks = i % KS;
    mma_ts_c<1>(tmem_d, tmem_a + (uint32_t)((pl * KS + ks) * 8), desc_b0 + (uint64_t)(ks * 16), idesc);
  }
}
// runtime dispatch: ksteps = Kmma / 16 in [1, 20]  (Kmma <= 320).  Returns false for an unsupported count.
template <int PLANES>
__device__ __forceinline__ bool mma_planes(int ksteps, uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b0,
                                           uint32_t idesc) {
  switch (ksteps) {
#define GSN_KS_CASE(n) case n: mma_planes_unrolled<n, PLANES>(tmem_d, tmem_a, desc_b0, idesc); return true;
    GSN_KS_CASE(1) GSN_KS_CASE(2) GSN_KS_CASE(3) GSN_KS_CASE(4) GSN_KS_CASE(5) GSN_KS_CASE(6) GSN_KS_CASE(7)
    GSN_KS_CASE(8) GSN_KS_CASE(9) GSN_KS_CASE(10) GSN_KS_CASE(11) GSN_KS_CASE(12) GSN_KS_CASE(13) GSN_KS_CASE(14)
    GSN_KS_CASE(15) GSN_KS_CASE(16) GSN_KS_CASE(17) GSN_KS_CASE(18) GSN_KS_CASE(19) GSN_KS_CASE(20)
#undef GSN_KS_CASE
    default: return false;
  }
}

// ---- bf16x3 x bf16x3 ("both operands real"): PAIRS of the 9 plane pairs x KS k steps, smallest terms first.
// PAIRS = 8 drops only lo x lo (<= 2^-32 |w||x|); PAIRS = 6 also drops lo x mid and mid x lo (<= 2^-23 |w||x| together:
// the rounding level of an fp32 product).  w plane pw at tmem_a + (pw*KS + ks)*8 columns, x plane px at
// desc_b0 + px*PB + ks*16 (PB = plane bytes >> 4 = NT*KS*2).  Plane index: 0 = lo, 1 = mid, 2 = hi.  The first MMA
// overwrites D.
template <int KS, int PB, int PAIRS>
__device__ __forceinline__ void mma_pairs_unrolled(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b0, uint32_t idesc) {
  constexpr int PW[8] = {0, 1, 0, 2, 1, 1, 2, 2};
  constexpr int PX[8] = {1, 0, 2, 0, 1, 2, 1, 2};
  constexpr int T0 = 8 - PAIRS;
  mma_ts_c<0>(tmem_d, tmem_a + (uint32_t)((PW[T0] * KS) * 8), desc_b0 + (uint64_t)(PX[T0] * PB), idesc);
#pragma unroll
  for (int i = T0 * KS + 1; i < 8 * KS; ++i) {
    const int term = i / KS, ks = i % KS;
    mma_ts_c<1>(tmem_d, tmem_a + (uint32_t)((PW[term] * KS + ks) * 8), desc_b0 + (uint64_t)(PX[term] * PB + ks * 16),
                idesc);
  }
}
template <int NT, int PAIRS = 8>
__device__ __forceinline__ bool mma_pairs(int ksteps, uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b0, uint32_t idesc) {
  switch (ksteps) {
#define GSN_KS_CASE(n) case n: mma_pairs_unrolled<n, NT * n * 2, PAIRS>(tmem_d, tmem_a, desc_b0, idesc); return true;
    GSN_KS_CASE(1) GSN_KS_CASE(2) GSN_KS_CASE(3) GSN_KS_CASE(4) GSN_KS_CASE(5) GSN_KS_CASE(6) GSN_KS_CASE(7)
    GSN_KS_CASE(8) GSN_KS_CASE(9) GSN_KS_CASE(10) GSN_KS_CASE(11) GSN_KS_CASE(12) GSN_KS_CASE(13) GSN_KS_CASE(14)
    GSN_KS_CASE(15) GSN_KS_CASE(16)
#undef GSN_KS_CASE
    default: return false;
  }
}

// kind::i8: D(int32)[tmem] (+)= A(int8/uint8)[tmem] * B(int8/uint8)[smem desc], K = 32 per instruction
__device__ __forceinline__ uint32_t make_idesc_i8(uint32_t M, uint32_t N, bool a_signed, bool b_signed) {
  return (2u << 4) | ((a_signed ? 1u : 0u) << 7) | ((b_signed ? 1u : 0u) << 10) | ((N >> 3) << 17) | ((M >> 4) << 24);
}
__device__ __forceinline__ void mma_i8_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// straight-line issue for kind::i8 (see mma_planes_unrolled): plane pl accumulates into its OWN int32 accumulator
// tmem_d + pl * d_stride; the top plane is signed (idesc_s), the others unsigned digits (idesc_u)
template <int ACC>
__device__ __forceinline__ void mma_i8_ts_c(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "n"(ACC)
      : "memory");
}
template <int KS, int PLANES>
__device__ __forceinline__ void mma_i8_planes_unrolled(uint32_t tmem_d, uint32_t d_stride, uint32_t tmem_a,
                                                       uint64_t desc_b0, uint32_t idesc_u, uint32_t idesc_s) {
#pragma unroll
  for (int pl = 0; pl < PLANES; ++pl) {
    const uint32_t idesc = pl == PLANES - 1 ? idesc_s : idesc_u;
    mma_i8_ts_c<0>(tmem_d + pl * d_stride, tmem_a + (uint32_t)(pl * KS * 8), desc_b0, idesc);
#pragma unroll
    for (int ks = 1; ks < KS; ++ks)
      mma_i8_ts_c<1>(tmem_d + pl * d_stride, tmem_a + (uint32_t)((pl * KS + ks) * 8), desc_b0 + (uint64_t)(ks * 16),
                     idesc);
  }
}
// runtime dispatch: ksteps = Kmma / 32 in [1, 16]  (Kmma <= 512)
template <int PLANES>
__device__ __forceinline__ bool mma_i8_planes(int ksteps, uint32_t tmem_d, uint32_t d_stride, uint32_t tmem_a,
                                              uint64_t desc_b0, uint32_t idesc_u, uint32_t idesc_s) {
  switch (ksteps) {
#define GSN_KS_CASE(n) \
  case n: mma_i8_planes_unrolled<n, PLANES>(tmem_d, d_stride, tmem_a, desc_b0, idesc_u, idesc_s); return true;
    GSN_KS_CASE(1) GSN_KS_CASE(2) GSN_KS_CASE(3) GSN_KS_CASE(4) GSN_KS_CASE(5) GSN_KS_CASE(6) GSN_KS_CASE(7)
    GSN_KS_CASE(8) GSN_KS_CASE(9) GSN_KS_CASE(10) GSN_KS_CASE(11) GSN_KS_CASE(12) GSN_KS_CASE(13) GSN_KS_CASE(14)
    GSN_KS_CASE(15) GSN_KS_CASE(16)
#undef GSN_KS_CASE
    default: return false;
  }
}

// 8 spike bits -> 8 bf16 {0, 1} (one 16-byte operand chunk) without a table: x * 0x10204080 moves bit i of a nibble to
// bit 8i + 7 (the 16 partial products land on distinct bits, so nothing carries), PRMT replicates those byte sign bits
// over half-words, and the mask leaves 0x3F80 = bf16 1.0 where the bit was set.
__device__ __forceinline__ uint4 spike_byte_to_bf16x8(uint32_t b8) {
  const uint32_t r0 = (b8 & 0xFu) * 0x10204080u, r1 = ((b8 >> 4) & 0xFu) * 0x10204080u;
  uint4 v;  // (prmt.b32 directly: __byte_perm drops the sign-replication bit of the selector nibbles)
  asm("prmt.b32 %0, %1, 0, 0x9988;" : "=r"(v.x) : "r"(r0));
  asm("prmt.b32 %0, %1, 0, 0xBBAA;" : "=r"(v.y) : "r"(r0));
  asm("prmt.b32 %0, %1, 0, 0x9988;" : "=r"(v.z) : "r"(r1));
  asm("prmt.b32 %0, %1, 0, 0xBBAA;" : "=r"(v.w) : "r"(r1));
  v.x &= 0x3F803F80u; v.y &= 0x3F803F80u; v.z &= 0x3F803F80u; v.w &= 0x3F803F80u;
  return v;
}

// all previously issued MMAs of this thread arrive (once) on `bar` when complete
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

}  // namespace tc
}  // namespace gsn
