// Streaming helper stages of the frame-granular pipeline on tcgen05 (one warp-specialised persistent kernel, two
// producers):
//   GATHER -- the front end of a sequence model: sub-band gather (MSF:241-312) + pre-LayerNorm
//             (MSF:111-112) + the layer-0 input-to-hidden product x_t . W_ih^T (ESN:141).  It follows the full-band
//             model frame by frame (the sub-band input of frame t needs the full-band output of frame t, MSF:441-447)
//             and feeds the layer-0 streaming recurrence through per-frame counters.
//   BITS   -- a linear layer whose left operand is a bit-packed spike trace a concurrently running recurrence emits:
//             the input-to-hidden product of layers >= 1 that do not fit the fused recurrence, and proj (MSF:118).
//
// Arithmetic.  BITS: spikes are exact in bf16, the fp32 weights are three exact bf16 planes (w = hi + mid + lo), every
// product is exact and the three planes accumulate in fp32, lo first (as gsn_linear_tc.cu).  GATHER: x is REAL valued,
// so BOTH operands are split by truncation into three bf16 planes; every bf16 x bf16 product is exact in fp32 and 8 of
// the 9 plane pairs are accumulated in tensor memory, smallest magnitude first (only lo x lo, <= 2^-32 |w||x|, is
// dropped): an fp32-faithful dot product whose error is below that of an fp32 FMA chain.
//
// Structure.  CTA (slice, p) owns output features [128 slice, +128) -- their weight planes stay in tensor memory as the
// A operand -- and walks over the row tiles p, p+P, ... of the flattened [T*R] rows.  Warp roles:
//   9 producer warps : warp w builds the B operand (K-major, no swizzle) of the tiles w, w+9, ... of this CTA, each in
//                      its own stage of a ring of NS <= 9 shared-memory stages, after an acquire poll of in_cnt for
//                      the frames the tile touches.  A lane owns (row n%8 of an 8-row group, 8 consecutive features):
//                      its loads are issued together, the LayerNorm reductions need two shuffles, the bf16 planes
//                      leave as 16-byte stores of which 8 lanes fill one 128-byte core matrix (conflict-free); the L2
//                      latency of one warp is covered by the eight others working on later tiles;
//   1 issue warp     : tcgen05.mma of tile i into accumulator i&1 as soon as its stage is full; tcgen05.commit frees
//                      the stage and hands the accumulator to
//   8 epilogue warps : tcgen05.ld, bias / activation, coalesced stores;
//   2 publisher warps: one gpu-scope release fence + one add per touched frame on out_cnt, tiles alternating between
//                      the two warps (a release costs an L2 round trip: it must neither sit on the epilogue's path
//                      nor be serialised tile after tile).
// All hand-overs are mbarriers; there is no __syncthreads in the tile loop.
#include <stdlib.h>

#include "gsn_common.cuh"
#include "gsn_tc.cuh"

namespace gsn {

struct StageParams {
  // GATHER
  const float* cm;      // [T, B, f_cm] compressed magnitude
  const float* fb;      // [T, B, f_fb] full-band output or null
  const float* ln_w;    // [K] or null
  const float* ln_b;
  float* x_out;         // [T, R, K] normalised input (all_layer_outputs[0]) or null
  int B, N, lo, ctr, nbr, f_cm, f_fb;
  float eps;
  // BITS
  const uint32_t* a_bits;  // [T*R, ceil(K/32)]
  // common
  const float* w;       // [Nout, K]
  const float* bias;    // [Nout] or null
  float* out;           // [T*R, Nout]
  float* out_act;       // or null
  int act;
  const unsigned int* in_cnt;  // [T] or null
  unsigned int in_target;
  unsigned int* out_cnt;       // [T] or null: += rows per (tile, slice); frame complete at R * slices
  int T, R, K, Kmma, Nout;
  int wpitch, ns;
  int pitch;  // GATHER: scratch row pitch in floats (>= Kmma, pitch % 8 == 4)
  unsigned int poll_ns;
  TraceBuf* trace;
};

constexpr int kSgEpiWarps = 8, kSgProdWarps = 9;  // 20 warps: 5 per SM sub-partition, 96 registers each
constexpr int kSgIssueWarp = kSgEpiWarps + kSgProdWarps;
constexpr int kSgPubWarps = 2, kSgPubRing = 4;
constexpr int kSgThreads = (kSgIssueWarp + 1 + kSgPubWarps) * 32;
constexpr int kSgMaxStages = 9;
enum { kStageBits = 0, kStageGather = 1 };

__device__ __forceinline__ float sg_act(float v, int act) {
  switch (act) {
    case 1: return tanhf(v);
    case 2: return 1.0f / (1.0f + expf(-v));
    case 3: return fmaxf(v, 0.f);
    default: return v;
  }
}

__device__ __forceinline__ void sg_split3(float w, uint32_t& hi, uint32_t& mid, uint32_t& lo) {
  const uint32_t wb = __float_as_uint(w);
  hi = wb >> 16;
  const float r1 = w - __uint_as_float(wb & 0xFFFF0000u);
  const uint32_t r1b = __float_as_uint(r1);
  mid = r1b >> 16;
  const float r2 = r1 - __uint_as_float(r1b & 0xFFFF0000u);
  lo = __float_as_uint(r2) >> 16;
}

__device__ __forceinline__ uint32_t sg_ld_cg(const void* p) {
  uint32_t v;
  asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// Frame counters of a concurrently running producer kernel: frames [0, ready) are known complete; polls the 32 frames
// from `ready` on (one acquire load per lane) until frame `need` is complete.  Bounded: false on timeout.
__device__ __forceinline__ bool sg_poll_frames(const unsigned int* cnt, unsigned int target, int T, int& ready, int need,
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

static inline __host__ __device__ int sg_wpitch(int K) { return ((K / 4) & 1) ? K : K + 4; }

template <int NT, int MODE, int J>
__global__ void __launch_bounds__(kSgThreads, 1) k_stage_stream(const StageParams p) {
  constexpr int HC = NT / 2;               // accumulator columns (rows of the tile) per epilogue warp
  constexpr int CH = HC < 16 ? HC : 16;
  constexpr int NPLANES = MODE == kStageGather ? 3 : 1;
  static_assert(NT == 16 || NT == 32 || NT == 64, "NT must be 16, 32 or 64");
  extern __shared__ __align__(1024) uint8_t smem[];
  const int tslot = trace_begin(p.trace, MODE == kStageGather ? 6 : 4, p.T, p.R, p.K);
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
  const int slice = blockIdx.x, P = gridDim.y;
  const int K = p.K, Kmma = p.Kmma, Nout = p.Nout, R = p.R, NS = p.ns;
  const int M = p.T * R;
  const int ntiles_all = (M + NT - 1) / NT;

  const uint32_t SBO = 16u * Kmma;
  const size_t plane_bytes = (size_t)NT * Kmma * 2;  // multiple of 512
  const size_t stage_bytes = NPLANES * plane_bytes;
  const size_t wst_bytes = p.wpitch > 0 ? (size_t)128 * p.wpitch * 4 : 0;
  const size_t ring_bytes = (size_t)NS * stage_bytes;
  const size_t ln_off = ((ring_bytes > wst_bytes ? ring_bytes : wst_bytes) + 127) / 128 * 128;
  float* s_lnw = reinterpret_cast<float*>(smem + ln_off);   // [Kmma] LayerNorm weight (1 when absent), bias behind it
  float* s_lnb = s_lnw + Kmma;
  float* s_scr = s_lnb + Kmma;                               // [NS][8][pitch] per-warp transpose scratch
  const size_t bar_off = ln_off + (MODE == kStageGather ? ((size_t)2 * Kmma + (size_t)NS * 8 * p.pitch) * 4 : 0);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + bar_off);
  uint64_t* bar_full = bars;             // [12] stage filled by its producer warp
  uint64_t* bar_free = bars + 12;        // [12] stage consumed (tcgen05.commit)
  uint64_t* bar_dfull = bars + 24;       // [2] accumulator complete (tcgen05.commit)
  uint64_t* bar_dfree = bars + 26;       // [2] accumulator read by the 8 epilogue warps
  uint64_t* bar_w = bars + 28;
  uint64_t* bar_pub = bars + 29;         // [4] tile stored by the 8 epilogue warps
  uint64_t* bar_pfree = bars + 33;       // [4] tile published
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 37);
  float* wst = reinterpret_cast<float*>(smem);

  if (tid == 0) {
    for (int i = 0; i < kSgMaxStages; ++i) {
      tc::mbar_init(&bar_full[i], 1);
      tc::mbar_init(&bar_free[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      tc::mbar_init(&bar_dfull[i], 1);
      tc::mbar_init(&bar_dfree[i], kSgEpiWarps);
    }
    tc::mbar_init(bar_w, 1);
    for (int i = 0; i < kSgPubRing; ++i) {
      tc::mbar_init(&bar_pub[i], kSgEpiWarps);
      tc::mbar_init(&bar_pfree[i], 1);
    }
    tc::fence_mbar_init();
    if (p.wpitch > 0) {
      const int nrows = Nout - slice * 128 < 128 ? Nout - slice * 128 : 128;
      tc::mbar_arrive_expect_tx(bar_w, (uint32_t)nrows * (uint32_t)K * 4u);
    }
  }
  if (MODE == kStageGather)
    for (int k = tid; k < Kmma; k += kSgThreads) {
      s_lnw[k] = (p.ln_w != nullptr && k < K) ? p.ln_w[k] : 1.f;
      s_lnb[k] = (p.ln_w != nullptr && k < K) ? p.ln_b[k] : 0.f;
    }
  if (warp == 0) tc::tmem_alloc<512>(tmem_slot);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tmem_d0 = tmem, tmem_d1 = tmem + NT;
  const uint32_t tmem_a = tmem + 2 * NT;
  const uint32_t plane_cols = Kmma / 2;

  // ---- prologue (warps 0..15): the 128 weight rows of this slice -> three exact bf16 planes in tensor memory --------
  if (warp < 16) {
    const int q = warp & 3, g = warp >> 2;
    const int tl = q * 32 + lane;
    const int j = slice * 128 + tl;
    const bool jv = j < Nout;
    const float* wrow = p.w + (size_t)(jv ? j : 0) * K;
    const bool staged = p.wpitch > 0;
    if (staged && g == 0 && jv) tc::bulk_g2s(wst + (size_t)tl * p.wpitch, wrow, (uint32_t)K * 4u, bar_w);
    if (staged && !tc::mbar_wait_cta(bar_w, 0)) __trap();
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    const float* srow = wst + (size_t)tl * p.wpitch;
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
        for (int e = 0; e < 2; ++e) sg_split3(wv[2 * u + e], h2[e], m2[e], l2[e]);
        vh[u] = h2[0] | (h2[1] << 16);
        vm[u] = m2[0] | (m2[1] << 16);
        vl[u] = l2[0] | (l2[1] << 16);
      }
      tc::tmem_st8(tmem_a + lane_base + 0 * plane_cols + c0, vl);
      tc::tmem_st8(tmem_a + lane_base + 1 * plane_cols + c0, vm);
      tc::tmem_st8(tmem_a + lane_base + 2 * plane_cols + c0, vh);
    }
    tc::tmem_wait_st();
  }
  tc::tc_fence_before();
  __syncthreads();  // weights in place; the staging bytes become the operand ring
  tc::tc_fence_after();

  const int first = blockIdx.y;
  const int my_tiles = first < ntiles_all ? (ntiles_all - first + P - 1) / P : 0;

  if (warp == kSgIssueWarp) {
    // =============================== MMA issue warp ===============================
    const bool leader = tc::elect_one();
    const uint32_t idesc = tc::make_idesc_f16(128, NT, true);
    const int ksteps = Kmma / 16;
    for (int i = 0; i < my_tiles; ++i) {
      const int slot = i % NS, buf = i & 1;
      if (!tc::mbar_wait_cta(&bar_full[slot], (uint32_t)((i / NS) & 1))) __trap();
      // Issue tile i only after the epilogue has READ tile i-1 (which also frees accumulator i&1, read before it): a
      // tcgen05.ld queued behind in-flight MMAs waits for all of them, so letting the MMAs of tile i run ahead would
      // serialise the pipe the other way round (MMA i -> ld i-1 -> MMA i+1) at twice the cost
      if (i >= 1 && !tc::mbar_wait_cta(&bar_dfree[(i - 1) & 1], (uint32_t)(((i - 1) >> 1) & 1))) __trap();
      tc::tc_fence_after();
      const uint64_t desc_b0 = tc::make_smem_desc(tc::smem_u32(smem + (size_t)slot * stage_bytes), 128, SBO);
      if (leader) {
        bool ok;
        if (MODE == kStageGather) ok = tc::mma_pairs<NT>(ksteps, buf ? tmem_d1 : tmem_d0, tmem_a, desc_b0, idesc);
        else ok = tc::mma_planes<3>(ksteps, buf ? tmem_d1 : tmem_d0, tmem_a, desc_b0, idesc);
        if (!ok) __trap();
        tc::mma_commit(&bar_free[slot]);
        tc::mma_commit(&bar_dfull[buf]);
      }
      __syncwarp();
    }
  } else if (warp > kSgIssueWarp) {
    // =============================== publisher warps ===============================
    if (lane == 0 && p.out_cnt != nullptr) {
      for (int i = warp - kSgIssueWarp - 1; i < my_tiles; i += kSgPubWarps) {
        const int tile = first + i * P, ps = i % kSgPubRing;
        if (!tc::mbar_wait_cta(&bar_pub[ps], (uint32_t)((i / kSgPubRing) & 1))) __trap();
        asm volatile("fence.acq_rel.gpu;" ::: "memory");  // the epilogue warps' stores, observed through the mbarrier
        const int ra = tile * NT;
        const int rb = ra + NT < M ? ra + NT : M;
        for (int t = ra / R; t * R < rb; ++t) {
          const int a = t * R > ra ? t * R : ra, bnd = (t + 1) * R < rb ? (t + 1) * R : rb;
          asm volatile("red.relaxed.gpu.global.add.u32 [%0], %1;" ::"l"(p.out_cnt + t), "r"((unsigned int)(bnd - a))
                       : "memory");
        }
        tc::mbar_arrive(&bar_pfree[ps]);
      }
    }
    __syncwarp();
  } else if (warp >= kSgEpiWarps) {
    // =============================== producer warps ===============================
    const int pw = warp - kSgEpiWarps;
    const int rg = lane & 7, cg = lane >> 3;  // row of an 8-row group / chunk group
    int ready = 0;  // frames [0, ready) of the input are known complete
    // GATHER: per-lane constants of feature jx = lane + 32 i: from the noisy band or the full-band output, inside K,
    // and (jx - k_noisy) mod f_fb
    bool g_noisy[J], g_valid[J];
    int g_off[J];
    const float* g_base[J];
    int lo_mod = 0;
    if (MODE == kStageGather) {
      const int kn = p.ctr + 2 * p.nbr;
#pragma unroll
      for (int i = 0; i < J; ++i) {
        const int jx = lane + 32 * i;
        g_noisy[i] = jx < kn;
        g_valid[i] = jx < K;
        g_off[i] = (!g_noisy[i] && p.f_fb > 0) ? (jx - kn) % p.f_fb : 0;
        g_base[i] = (g_noisy[i] || p.fb == nullptr) ? p.cm : p.fb;
      }
      lo_mod = p.f_fb > 0 ? p.lo % p.f_fb : 0;
    }
    // warp pw owns stage pw (NS <= 9 active producer warps): every mbarrier wait below is for the NEXT phase of a
    // barrier only this warp waits on -- a parity wait two phases ahead would pass spuriously
    for (int i = pw; i < my_tiles && pw < NS; i += NS) {
      const int tile = first + i * P, slot = pw;
      if (i >= NS && !tc::mbar_wait_cta(&bar_free[slot], (uint32_t)(((i / NS) - 1) & 1))) __trap();
      uint8_t* dst = smem + (size_t)slot * stage_bytes;
      const int m0 = tile * NT;  // first (flattened) row of the tile
      if (p.in_cnt != nullptr) {
        const int ml = m0 + NT - 1 < M ? m0 + NT - 1 : M - 1;
        const int t_hi = ml / R;
        if (ready <= t_hi && !sg_poll_frames(p.in_cnt, p.in_target, p.T, ready, t_hi, lane, p.poll_ns)) __trap();
      }
      if (MODE == kStageBits) {
        // task = (row, one 32-bit word of its packed trace) -> up to four 16-byte operand chunks; lanes run over the
        // rows of an 8-row group first, and every word of the tile is loaded before any is expanded
        const int k8n = Kmma / 8, nw = (k8n + 3) / 4, Wk = (K + 31) / 32;
        constexpr int MAXW = (NT * 10 + 31) / 32;  // Kmma <= 320: at most 10 words per row
        uint32_t wd[MAXW];
#pragma unroll
        for (int it = 0; it < MAXW; ++it) {
          const int idx = lane + 32 * it;
          const int nlo = idx & 7, wi = (idx >> 3) % nw, nhi = (idx >> 3) / nw;
          const int row = m0 + nhi * 8 + nlo;
          wd[it] = (idx < NT * nw && wi < Wk && row < M) ? sg_ld_cg(p.a_bits + (size_t)row * Wk + wi) : 0u;
        }
#pragma unroll
        for (int it = 0; it < MAXW; ++it) {
          const int idx = lane + 32 * it;
          if (idx >= NT * nw) break;
          const int nlo = idx & 7, wi = (idx >> 3) % nw, nhi = (idx >> 3) / nw;
          const uint32_t rowoff = (uint32_t)(nhi * SBO + nlo * 16);
#pragma unroll
          for (int e4 = 0; e4 < 4; ++e4) {
            const int k8 = 4 * wi + e4;
            if (k8 >= k8n) break;
            *reinterpret_cast<uint4*>(dst + rowoff + (uint32_t)(k8 * 128)) = tc::spike_byte_to_bf16x8(wd[it] >> (8 * e4));
          }
        }
      } else {
        // gather + LayerNorm + split, 8 rows per pass.  (1) the raw features of the 8 rows are loaded with lanes running
        // over the features (coalesced: 4 sectors per load) into a per-warp scratch [8][pitch]; (2) each lane reads
        // back (row rg, chunks cg, cg+4, ... of 8 features) as 16-byte words (pitch % 8 == 4: conflict-free) and does
        // the LayerNorm, the split and the operand stores in that form
        const int k_noisy = p.ctr + 2 * p.nbr;
        const int k8n = Kmma / 8;
        const bool use_ln = p.ln_w != nullptr;
        const float inv_k = 1.0f / (float)K;
        float* scr = s_scr + (size_t)pw * 8 * p.pitch;
        // row state of the walk over the tile's rows (warp-uniform): m = (t*B + b)*N + ns; no division per row
        int rt = m0 / R, rr = m0 - rt * R, rb = rr / p.N, rn = rr - rb * p.N;
        int bmod = p.f_fb > 0 ? (p.lo + rn * p.ctr) % p.f_fb : 0;  // (lo + ns*ctr) mod f_fb; ctr <= f_fb
#pragma unroll 1
        for (int it = 0; it < NT / 8; ++it) {
          {
            float raw[8][J];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              const bool rvu = m0 + it * 8 + u < M;
              // 32-bit element offsets from the two base pointers (the tensors hold < 2^31 elements)
              const int tb = rt * p.B + rb;
              const int row_cm = tb * p.f_cm, row_fb = tb * p.f_fb;
              const int q0 = p.lo + rn * p.ctr - p.nbr + lane;
#pragma unroll
              for (int i = 0; i < J; ++i) {
                int qq = q0 + 32 * i;                       // reflect padding at both ends of the spectrum (MSF:262)
                qq = qq < 0 ? -qq : qq;
                qq = min(qq, 2 * (p.f_cm - 1) - qq);
                int fi = bmod + g_off[i];
                fi = fi >= p.f_fb ? fi - p.f_fb : fi;
                const int off = g_noisy[i] ? row_cm + qq : row_fb + fi;
                raw[u][i] = (rvu && g_valid[i]) ? __uint_as_float(sg_ld_cg(g_base[i] + off)) : 0.f;
              }
              // next row
              ++rr; ++rn;
              bmod += p.ctr;
              bmod = bmod >= p.f_fb ? bmod - p.f_fb : bmod;
              if (rn == p.N) { rn = 0; ++rb; bmod = lo_mod; }
              if (rr == R) { rr = 0; ++rt; rb = 0; }
            }
#pragma unroll
            for (int u = 0; u < 8; ++u)
#pragma unroll
              for (int i = 0; i < J; ++i)
                if (lane + 32 * i < Kmma) scr[u * p.pitch + lane + 32 * i] = raw[u][i];
          }
          __syncwarp();
          const int n = it * 8 + rg;
          const int m = m0 + n;
          const bool rv = m < M;
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
          __syncwarp();  // the scratch may be overwritten by the next pass
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
          float* xo = (p.x_out != nullptr && rv && slice == 0) ? p.x_out + (size_t)m * K : nullptr;
          const uint32_t rowoff = (uint32_t)(it * SBO + rg * 16);
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
              for (int u = 0; u < 4; ++u) {
                const uint32_t a0 = __float_as_uint(v[jc][2 * u]), a1 = __float_as_uint(v[jc][2 * u + 1]);
                wh[u] = __byte_perm(a0, a1, 0x7632);
                const float r0 = v[jc][2 * u] - __uint_as_float(a0 & 0xFFFF0000u);
                const float r1 = v[jc][2 * u + 1] - __uint_as_float(a1 & 0xFFFF0000u);
                const uint32_t c0 = __float_as_uint(r0), c1 = __float_as_uint(r1);
                wm[u] = __byte_perm(c0, c1, 0x7632);
                const float q0 = r0 - __uint_as_float(c0 & 0xFFFF0000u);
                const float q1 = r1 - __uint_as_float(c1 & 0xFFFF0000u);
                wl[u] = __byte_perm(__float_as_uint(q0), __float_as_uint(q1), 0x7632);
              }
              uint8_t* d0 = dst + rowoff + (uint32_t)(ch * 128);
              *reinterpret_cast<uint4*>(d0) = make_uint4(wl[0], wl[1], wl[2], wl[3]);
              *reinterpret_cast<uint4*>(d0 + plane_bytes) = make_uint4(wm[0], wm[1], wm[2], wm[3]);
              *reinterpret_cast<uint4*>(d0 + 2 * plane_bytes) = make_uint4(wh[0], wh[1], wh[2], wh[3]);
            }
          }
        }
      }
      tc::fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&bar_full[slot]);
    }
  } else {
    // =============================== epilogue warps ===============================
    const int q = warp & 3, half = warp >> 2;
    const int j = slice * 128 + q * 32 + lane;
    const bool jv = j < Nout;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    const float bj = (p.bias != nullptr && jv) ? p.bias[j] : 0.f;
    for (int i = 0; i < my_tiles; ++i) {
      const int tile = first + i * P, buf = i & 1;
      if (!tc::mbar_wait_cta(&bar_dfull[buf], (uint32_t)((i >> 1) & 1))) __trap();
      tc::tc_fence_after();
      uint32_t zr[HC / CH][CH];
#pragma unroll
      for (int c = 0; c < HC / CH; ++c) tc::tmem_ld<CH>((buf ? tmem_d1 : tmem_d0) + lane_base + half * HC + c * CH, zr[c]);
      tc::tmem_wait_ld();
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&bar_dfree[buf]);
      // stores: thread = feature j, rows [half*HC, +HC) of the tile; coalesced over features
      const int r0 = tile * NT + half * HC;
      const int left = M - r0;
      if (jv && left > 0) {
        float* po = p.out + (size_t)r0 * Nout + j;
        if (left >= HC) {
#pragma unroll
          for (int u = 0; u < HC; ++u) po[(size_t)u * Nout] = __uint_as_float(zr[u / CH][u % CH]) + bj;
        } else {
#pragma unroll
          for (int u = 0; u < HC; ++u)
            if (u < left) po[(size_t)u * Nout] = __uint_as_float(zr[u / CH][u % CH]) + bj;
        }
        if (p.out_act != nullptr) {
          float* pa = p.out_act + (size_t)r0 * Nout + j;
#pragma unroll
          for (int u = 0; u < HC; ++u)
            if (u < left) pa[(size_t)u * Nout] = sg_act(__uint_as_float(zr[u / CH][u % CH]) + bj, p.act);
        }
      }
      if (p.out_cnt != nullptr) {
        __syncwarp();
        if (lane == 0) {
          const int ps = i % kSgPubRing;
          if (i >= kSgPubRing && !tc::mbar_wait_cta(&bar_pfree[ps], (uint32_t)(((i / kSgPubRing) - 1) & 1))) __trap();
          tc::mbar_arrive(&bar_pub[ps]);
        }
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc<512>(tmem);
  trace_end(p.trace, tslot);
}

template <int NT, int MODE, int J>
static int launch_stage(StageParams p, int ctas_per_slice, cudaStream_t st) {
  const size_t stage = (size_t)(MODE == kStageGather ? 3 : 1) * NT * p.Kmma * 2;
  p.pitch = p.Kmma + 4;  // Kmma % 16 == 0, so pitch % 8 == 4
  // per ring stage: the operand planes + (GATHER) the producer warp's transpose scratch; fixed: LayerNorm params, barriers
  const size_t per_stage = stage + (MODE == kStageGather ? (size_t)8 * p.pitch * 4 : 0);
  const size_t extra = (MODE == kStageGather ? (size_t)2 * p.Kmma * 4 : 0) + 512;
  p.wpitch = 0;
  size_t wst = 0;
  if (p.K % 4 == 0 && (reinterpret_cast<uintptr_t>(p.w) & 15) == 0) {
    p.wpitch = sg_wpitch(p.K);
    wst = (size_t)128 * p.wpitch * sizeof(float);
    if (wst + extra + 128 > tc::kMaxDynamicSmem) { p.wpitch = 0; wst = 0; }
  }
  int ns = (int)((tc::kMaxDynamicSmem - extra - 128) / per_stage);
  if (ns > kSgMaxStages) ns = kSgMaxStages;
  if (ns < 2) return fail(GSN_ENOSUP, "gsn stage stream: K=%d does not fit shared memory", p.K);
  p.ns = ns;
  static const unsigned int poll_ns = getenv("GSN_POLL_NS") ? (unsigned int)atoi(getenv("GSN_POLL_NS")) : 100u;
  p.poll_ns = poll_ns;
  const size_t ring = (size_t)ns * stage;
  size_t smem = ((ring > wst ? ring : wst) + 127) / 128 * 128 + extra + (per_stage - stage) * ns;
  if (smem < tc::kTmemExclusiveSmem) smem = tc::kTmemExclusiveSmem;
  GSN_CUDA(cudaFuncSetAttribute(k_stage_stream<NT, MODE, J>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int slices = (p.Nout + 127) / 128;
  const long long ntiles = ((long long)p.T * p.R + NT - 1) / NT;
  long long P = ctas_per_slice < 1 ? 1 : ctas_per_slice;
  if (P > ntiles) P = ntiles;
  dim3 grid((unsigned)slices, (unsigned)P);
  k_stage_stream<NT, MODE, J><<<grid, kSgThreads, smem, st>>>(p);
  GSN_LAUNCH_CHECK("k_stage_stream");
  return GSN_OK;
}

// row tile: 64 rows when at least 6 ring stages of that size fit (and tensor memory holds the planes), else 32 / 16
template <int MODE, int J>
static int dispatch_stage(const StageParams& p, int ctas_per_slice, cudaStream_t st) {
  const size_t per_row = (size_t)(MODE == kStageGather ? 3 : 1) * p.Kmma * 2;
  const size_t room = tc::kMaxDynamicSmem - 4096 - (MODE == kStageGather ? (size_t)6 * 8 * (p.Kmma + 4) * 4 : 0);
  const int acols = 3 * (p.Kmma / 2);
  if (acols + 2 * 64 <= 512 && 6 * 64 * per_row <= room) return launch_stage<64, MODE, J>(p, ctas_per_slice, st);
  if (acols + 2 * 32 <= 512 && 4 * 32 * per_row <= room) return launch_stage<32, MODE, J>(p, ctas_per_slice, st);
  if (acols + 2 * 16 <= 512) return launch_stage<16, MODE, J>(p, ctas_per_slice, st);
  return fail(GSN_ENOSUP, "gsn stage stream: K=%d does not fit tensor memory (K <= 320)", p.K);
}

int preload_stage_stream() {
  cudaFuncAttributes a;
#define GSN_PRE(NT, MODE, J) GSN_CUDA(cudaFuncGetAttributes(&a, k_stage_stream<NT, MODE, J>));
#define GSN_PRE3(MODE, J) GSN_PRE(64, MODE, J) GSN_PRE(32, MODE, J) GSN_PRE(16, MODE, J)
  GSN_PRE3(kStageBits, 1) GSN_PRE3(kStageGather, 2) GSN_PRE3(kStageGather, 3) GSN_PRE3(kStageGather, 5)
  GSN_PRE3(kStageGather, 8)
#undef GSN_PRE3
#undef GSN_PRE
  return GSN_OK;
}

}  // namespace gsn

extern "C" int gsn_pre_stream_supported(int K, int H) {
  const int Kmma = (K + 15) / 16 * 16;
  return (K >= 1 && K <= 256 && H >= 1 && 3 * (Kmma / 2) + 2 * 16 <= 512) ? 1 : 0;
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
  GSN_REQUIRE((long long)T * B * N < (1ll << 30), "gsn_pre_stream: T*R too large");
  GSN_REQUIRE(lo + N * ctr <= f_cm, "gsn_pre_stream: band leaves the spectrum");
  GSN_REQUIRE(lo == 0 || lo - nbr >= 0, "gsn_pre_stream: lower neighbourhood out of range");
  GSN_REQUIRE(lo + N * ctr == f_cm || lo + N * ctr + nbr <= f_cm, "gsn_pre_stream: upper neighbourhood out of range");
  GSN_REQUIRE(!fb || (f_fb > 0 && ctr <= f_fb), "gsn_pre_stream: f_fb=%d must be >= ctr=%d", f_fb, ctr);
  GSN_REQUIRE((long long)T * B * (f_cm > f_fb ? f_cm : f_fb) < (1ll << 31), "gsn_pre_stream: inputs too large");
  GSN_REQUIRE((ln_weight == nullptr) == (ln_bias == nullptr), "gsn_pre_stream: ln params");
  StageParams p{};
  p.cm = cm; p.fb = fb; p.ln_w = ln_weight; p.ln_b = ln_bias; p.x_out = x_out;
  p.B = B; p.N = N; p.lo = lo; p.ctr = ctr; p.nbr = nbr; p.f_cm = f_cm; p.f_fb = f_fb; p.eps = ln_eps;
  p.w = w_ih; p.bias = nullptr; p.out = xproj; p.out_act = nullptr; p.act = 0;
  p.in_cnt = in_cnt; p.in_target = in_target; p.out_cnt = out_cnt;
  p.T = T; p.R = B * N; p.K = K; p.Kmma = (K + 15) / 16 * 16; p.Nout = H; p.trace = trace_buffer();
  cudaStream_t st = as_stream(stream);
  // J = 8-feature chunks per lane (4 chunk groups): ceil(Kmma / 32)
  if (p.Kmma <= 64) return dispatch_stage<kStageGather, 2>(p, ctas_per_slice, st);
  if (p.Kmma <= 96) return dispatch_stage<kStageGather, 3>(p, ctas_per_slice, st);
  if (p.Kmma <= 160) return dispatch_stage<kStageGather, 5>(p, ctas_per_slice, st);
  return dispatch_stage<kStageGather, 8>(p, ctas_per_slice, st);
}

extern "C" int gsn_linear_spike_bits_stream(const uint32_t* a_bits, const float* w, const float* bias, float* out,
                                            float* out_act, int act, int T, int R, int K, int N, int ctas,
                                            const unsigned int* in_cnt, unsigned int in_target,
                                            unsigned int* out_cnt, gsn_stream_t stream) {
  using namespace gsn;
  GSN_REQUIRE(a_bits && w && out, "gsn_linear_spike_bits_stream: null pointer");
  GSN_REQUIRE(T > 0 && R > 0 && K > 0 && N > 0, "gsn_linear_spike_bits_stream: bad shape");
  GSN_REQUIRE(act >= 0 && act <= 3, "gsn_linear_spike_bits_stream: unknown activation %d", act);
  GSN_REQUIRE((long long)T * R < (1ll << 30), "gsn_linear_spike_bits_stream: T*R too large");
  StageParams p{};
  p.a_bits = a_bits; p.w = w; p.bias = bias; p.out = out; p.out_act = out_act; p.act = act;
  p.in_cnt = in_cnt; p.in_target = in_target; p.out_cnt = out_cnt;
  p.T = T; p.R = R; p.K = K; p.Kmma = (K + 15) / 16 * 16; p.Nout = N; p.trace = trace_buffer();
  const int slices = (N + 127) / 128;
  const int per_slice = ctas < slices ? 1 : ctas / slices;
  return dispatch_stage<kStageBits, 1>(p, per_slice, as_stream(stream));
}
