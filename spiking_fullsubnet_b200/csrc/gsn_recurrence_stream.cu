// Streaming tcgen05 recurrence: GSULayer.forward ESN:75-81 over GSUCell.forward ESN:132-153 as ONE persistent,
// warp-specialised launch per (sequence model, layer) for all T frames, chained to its producers / consumers through
// per-frame counters in global memory instead of kernel boundaries (StackedGSU.forward ESN:50-62 and the
// full-band -> sub-band hand-over MSF:441-447 become a frame-granular pipeline of co-resident kernels).
//
// Same arithmetic and decomposition as gsn_recurrence_tc.cu (cluster = one tile of NT rows, CTA s = neurons
// [128 s, 128 s + 128), recurrent weights as three exact bf16 planes stationary in tensor memory, spikes exchanged as
// ballot words over DSMEM); what changes is everything around the MMAs:
//   * warp roles: 16 epilogue warps (thread = neuron x 4 row groups), one MMA-issue warp, one loader warp, two
//     publisher warps.  All hand-overs are mbarriers; there is no __syncthreads in the frame loop;
//   * FUSED (layers >= 1): the input-to-hidden product W_ih . h^{l-1}_t runs in the SAME kernel: W_ih sits in tensor
//     memory next to W_hh (H <= 160), the loader warp expands the previous layer's bit-packed spikes of frame t+1 into
//     a second B operand while frame t is in flight, and the issue warp queues those MMAs behind the recurrent ones of
//     frame t, off the critical path.  xproj of layers >= 1 never exists in HBM.  The product lands in its own
//     accumulator and joins as (xproj + bias) + z exactly like the unfused path: bit-identical results;
//   * !FUSED (layer 0 / wide layers): the loader warp stages xproj[t+1..t+RX] tiles in shared memory with bulk
//     asynchronous copies (TMA unit, one per row) counted on an mbarrier: no per-thread global loads in the loop;
//   * the spike trace leaves the kernel bit-packed only (the ballot words); the fp32 [T,R,H] trace the reference
//     returns is optional (strict outputs);
//   * in_cnt / out_cnt: the loader polls in_cnt[t] (acquire) before it touches frame t of its input; the publisher
//     adds 1 to out_cnt[t] (release) once every epilogue warp's stores of frame t are done.
#include <stdlib.h>

#include <type_traits>

#include "gsn_common.cuh"
#include "gsn_tc.cuh"

namespace gsn {

struct RecStreamParams {
  const float* xproj;        // !FUSED: [T, R, H]
  const uint32_t* in_bits;   // IN_BITS: [T, R, Wi] bit-packed spikes of the layer below, Wi = ceil(K_in / 32)
  const uint8_t* in_planes;  // IN_PLANES: operand images of gsn_xplanes_stream, [ring][tiles][3][NT x Kin_mma] bf16
  int planes_ring;           // frames the image buffer holds (frame t in slot t % planes_ring)
  uint8_t* img_out;          // or null: bf16 operand image of this layer's spikes, [img_ring][tiles][NT x Kmma] (16-row tiles)
  int img_ring;              // frames img_out holds
  const unsigned int* bp_cnt;  // or null: frame t may overwrite slot t % img_ring once bp_cnt[t - img_ring] >= bp_target
  unsigned int bp_target;
  const float* w_ih;         // FUSED: [H, K_in]
  const float* w_hh;         // [H, H]
  const float* bias;         // [2H]
  const float* bn_scale;     // [H] or null
  const float* bn_shift;
  uint32_t* h_bits;          // [T, R, Wb]
  float* h_out;              // [T, R, H] or null
  float* c_out;              // [T, R, H] or null
  float* hT;                 // [R, H] or null
  float* cT;
  const unsigned int* in_cnt;   // [T] or null: frame t of the input is complete when in_cnt[t] >= in_target
  unsigned int in_target;
  unsigned int poll_ns;         // back-off between two polls of in_cnt
  int pub_nofence;              // TIMING EXPERIMENT ONLY (GSN_PUB_NOFENCE=1): publish without the release fence
  int dbg;                      // PROF builds only (GSN_TC_DBG): timing experiments, wrong results possible
  int direct;                   // epilogue warps write the bf16 hh operand of the next frame straight into every CTA
  unsigned int* out_cnt;        // [T] or null: += 1 per CTA when its part of frame t is globally visible
  unsigned long long* spike_count;  // or null: += number of spikes emitted by this launch (SynOps accounting)
  unsigned long long* prof;     // PROF builds: cycle counters
  int T, R, H, Kmma;
  int K_in, Kin_mma;
  int wpitch, wpitch_in;
  TraceBuf* trace;
};

constexpr uint32_t kStTmemCols = 512;
constexpr int kStPlanes = 3;
constexpr int kEpiWarps = 16;
constexpr int kIssueWarp = 16, kLoadWarp = 17, kPubWarp = 18;  // publishers: warps 18 .. 18 + kPubWarps - 1
constexpr int kPubWarps = 2;
constexpr int kStThreads = (kPubWarp + kPubWarps) * 32;
constexpr int kRI = 4;        // ring depth of the fused input operand
constexpr int kPubRing = 4;
constexpr uint32_t kOneBf = 0x3F80u;
// Fused real-input product: the recurrence CTA's tensor pipe has room for about 60 extra MMAs per frame next to the 30-45
// recurrent ones (each costs ~25 cycles at 16-row tiles).  From K_in > 112 on (8 k steps: 64 / 80 MMAs with 8 pairs)
// the two smallest plane pairs (lo x mid, mid x lo: together <= 2^-23 |w||x|, the rounding level of an fp32 product) are
// dropped so that the layer keeps the frame rate of the others.
constexpr int kWidePairsKsteps = 8;

template <int NT>
struct StCfg {
  static constexpr int RX = NT <= 16 ? 4 : (NT <= 32 ? 3 : 2);  // ring depth of the staged xproj tiles
};

__host__ __device__ inline int st_kw_padded(int C) { return 4 * C + 1; }

// shared-memory carve-up (bytes); everything is a function of (NT, Kmma, Kin_mma, C, FUSED)
struct StLayout {
  size_t sB, bits, ring, bars, lut, stage, total;
};
enum { kInXproj = 0, kInBits = 1, kInPlanes = 2, kInImage = 3 };
template <int NT, int IN>
__host__ __device__ inline StLayout st_layout(int Kmma, int Kin_mma, int C, int wpitch_max, bool direct) {
  constexpr bool FUSED = IN != kInXproj;
  StLayout L;
  size_t off = 0;
  L.sB = off;
  off += (direct ? 2 : 1) * (((size_t)NT * Kmma * 2 + 127) / 128 * 128);  // direct: one operand buffer per frame parity
  L.bits = off;
  off += ((size_t)2 * NT * st_kw_padded(C) * 4 + 127) / 128 * 128;
  L.ring = off;
  if (FUSED) off += (size_t)kRI * (IN == kInPlanes ? 3 : 1) * (((size_t)NT * Kin_mma * 2 + 127) / 128 * 128);
  else off += (size_t)StCfg<NT>::RX * NT * 128 * 4;
  L.bars = off;
  off += 256;
  L.lut = off;  // 256 x 16 bytes: 8 spike bits -> 8 bf16 {0, 1} (one operand chunk)
  off += 4096;
  L.stage = off;  // weight staging (prologue only)
  off += (size_t)128 * wpitch_max * 4;
  L.total = (off + 127) / 128 * 128;
  return L;
}

__device__ __forceinline__ void st_split3(float w, uint32_t& hi, uint32_t& mid, uint32_t& lo) {
  const uint32_t wb = __float_as_uint(w);
  hi = wb >> 16;
  const float r1 = w - __uint_as_float(wb & 0xFFFF0000u);
  const uint32_t r1b = __float_as_uint(r1);
  mid = r1b >> 16;
  const float r2 = r1 - __uint_as_float(r1b & 0xFFFF0000u);
  lo = __float_as_uint(r2) >> 16;
}

__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_add(unsigned int* p, unsigned int v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_cg_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// Frame counters of a concurrently running producer kernel.  Frames [0, ready) are known complete; one call polls the
// 32 frames from `ready` on (one acquire load per lane, in parallel) and advances `ready` over the complete prefix.
// block: repeat until frame `need` is complete (bounded; false on timeout).  An acquire poll costs an L2 round trip, so
// callers issue the loads of the frame they already own first and poll ahead under them.
__device__ __forceinline__ bool poll_frames(const unsigned int* cnt, unsigned int target, int T, int& ready, int need,
                                            bool block, unsigned int ns, int lane) {
  unsigned long long t0 = 0;
  for (unsigned int spins = 0;; ++spins) {
    const int t = ready + lane;
    const bool ok = t < T && ld_acquire_u32(cnt + t) >= target;
    const unsigned int m = __ballot_sync(0xffffffffu, ok);
    ready += m == 0xffffffffu ? 32 : __ffs(~m) - 1;
    if (ready > need || !block) return true;
    if ((spins & 0x3FFu) == 0x3FFu) {  // wall-clock bound: the producer kernel may start late (lazy module loading)
      const unsigned long long now = tc::wait_clock_ns();
      if (t0 == 0) t0 = now;
      else if (now - t0 > tc::kWaitTimeoutNs) return false;
    }
    __nanosleep(ns);
  }
}

// 128 weight rows (this CTA's neuron slice) -> three exact bf16 planes in tensor memory (lane = neuron).
// Rows are staged in shared memory with one bulk copy per row when aligned, else read directly.
__device__ __forceinline__ void st_weights_to_tmem(const float* w, int ld, int K, int Kmma, int first_row, int nrows,
                                                   float* wst, int wpitch, uint64_t* bar_w, uint32_t bar_parity,
                                                   uint32_t tmem_a, int warp, int lane, bool epi) {
  const int q = warp & 3, g = warp >> 2, tl = q * 32 + lane;
  const bool rv = tl < nrows;
  const float* wrow = w + (size_t)(first_row + (rv ? tl : 0)) * ld;
  const bool staged = wpitch > 0;
  if (staged) {
    if (warp == 0 && lane == 0) tc::mbar_arrive_expect_tx(bar_w, (uint32_t)nrows * (uint32_t)K * 4u);
    __syncthreads();
    if (epi && g == 0 && rv) tc::bulk_g2s(wst + (size_t)tl * wpitch, wrow, (uint32_t)K * 4u, bar_w);
    if (epi && !tc::mbar_wait_cta(bar_w, bar_parity)) __trap();
  }
  if (epi) {
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    const uint32_t plane_cols = Kmma / 2;
    const float* srow = wst + (size_t)tl * wpitch;
    for (int c0 = 8 * g; c0 < (int)plane_cols; c0 += 8 * 4) {
      float wv[16];
      if (staged) {
#pragma unroll
        for (int v4 = 0; v4 < 4; ++v4) {
          const int k = 2 * c0 + 4 * v4;
          const float4 x = (rv && k < K) ? *reinterpret_cast<const float4*>(srow + k) : make_float4(0.f, 0.f, 0.f, 0.f);
          wv[4 * v4 + 0] = x.x; wv[4 * v4 + 1] = x.y; wv[4 * v4 + 2] = x.z; wv[4 * v4 + 3] = x.w;
        }
      } else {
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          const int k = 2 * c0 + e;
          wv[e] = (rv && k < K) ? __ldg(wrow + k) : 0.f;
        }
      }
      uint32_t vh[8], vm[8], vl[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        uint32_t h2[2], m2[2], l2[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) st_split3(wv[2 * u + e], h2[e], m2[e], l2[e]);
        vh[u] = h2[0] | (h2[1] << 16);
        vm[u] = m2[0] | (m2[1] << 16);
        vl[u] = l2[0] | (l2[1] << 16);
      }
      tc::tmem_st8(tmem_a + lane_base + 0 * plane_cols + c0, vl);  // plane 0 = lo (issued first)
      tc::tmem_st8(tmem_a + lane_base + 1 * plane_cols + c0, vm);
      tc::tmem_st8(tmem_a + lane_base + 2 * plane_cols + c0, vh);
    }
    tc::tmem_wait_st();
  }
  __syncthreads();  // the staging area may be reused
}

template <int NT, int IN, bool PROF>
__global__ void __launch_bounds__(kStThreads, 1) k_recurrence_stream(const RecStreamParams p) {
  constexpr bool FUSED = IN != kInXproj;  // the input-to-hidden product runs in this kernel
  constexpr int CPT = NT / 4;                       // rows per epilogue thread
  constexpr int CH = CPT < 8 ? CPT : 8;
  constexpr int RX = StCfg<NT>::RX;
  constexpr int MAXT = (NT * 40 + 511) / 512;       // operand rebuild tasks per epilogue thread (Kmma <= 320)
  static_assert(CPT == 4 || CPT == 8 || CPT == 16, "NT must be 16, 32 or 64");
  extern __shared__ __align__(1024) uint8_t smem[];
  int tslot = trace_begin(p.trace, 2, p.T, p.R, p.H);
  const long long prof_c0 = (PROF || tslot >= 0) ? clock64() : 0;
  const unsigned long long prof_t0 = PROF ? global_ns() : 0;
  // warp index through a shuffle: tells the compiler it is warp-uniform, so the role branches below are uniform
  // control flow and the MMA-issue code may live in uniform registers (CUTLASS's canonical_warp_idx_sync trick)
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
  const bool epi = warp < kEpiWarps;
  const int q = warp & 3, g = (warp >> 2) & 3;
  const uint32_t C = tc::cluster_nctarank(), slice = tc::cluster_ctarank();
  const int row0 = (blockIdx.x / C) * NT;
  const int H = p.H, R = p.R, T = p.T, Kmma = p.Kmma;
  const int tl = q * 32 + lane;
  const int j = slice * 128 + tl;
  const bool comp = epi && j < H;
  const int KWp = st_kw_padded(C);
  const int Wb = (H + 31) / 32;
  const int Wi = FUSED ? (p.K_in + 31) / 32 : 0;
  const int wp_max = p.wpitch > p.wpitch_in ? p.wpitch : p.wpitch_in;
  const bool direct = p.direct != 0;
  const StLayout L = st_layout<NT, IN>(Kmma, p.Kin_mma, (int)C, wp_max, direct);

  uint8_t* sB = smem + L.sB;
  uint32_t* bits = reinterpret_cast<uint32_t*>(smem + L.bits);  // [2][NT][KWp]
  uint8_t* ring = smem + L.ring;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.bars);
  uint64_t* bar_w = bars + 0;
  uint64_t* bar_mma = bars + 1;
  uint64_t* bar_B = bars + 2;
  uint64_t* bar_bits = bars + 3;       // [2]
  uint64_t* bar_in_full = bars + 5;    // [4] FUSED: operand of frame tt ready; else xproj tile landed (tx)
  uint64_t* bar_in_free = bars + 9;    // [4]
  uint64_t* bar_dih_full = bars + 13;
  uint64_t* bar_dih_free = bars + 14;
  uint64_t* bar_pub = bars + 15;       // [kPubRing]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 20);
  volatile int* pub_done = reinterpret_cast<volatile int*>(bars + 21);
  float* wst = reinterpret_cast<float*>(smem + L.stage);
  const uint4* lut = reinterpret_cast<const uint4*>(smem + L.lut);
  if (tid < 256) {  // spike byte -> operand chunk (replaces ~20 shift / mask instructions per chunk in the frame loop)
    uint32_t v[4];
#pragma unroll
    for (int e = 0; e < 4; ++e)
      v[e] = ((tid >> (2 * e)) & 1 ? kOneBf : 0u) | ((tid >> (2 * e + 1)) & 1 ? (kOneBf << 16) : 0u);
    reinterpret_cast<uint4*>(smem + L.lut)[tid] = make_uint4(v[0], v[1], v[2], v[3]);
  }

  if (tid == 0) {
    tc::mbar_init(bar_w, 1);
    tc::mbar_init(bar_mma, 1);
    tc::mbar_init(bar_B, kEpiWarps);
    // direct mode: + one arrival per epilogue warp (its chunks for THIS CTA are plain shared-memory stores)
    tc::mbar_init(&bar_bits[0], p.direct ? 1 + kEpiWarps : 1);
    tc::mbar_init(&bar_bits[1], p.direct ? 1 + kEpiWarps : 1);
    for (int i = 0; i < 4; ++i) {
      tc::mbar_init(&bar_in_full[i], 1);
      tc::mbar_init(&bar_in_free[i], FUSED ? 1 : kEpiWarps);
    }
    tc::mbar_init(bar_dih_full, 1);
    tc::mbar_init(bar_dih_free, kEpiWarps);
    for (int i = 0; i < kPubRing; ++i) tc::mbar_init(&bar_pub[i], kEpiWarps);
    pub_done[0] = 0;
    pub_done[1] = 0;
    tc::fence_mbar_init();
  }
  if (warp == 0) tc::tmem_alloc<kStTmemCols>(tmem_slot);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t lane_base = (uint32_t)(q * 32) << 16;
  const uint32_t tmem_dhh = tmem;
  const uint32_t tmem_dih = tmem + NT;
  const uint32_t tmem_ahh = tmem + (FUSED ? 2 * NT : NT);
  const uint32_t tmem_aih = tmem_ahh + kStPlanes * (Kmma / 2);

  // ---- prologue: weights -> tensor memory; hh operand of frame 0 = zeros -----------------------------------------
  {
    const int first = (int)slice * 128;
    const int nrows = H - first < 128 ? H - first : 128;
    st_weights_to_tmem(p.w_hh, H, H, Kmma, first, nrows, wst, p.wpitch, bar_w, 0, tmem_ahh, warp, lane, epi);
    if (FUSED)
      st_weights_to_tmem(p.w_ih, p.K_in, p.K_in, p.Kin_mma, first, nrows, wst, p.wpitch_in, bar_w, p.wpitch > 0 ? 1 : 0,
                         tmem_aih, warp, lane, epi);
  }
  const uint32_t SBO = 16u * Kmma;
  const int k8n = Kmma / 8;
  uint32_t task_dst[MAXT], task_src[MAXT];
  if (epi) {
#pragma unroll
    for (int it = 0; it < MAXT; ++it) {
      const int i = tid + 512 * it;
      const int nlo = i & 7, k8 = (i >> 3) % k8n, nhi = (i >> 3) / k8n;
      const int n = nhi * 8 + nlo;
      task_dst[it] = i < NT * k8n ? (uint32_t)(nhi * SBO + k8 * 128 + nlo * 16) : 0xFFFFFFFFu;
      task_src[it] = (uint32_t)(n * KWp + (k8 >> 2)) | ((uint32_t)(8 * (k8 & 3)) << 24);
      if (task_dst[it] != 0xFFFFFFFFu) *reinterpret_cast<uint4*>(sB + task_dst[it]) = make_uint4(0, 0, 0, 0);
    }
    tc::fence_proxy_async_smem();
  }

  const int jj = comp ? j : 0;
  const float bf = p.bias[jj], bc = p.bias[H + jj];
  const float bs = p.bn_scale ? p.bn_scale[jj] : 1.0f;
  const float bt = p.bn_shift ? p.bn_shift[jj] : 0.0f;
  const int rfirst = row0 + g * CPT;
  const int nv = comp ? (R - rfirst < CPT ? (R - rfirst > 0 ? R - rfirst : 0) : CPT) : 0;
  const uint32_t hstride = (uint32_t)H * 4u;
  const uint32_t boff0 = ((uint32_t)rfirst * (uint32_t)H + (uint32_t)j) * 4u;
  const size_t frame_bytes = (size_t)R * H * sizeof(float);

  uint32_t snd_cell0 = 0, snd_cell1 = 0, snd_bar0 = 0, snd_bar1 = 0;
  const bool sender = epi && lane < CPT * (int)C;
  if (sender) {
    const int i = lane % CPT;
    const uint32_t r = lane / CPT;
    snd_cell0 = tc::map_to_rank(bits + ((size_t)0 * NT + g * CPT + i) * KWp + slice * 4 + q, r);
    snd_cell1 = tc::map_to_rank(bits + ((size_t)1 * NT + g * CPT + i) * KWp + slice * 4 + q, r);
    snd_bar0 = tc::map_to_rank(&bar_bits[0], r);
    snd_bar1 = tc::map_to_rank(&bar_bits[1], r);
  }
  // direct mode: lane = (destination CTA of a pair, one of the warp's 16 operand chunks = (row i, 8 neurons)); the chunk
  // goes to byte dir_off of the destination's operand buffer of the NEXT frame's parity
  const uint32_t sB_bytes = (uint32_t)(((size_t)NT * Kmma * 2 + 127) / 128 * 128);
  const int dir_i = (lane & 15) >> 2, dir_sub = lane & 3;
  const int dir_k8 = (int)slice * 16 + q * 4 + dir_sub;
  const int dir_n = g * CPT + dir_i;
  const uint32_t dir_off = (uint32_t)((dir_n >> 3) * SBO + dir_k8 * 128 + (dir_n & 7) * 16);
  const bool dir_ok = epi && CPT == 4 && dir_k8 < Kmma / 8;
  // every barrier of the cluster is initialised (and the frame-0 operand written) before anybody stores remotely
  tc::tc_fence_before();
  tc::cluster_sync_all();
  tc::tc_fence_after();

  const uint32_t idesc = tc::make_idesc_f16(128, NT, true);
  const uint32_t bits_bytes = (uint32_t)NT * 4u * C * 4u;
  const int ksteps = Kmma / 16;
  bool alive = true;

  if (warp == kIssueWarp) {
    // =============================== MMA issue warp ===============================
    // The whole warp runs the loop and the barrier waits; only the tcgen05.mma / commit instructions sit in a small
    // elect.sync region, so the compiler keeps their operands in uniform registers (base + immediate).  With the loop
    // itself inside a single-thread region every MMA issue cost ~55 cycles of R2UR traffic.
    const bool leader = tc::elect_one();
    const uint64_t desc_b0 = tc::make_smem_desc(tc::smem_u32(sB), 128, SBO);
    const uint64_t desc_b1 = tc::make_smem_desc(tc::smem_u32(sB + sB_bytes), 128, SBO);
    // bytes of the next frame's operand that arrive as st.async from the OTHER CTAs of the cluster
    const int own_k8 = Kmma / 8 - (int)slice * 16 < 16 ? Kmma / 8 - (int)slice * 16 : 16;
    const uint32_t op_bytes = (uint32_t)NT * 16u * (uint32_t)(Kmma / 8 - (own_k8 > 0 ? own_k8 : 0));
    const uint32_t ih_slot_bytes = (uint32_t)((IN == kInPlanes ? 3 : 1) * (((size_t)NT * p.Kin_mma * 2 + 127) / 128 * 128));
    const int ksteps_in = FUSED ? p.Kin_mma / 16 : 0;
    auto issue_ih = [&](int tt, auto ksi_tag) {  // input-to-hidden product of frame tt -> D_ih
      constexpr int KSI = decltype(ksi_tag)::value;  // k steps of the input product; 0 = runtime count
      const int slot = tt % kRI;
      if (!tc::mbar_wait_cta(&bar_in_full[slot], (uint32_t)((tt / kRI) & 1))) __trap();
      if (tt > 0 && !tc::mbar_wait_cta(bar_dih_free, (uint32_t)((tt - 1) & 1))) __trap();
      tc::tc_fence_after();
      const uint64_t db = tc::make_smem_desc(tc::smem_u32(ring + (size_t)slot * ih_slot_bytes), 128, 16u * p.Kin_mma);
      if (leader) {
        if constexpr (IN == kInPlanes) {  // real-valued input: 8 plane pairs; 6 for wide inputs (see kWidePairsKsteps)
          if constexpr (KSI >= kWidePairsKsteps) tc::mma_pairs_unrolled<KSI, NT * KSI * 2, 6>(tmem_dih, tmem_aih, db, idesc);
          else if constexpr (KSI > 0) tc::mma_pairs_unrolled<KSI, NT * KSI * 2, 8>(tmem_dih, tmem_aih, db, idesc);
          else if (ksteps_in >= kWidePairsKsteps) tc::mma_pairs<NT, 6>(ksteps_in, tmem_dih, tmem_aih, db, idesc);
          else tc::mma_pairs<NT, 8>(ksteps_in, tmem_dih, tmem_aih, db, idesc);
        } else if constexpr (KSI > 0) {
          tc::mma_planes_unrolled<KSI, kStPlanes>(tmem_dih, tmem_aih, db, idesc);
        } else {
          tc::mma_planes<kStPlanes>(ksteps_in, tmem_dih, tmem_aih, db, idesc);
        }
        tc::mma_commit(&bar_in_free[slot]);
        tc::mma_commit(bar_dih_full);
      }
      __syncwarp();
    };
    long long ic[3] = {0, 0, 0};
    // The frame loop is instantiated per number of k steps (KS = 0: runtime count, switch per frame): with the
    // straight-line MMA list selected OUTSIDE the loop its descriptors / tensor-memory addresses are loop invariants
    // (one list per operand buffer), and the per-frame jump table + ~40 descriptor instructions in front of the
    // first MMA (about 270 cycles of every frame) are gone.
    auto issue_loop = [&](auto ks_tag, auto ksi_tag) {
      constexpr int KS = decltype(ks_tag)::value;
      if (FUSED) issue_ih(0, ksi_tag);
      for (int t = 0; t < T; ++t) {
        const long long i0 = PROF ? clock64() : 0;
        if (direct) {
          // operand of frame t (buffer t&1) = st.async stores of every epilogue warp of the cluster, counted in bytes
          if (t > 0 && !tc::mbar_wait_cta(&bar_bits[t & 1], (uint32_t)(((t - 1) >> 1) & 1))) __trap();
          if (!(PROF && (p.dbg & 1))) tc::fence_proxy_async_smem();
        } else if (!tc::mbar_wait_cta(bar_B, (uint32_t)(t & 1))) {
          __trap();
        }
        tc::tc_fence_after();
        const long long i1 = PROF ? clock64() : 0;
        if (leader) {
          if (direct) {
            if constexpr (KS > 0) {
              if (t & 1) tc::mma_planes_unrolled<KS, kStPlanes>(tmem_dhh, tmem_ahh, desc_b1, idesc);
              else tc::mma_planes_unrolled<KS, kStPlanes>(tmem_dhh, tmem_ahh, desc_b0, idesc);
            } else {
              tc::mma_planes<kStPlanes>(ksteps, tmem_dhh, tmem_ahh, (t & 1) ? desc_b1 : desc_b0, idesc);
            }
            tc::mma_commit(bar_mma);
            // arm the next frame's operand barrier (its phase cannot complete before this arrival, so peers' bytes
            // that land earlier are only counted early)
            if (t + 1 < T) tc::mbar_arrive_expect_tx(&bar_bits[(t + 1) & 1], (PROF && (p.dbg & 4)) ? 0u : op_bytes);
          } else {
            tc::mbar_arrive_expect_tx(&bar_bits[t & 1], bits_bytes);  // arm this frame's spike-bit exchange
            if constexpr (KS > 0) tc::mma_planes_unrolled<KS, kStPlanes>(tmem_dhh, tmem_ahh, desc_b0, idesc);
            else tc::mma_planes<kStPlanes>(ksteps, tmem_dhh, tmem_ahh, desc_b0, idesc);
            tc::mma_commit(bar_mma);
          }
        }
        __syncwarp();
        const long long i2 = PROF ? clock64() : 0;
        if (FUSED && t + 1 < T) issue_ih(t + 1, ksi_tag);
        if (PROF) { ic[0] += i1 - i0; ic[1] += i2 - i1; ic[2] += clock64() - i2; }
      }
    };
    // instantiated for the (hidden, input) sizes of the S / M / L recipes and cirm_gsn that take this input mode;
    // everything else runs the runtime lists
#define GSN_LOOP(a, b) issue_loop(std::integral_constant<int, a>{}, std::integral_constant<int, b>{})
    if constexpr (IN == kInXproj) {
      switch (ksteps) {
        case 10: GSN_LOOP(10, 0); break;
        case 14: GSN_LOOP(14, 0); break;
        case 15: GSN_LOOP(15, 0); break;
        case 16: GSN_LOOP(16, 0); break;
        case 17: GSN_LOOP(17, 0); break;
        case 20: GSN_LOOP(20, 0); break;
        default: GSN_LOOP(0, 0); break;
      }
    } else if constexpr (IN == kInBits || IN == kInImage) {
      if (ksteps == 10 && ksteps_in == 10) GSN_LOOP(10, 10);
      else GSN_LOOP(0, 0);
    } else {
      const int combo = ksteps * 100 + ksteps_in;
      switch (combo) {
        case 1003: GSN_LOOP(10, 3); break;
        case 1006: GSN_LOOP(10, 6); break;
        case 1010: GSN_LOOP(10, 10); break;
        case 1504: GSN_LOOP(15, 4); break;
        case 1403: GSN_LOOP(14, 3); break;
        case 1406: GSN_LOOP(14, 6); break;
        case 1603: GSN_LOOP(16, 3); break;
        default: GSN_LOOP(0, 0); break;
      }
    }
#undef GSN_LOOP
    if (PROF && p.prof && blockIdx.x == 0 && leader)
      for (int i = 0; i < 3; ++i) p.prof[8 + i] = (unsigned long long)ic[i];
  } else if (warp == kLoadWarp) {
    // =============================== loader warp ===============================
    int ready = 0;  // frames [0, ready) of the input are known complete
    int ready_bp = 0;  // frames [0, ready_bp) of this layer's operand images have been consumed downstream
    // ring reuse of img_out: the epilogue runs at most a ring of input slots behind this warp, so gating the INPUT of
    // frame tt on the consumer having finished frame tt - img_ring keeps slot tt % img_ring free for the epilogue
    auto bp_wait = [&](int tt) -> bool {
      if (p.bp_cnt == nullptr || tt < p.img_ring || ready_bp > tt - p.img_ring) return true;
      return poll_frames(p.bp_cnt, p.bp_target, T, ready_bp, tt - p.img_ring, true, p.poll_ns, lane);
    };
    if (IN == kInPlanes || IN == kInImage) {
      // one bulk copy per frame: the operand planes of this row tile (three for a real-valued input, one for the spike
      // image the layer below wrote), already in the B-operand layout
      const uint32_t blk_bytes = (IN == kInPlanes ? 3u : 1u) * (uint32_t)NT * (uint32_t)p.Kin_mma * 2u;
      const int ntiles = (R + NT - 1) / NT, tile = blockIdx.x / C;
      for (int tt = 0; tt < T; ++tt) {
        const int slot = tt % kRI;
        if (tt >= kRI && !tc::mbar_wait_cta(&bar_in_free[slot], (uint32_t)(((tt / kRI) - 1) & 1))) { alive = false; break; }
        if (!bp_wait(tt)) { alive = false; break; }
        if (p.in_cnt && ready <= tt) {
          if (!poll_frames(p.in_cnt, p.in_target, T, ready, tt, true, p.poll_ns, lane)) { alive = false; break; }
          asm volatile("fence.proxy.async;" ::: "memory");  // the bulk copy below reads what the producer wrote
        }
        if (lane == 0) {
          tc::mbar_arrive_expect_tx(&bar_in_full[slot], blk_bytes);
          tc::bulk_g2s(ring + (size_t)slot * blk_bytes, p.in_planes + ((size_t)(tt % p.planes_ring) * ntiles + tile) * blk_bytes, blk_bytes,
                       &bar_in_full[slot]);
        }
        __syncwarp();
      }
    } else if (FUSED) {
      const uint32_t slot_bytes = (uint32_t)(((size_t)NT * p.Kin_mma * 2 + 127) / 128 * 128);
      const uint32_t SBOi = 16u * p.Kin_mma;
      const int k8i = p.Kin_mma / 8;
      const int nw = (k8i + 3) / 4;                 // words per row that carry operand columns
      constexpr int MAXW = (NT * 10 + 31) / 32;      // K_in <= 320
      for (int tt = 0; tt < T; ++tt) {
        const int slot = tt % kRI;
        if (tt >= kRI && !tc::mbar_wait_cta(&bar_in_free[slot], (uint32_t)(((tt / kRI) - 1) & 1))) { alive = false; break; }
        if (!bp_wait(tt)) { alive = false; break; }
        if (p.in_cnt && ready <= tt && !poll_frames(p.in_cnt, p.in_target, T, ready, tt, true, p.poll_ns, lane)) {
          alive = false;
          break;
        }
        uint8_t* dst = ring + (size_t)slot * slot_bytes;
        uint32_t wd[MAXW];
#pragma unroll
        for (int it = 0; it < MAXW; ++it) {
          const int i = lane + 32 * it;
          const int nlo = i & 7, wi = (i >> 3) % nw, nhi = (i >> 3) / nw;
          const int row = row0 + nhi * 8 + nlo;
          wd[it] = (i < NT * nw && row < R && wi < Wi && !(p.dbg & 16)) ? ld_cg_u32(p.in_bits + ((size_t)tt * R + row) * Wi + wi) : 0u;
        }
        // poll ahead while the words of frame tt are in flight
        if (p.in_cnt && ready <= tt + 1 && tt + 1 < T) poll_frames(p.in_cnt, p.in_target, T, ready, tt + 1, false, 0, lane);
#pragma unroll
        for (int it = 0; it < MAXW; ++it) {
          const int i = lane + 32 * it;
          if (i >= NT * nw || ((p.dbg & 8) && tt >= kRI)) break;
          const int nlo = i & 7, wi = (i >> 3) % nw, nhi = (i >> 3) / nw;
#pragma unroll
          for (int e4 = 0; e4 < 4; ++e4) {
            const int k8 = 4 * wi + e4;
            if (k8 >= k8i) break;
            const uint32_t b8 = (wd[it] >> (8 * e4)) & 0xFFu;
            *reinterpret_cast<uint4*>(dst + (uint32_t)(nhi * SBOi + k8 * 128 + nlo * 16)) = lut[b8];
          }
        }
        tc::fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&bar_in_full[slot]);
      }
    } else {
      const int first = (int)slice * 128;
      const int ncols = H - first < 128 ? H - first : 128;
      const uint32_t row_bytes = (uint32_t)ncols * 4u;
      const int nrows = R - row0 < NT ? R - row0 : NT;
      const bool bulk_ok = (H % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.xproj) & 15) == 0);
      for (int tt = 0; tt < T; ++tt) {
        const int slot = tt % RX;
        if (tt >= RX && !tc::mbar_wait_cta(&bar_in_free[slot], (uint32_t)(((tt / RX) - 1) & 1))) { alive = false; break; }
        if (!bp_wait(tt)) { alive = false; break; }
        if (p.in_cnt && ready <= tt) {
          if (!poll_frames(p.in_cnt, p.in_target, T, ready, tt, true, p.poll_ns, lane)) { alive = false; break; }
          asm volatile("fence.proxy.async;" ::: "memory");  // the bulk copies below read what the producer wrote
        }
        float* xs = reinterpret_cast<float*>(ring) + (size_t)slot * NT * 128;
        const float* src = p.xproj + ((size_t)tt * R + row0) * H + first;
        if (bulk_ok) {
          if (lane == 0) tc::mbar_arrive_expect_tx(&bar_in_full[slot], (uint32_t)nrows * row_bytes);
          __syncwarp();
          for (int r = lane; r < nrows; r += 32)
            tc::bulk_g2s(xs + (size_t)r * 128, src + (size_t)r * H, row_bytes, &bar_in_full[slot]);
        } else {
          for (int i = lane; i < nrows * ncols; i += 32) {
            const int r = i / ncols, cix = i - r * ncols;
            xs[r * 128 + cix] = __uint_as_float(ld_cg_u32(reinterpret_cast<const uint32_t*>(src + (size_t)r * H + cix)));
          }
          __syncwarp();
          if (lane == 0) tc::mbar_arrive(&bar_in_full[slot]);
        }
      }
    }
    if (!alive) __trap();
  } else if (warp >= kPubWarp) {
    // =============================== publisher warps ===============================
    // One gpu-scope release per frame costs an L2 round trip (about a frame time): the frames alternate between the
    // publisher warps so that two releases are in flight, and neither sits on the epilogue's path.
    if (lane == 0 && p.out_cnt != nullptr) {
      const int pw = warp - kPubWarp;
      for (int t = pw; t < T; t += kPubWarps) {
        if (!tc::mbar_wait_cta(&bar_pub[t % kPubRing], (uint32_t)((t / kPubRing) & 1))) { alive = false; break; }
        // the epilogue warps' stores of frame t were observed through the mbarrier: fence + relaxed add = release
        if (!p.pub_nofence) asm volatile("fence.acq_rel.gpu;" ::: "memory");
        asm volatile("red.relaxed.gpu.global.add.u32 [%0], %1;" ::"l"(p.out_cnt + t), "r"(1u) : "memory");
        pub_done[pw] = t + 1;
      }
      if (!alive) __trap();
    }
    __syncwarp();
  } else {
    // =============================== epilogue warps ===============================
    float c[CPT];
#pragma unroll
    for (int i = 0; i < CPT; ++i) c[i] = 0.f;
    // Rows are independent columns of the MMA, so rows past R (and neurons past H: their lanes never spike, bs_ = 0 and
    // bt_ = -1 below) need no per-element predicate: their ballot words are masked once, here
    const int rows_w = R - rfirst < CPT ? (R - rfirst > 0 ? R - rfirst : 0) : CPT;  // valid rows of this warp
    const uint32_t row_mask = (lane % CPT) < rows_w ? 0xFFFFFFFFu : 0u;
    const float bs_ = comp ? bs : 0.f, bt_ = comp ? bt : -1.f;
    uint32_t lastw[CPT];  // ballot words of the last frame (hT)
    unsigned int nspk = 0;
    // running output pointers of this thread
    const int hb_row = row0 + g * CPT + lane;
    const bool hb_ok = lane < CPT && (int)slice * 4 + q < Wb && hb_row < R;
    uint32_t* hb_ptr = p.h_bits + (size_t)(hb_ok ? hb_row : 0) * Wb + (hb_ok ? slice * 4 + q : 0);
    const size_t hb_step = (size_t)R * Wb;
    const bool do_h = p.h_out != nullptr, do_c = p.c_out != nullptr, do_pub = p.out_cnt != nullptr;
    // operand image of my spikes for the layer above: [img_ring][tiles][NT x Kmma bf16] in this kernel's own operand layout
    const bool do_img = direct && p.img_out != nullptr;
    const bool dir_mine = dir_ok && (slice & 1u) == (uint32_t)(lane >> 4);
    const size_t img_step = (size_t)((R + NT - 1) / NT) * sB_bytes;
    uint8_t* const img_base = p.img_out + (size_t)(blockIdx.x / C) * sB_bytes;
    uint8_t* img_ptr = img_base;
    int img_slot = 0;
    char* h_ptr = reinterpret_cast<char*>(p.h_out) + boff0;
    char* c_ptr = reinterpret_cast<char*>(p.c_out) + boff0;
    // frame-0 operand is in place (zeros)
    __syncwarp();
    if (lane == 0 && !direct) tc::mbar_arrive(bar_B);
    long long pc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int t = 0; t < T; ++t) {
      const int par = t & 1;
      const long long q0 = PROF ? clock64() : 0;
      // ---- input projection of frame t -> registers (under the recurrent MMAs) ----
      float xn[CPT];
      if (FUSED) {
        if (!tc::mbar_wait_cta(bar_dih_full, (uint32_t)par)) { alive = false; break; }
        tc::tc_fence_after();
#pragma unroll
        for (int i0 = 0; i0 < CPT; i0 += CH) {
          uint32_t zr[CH];
          tc::tmem_ld<CH>(tmem_dih + lane_base + g * CPT + i0, zr);
          tc::tmem_wait_ld();
#pragma unroll
          for (int u = 0; u < CH; ++u) xn[i0 + u] = __uint_as_float(zr[u]);
        }
        tc::tc_fence_before();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(bar_dih_free);
      } else {
        const int slot = t % RX;
        if (!tc::mbar_wait_cta(&bar_in_full[slot], (uint32_t)((t / RX) & 1))) { alive = false; break; }
        const float* xs = reinterpret_cast<const float*>(ring) + (size_t)slot * NT * 128 + (size_t)(g * CPT) * 128 + tl;
#pragma unroll
        for (int i = 0; i < CPT; ++i) xn[i] = xs[i * 128];  // (rows past R: stale ring contents, masked below)
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&bar_in_free[slot]);
      }
      float xf_[CPT], xg_[CPT];
#pragma unroll
      for (int i = 0; i < CPT; ++i) {  // reference order: (x W_ih^T + bias) + h W_hh^T   (ESN:140-145)
        xf_[i] = __fadd_rn(xn[i], bf);
        xg_[i] = __fadd_rn(xn[i], bc);
      }
      const long long q1 = PROF ? clock64() : 0;
      if (!tc::mbar_wait_cta(bar_mma, (uint32_t)par)) { alive = false; break; }
      tc::tc_fence_after();
      const long long q2 = PROF ? clock64() : 0;
      // ---- leak / BatchNorm / threshold; spikes -> one bit each ----
      uint32_t myw = 0;
#pragma unroll
      for (int i0 = 0; i0 < CPT; i0 += CH) {
        uint32_t zr[CH];
        tc::tmem_ld<CH>(tmem_dhh + lane_base + g * CPT + i0, zr);
        tc::tmem_wait_ld();
#pragma unroll
        for (int u = 0; u < CH; ++u) {
          // (scalar on purpose: the packed FADD2 / FMUL2 / FFMA2 forms of this block measured 2 % SLOWER per frame)
          const float z = __uint_as_float(zr[u]);
          const float sg = sigmoid_f32(__fadd_rn(xf_[i0 + u], z));
          const float gh = __fadd_rn(xg_[i0 + u], z);
          const float ctil = __fadd_rn(__fmul_rn(sg, c[i0 + u]), __fmul_rn(__fsub_rn(1.0f, sg), gh));
          const float cn = __fadd_rn(__fmul_rn(ctil, bs_), bt_);
          c[i0 + u] = cn;
          const uint32_t w = __ballot_sync(0xffffffffu, cn >= 0.f);
          lastw[i0 + u] = w;
          myw = (lane % CPT) == (i0 + u) ? w : myw;
        }
      }
      myw &= row_mask;
      tc::tc_fence_before();
      const long long q3 = PROF ? clock64() : 0;
      // ---- exchange first (critical path), then the trace (running pointers: no per-frame address arithmetic) ----
      if (direct) {
        const bool more = t + 1 < T;
        if (more || do_img) {
          // my warp's 32 neurons x 4 rows as sixteen 16-byte operand chunks, to every CTA of the cluster
          const uint32_t wrow = __shfl_sync(0xffffffffu, myw, dir_i);
          const uint4 v4 = tc::spike_byte_to_bf16x8(wrow >> (8 * dir_sub));
          const uint32_t v[4] = {v4.x, v4.y, v4.z, v4.w};
          if (more) {
            const uint32_t local = tc::smem_u32(sB + (par ? 0u : sB_bytes)) + dir_off;  // buffer (t+1)&1
            const uint32_t lbar = tc::smem_u32(&bar_bits[par ^ 1]);
            // The chunk for this CTA is a plain store (the st.async path takes ~1.4 cycles per 16-byte packet: 512
            // packets per frame would cost more than the bit exchange it replaces), made visible to the tensor core's
            // proxy and announced with one arrival per warp; the remote chunks follow as st.async, counted in bytes on
            // the peer's barrier (issued after the proxy fence, which would otherwise wait for them)
            if (dir_mine) *reinterpret_cast<uint4*>(sB + (par ? 0u : sB_bytes) + dir_off) = v4;
            if (!(PROF && (p.dbg & 2))) tc::fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&bar_bits[par ^ 1]);
            if (PROF) pc[6] += clock64() - q3;
            if (C <= 2) {  // one peer at most: lanes 16-31 of CTA 0 / lanes 0-15 of CTA 1 send
              const uint32_t r = (uint32_t)(lane >> 4);
              if (dir_ok && r < C && r != slice && !(PROF && (p.dbg & 4)))
                tc::st_async_v4(tc::map_shared_rank(local, r), v, tc::map_shared_rank(lbar, r));
            } else {
              for (uint32_t r0 = 0; r0 < C; r0 += 2) {
                const uint32_t r = r0 + (lane >> 4);
                if (dir_ok && r < C && r != slice && !(PROF && (p.dbg & 4)))
                  tc::st_async_v4(tc::map_shared_rank(local, r), v, tc::map_shared_rank(lbar, r));
              }
            }
            if (PROF) pc[7] += clock64() - q3;
          }
          // the same chunks as this frame's operand image for the layer above (it fetches its input with one bulk
          // copy per frame instead of expanding spike bits): off the critical path, before the frame is published
          if (do_img) {
            if (dir_mine) *reinterpret_cast<uint4*>(img_ptr + dir_off) = v4;
            img_ptr += img_step;
            if (++img_slot == p.img_ring) { img_slot = 0; img_ptr = img_base; }
          }
        }
      } else if (sender) {
        tc::st_async_u32(par ? snd_cell1 : snd_cell0, myw, par ? snd_bar1 : snd_bar0);
      }
      if (lane < CPT) {
        if (hb_ok) *hb_ptr = myw;
        hb_ptr += hb_step;
        nspk += __popc(myw);
      }
      if (do_h) {
#pragma unroll
        for (int i = 0; i < CPT; ++i)
          if (i < nv) *reinterpret_cast<float*>(h_ptr + i * hstride) = (lastw[i] >> lane) & 1u ? 1.0f : 0.0f;
        h_ptr += frame_bytes;
      }
      if (do_c) {
#pragma unroll
        for (int i = 0; i < CPT; ++i)
          if (i < nv) *reinterpret_cast<float*>(c_ptr + i * hstride) = c[i];
        c_ptr += frame_bytes;
      }
      if (do_pub) {
        __syncwarp();
        if (lane == 0) {
          // never run a whole ring ahead of the publisher (mbarrier phases would alias); the phase of frame t needs
          // every warp's arrival, so holding back ONE warp holds the phase
          if (warp == 0 && t >= kPubRing) {
            unsigned int spins = 0;
            while (pub_done[t % kPubWarps] < t - kPubRing + 1) {
              if (++spins > (1u << 24)) __trap();
            }
          }
          tc::mbar_arrive(&bar_pub[t % kPubRing]);
        }
      }
      const long long q4 = PROF ? clock64() : 0;
      if (t + 1 < T && !direct) {
        if (!tc::mbar_wait_cta(&bar_bits[par], (uint32_t)((t >> 1) & 1))) { alive = false; break; }
        const long long q5 = PROF ? clock64() : 0;
        // ---- rebuild the bf16 hh operand (spikes of frame t, all H neurons of my rows) from the bits ----
        const uint32_t* src = bits + (size_t)par * NT * KWp;
#pragma unroll
        for (int it = 0; it < MAXT; ++it) {
          if (task_dst[it] == 0xFFFFFFFFu) continue;
          const uint32_t b8 = (src[task_src[it] & 0xFFFFFFu] >> (task_src[it] >> 24)) & 0xFFu;
          *reinterpret_cast<uint4*>(sB + task_dst[it]) = lut[b8];
        }
        tc::fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(bar_B);
        if (PROF) {
          const long long q6 = clock64();
          pc[4] += q5 - q4; pc[5] += q6 - q5;
        }
      }
      if (PROF) { pc[0] += q1 - q0; pc[1] += q2 - q1; pc[2] += q3 - q2; pc[3] += q4 - q3; }
    }
    if (!__all_sync(0xffffffffu, alive)) __trap();
    if (PROF && p.prof && blockIdx.x == 0 && tid == 0)
      for (int i = 0; i < 8; ++i) p.prof[i] = (unsigned long long)pc[i];
#pragma unroll
    for (int i = 0; i < CPT; ++i) {
      if (i < nv) {
        if (p.cT) *reinterpret_cast<float*>(reinterpret_cast<char*>(p.cT) + boff0 + i * hstride) = c[i];
        if (p.hT) *reinterpret_cast<float*>(reinterpret_cast<char*>(p.hT) + boff0 + i * hstride) = (lastw[i] >> lane) & 1u ? 1.0f : 0.0f;
      }
    }
    if (p.spike_count && lane < CPT && nspk) atomicAdd(p.spike_count, (unsigned long long)nspk);
  }
  tc::tc_fence_before();
  tc::cluster_sync_all();  // nobody leaves while a peer may still store into its staging buffer
  if (warp == 0) tc::tmem_dealloc<kStTmemCols>(tmem);
  if (PROF && p.prof && blockIdx.x == 0 && threadIdx.x == 0) {  // SM cycles and nanoseconds of the whole launch
    p.prof[12] = (unsigned long long)(clock64() - prof_c0);
    p.prof[13] = global_ns() - prof_t0;
  }
  trace_end(p.trace, tslot);
  if (tslot >= 0) p.trace->rec[tslot].c = (int)((clock64() - prof_c0) >> 10);  // SM kilo-cycles of the launch (tools/timeline.py)
}

// ------------------------------------------------------------------------------------------------
static inline int st_wpitch(int K) { return ((K / 4) & 1) ? K : K + 4; }

bool recurrence_stream_fused_fits(int H, int K_in, int nt) {
  const int Kmma = (H + 15) / 16 * 16, Kin = (K_in + 15) / 16 * 16;
  return 2 * nt + kStPlanes * (Kmma / 2) + kStPlanes * (Kin / 2) <= (int)kStTmemCols && K_in <= 320;
}

int recurrence_stream_tile(int R, int H, int K_in, int fused, int sms) {
  const int C = (H + 127) / 128;
  const int Kmma = (H + 15) / 16 * 16;
  int best = 0;
  for (int nt : {16, 32, 64}) {
    if (fused ? !recurrence_stream_fused_fits(H, K_in, nt) : (nt + kStPlanes * (Kmma / 2) > (int)kStTmemCols)) break;
    if ((nt / 4) * C > 32) break;
    best = nt;
    if ((long long)((R + nt - 1) / nt) * C <= sms) break;
  }
  return best;
}

template <int NT, int IN>
static int launch_stream(RecStreamParams p, int C, cudaStream_t st) {
  constexpr bool FUSED = IN != kInXproj;
  p.wpitch = 0;
  p.wpitch_in = 0;
  if (p.H % 4 == 0 && (reinterpret_cast<uintptr_t>(p.w_hh) & 15) == 0) p.wpitch = st_wpitch(p.H);
  if (FUSED && p.K_in % 4 == 0 && (reinterpret_cast<uintptr_t>(p.w_ih) & 15) == 0) p.wpitch_in = st_wpitch(p.K_in);
  auto layout = [&]() {
    return st_layout<NT, IN>(p.Kmma, p.Kin_mma, C, p.wpitch > p.wpitch_in ? p.wpitch : p.wpitch_in, p.direct != 0);
  };
  if (layout().total > tc::kMaxDynamicSmem) {  // no room for the staging area: read the weights directly
    p.wpitch = 0;
    p.wpitch_in = 0;
  }
  size_t smem = layout().total;
  if (smem > tc::kMaxDynamicSmem) return fail(GSN_ENOSUP, "gsn_recurrence_stream: shared memory (%zu B) exceeded", smem);
  if (smem < tc::kTmemExclusiveSmem) smem = tc::kTmemExclusiveSmem;
  static const bool prof = getenv("GSN_TC_PROF") != nullptr;
  auto kern = prof ? k_recurrence_stream<NT, IN, true> : k_recurrence_stream<NT, IN, false>;
  GSN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(((p.R + NT - 1) / NT) * C));
  cfg.blockDim = dim3(kStThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = C;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  GSN_CUDA(cudaLaunchKernelEx(&cfg, kern, p));
  return GSN_OK;
}

}  // namespace gsn

namespace gsn {
// force the (lazily loaded) kernels into the context: see gsn_stream_preload
int preload_recurrence_stream() {
  cudaFuncAttributes a;
#define GSN_PRE(NT, IN) GSN_CUDA(cudaFuncGetAttributes(&a, k_recurrence_stream<NT, IN, false>));
  GSN_PRE(16, kInXproj) GSN_PRE(32, kInXproj) GSN_PRE(64, kInXproj)
  GSN_PRE(16, kInBits) GSN_PRE(32, kInBits) GSN_PRE(64, kInBits)
  GSN_PRE(16, kInPlanes) GSN_PRE(32, kInPlanes) GSN_PRE(64, kInPlanes)
  GSN_PRE(16, kInImage)
#undef GSN_PRE
  return GSN_OK;
}
}  // namespace gsn

extern "C" int gsn_recurrence_stream_tile(int R, int H, int K_in, int fused, int sm_budget) {
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (sm_budget > 0 && sm_budget < sms) sms = sm_budget;
  if (H < 16 || (H + 127) / 128 > 8) return 0;
  return gsn::recurrence_stream_tile(R, H, K_in, fused, sms);
}

extern "C" size_t gsn_spike_image_bytes(int frames, int R, int H) {
  if (frames <= 0 || R <= 0 || H <= 0) return 0;
  return (size_t)frames * (size_t)((R + 15) / 16) * 16u * (size_t)((H + 15) / 16 * 16) * 2u;
}

extern "C" int gsn_recurrence_stream_ctas(int R, int H, int K_in, int fused, int sm_budget) {
  const int nt = gsn_recurrence_stream_tile(R, H, K_in, fused, sm_budget);
  return nt ? ((R + nt - 1) / nt) * ((H + 127) / 128) : 0;
}

extern "C" int gsn_recurrence_stream(const float* xproj, const uint32_t* in_bits, const void* in_planes,
                                     int planes_ring, const float* w_ih, int K_in, const float* w_hh, const float* bias,
                                     const float* bn_scale, const float* bn_shift, uint32_t* h_bits, float* h_out,
                                     float* c_out, float* hT, float* cT, const unsigned int* in_cnt,
                                     unsigned int in_target, unsigned int* out_cnt, unsigned long long* spike_count,
                                     const void* in_image, void* img_out, int img_ring, const unsigned int* bp_cnt,
                                     unsigned int bp_target, int T, int R, int H, int sm_budget, void* workspace,
                                     gsn_stream_t stream) {
  using namespace gsn;
  const int in_mode = in_bits != nullptr ? kInBits : (in_planes != nullptr ? kInPlanes : (in_image != nullptr ? kInImage : kInXproj));
  const bool fused = in_mode != kInXproj;
  GSN_REQUIRE(w_hh && bias && h_bits, "gsn_recurrence_stream: null pointer");
  GSN_REQUIRE((xproj != nullptr) + (in_bits != nullptr) + (in_planes != nullptr) + (in_image != nullptr) == 1,
              "gsn_recurrence_stream: pass exactly one of xproj, (in_bits, w_ih), (in_planes, w_ih), (in_image, w_ih)");
  GSN_REQUIRE(in_mode != kInImage || (reinterpret_cast<uintptr_t>(in_image) & 15) == 0,
              "gsn_recurrence_stream: in_image needs 16-byte alignment");
  GSN_REQUIRE(img_out == nullptr || (reinterpret_cast<uintptr_t>(img_out) & 15) == 0,
              "gsn_recurrence_stream: img_out needs 16-byte alignment");
  GSN_REQUIRE(!fused || (w_ih && K_in > 0), "gsn_recurrence_stream: fused input needs w_ih and K_in");
  GSN_REQUIRE(in_mode != kInPlanes || (K_in <= 256 && (reinterpret_cast<uintptr_t>(in_planes) & 15) == 0),
              "gsn_recurrence_stream: in_planes needs K_in <= 256 and 16-byte alignment");
  GSN_REQUIRE(T > 0 && R > 0 && H >= 16, "gsn_recurrence_stream: bad shape T=%d R=%d H=%d", T, R, H);
  GSN_REQUIRE((bn_scale == nullptr) == (bn_shift == nullptr), "gsn_recurrence_stream: bn params");
  const int nt = gsn_recurrence_stream_tile(R, H, K_in, fused ? 1 : 0, sm_budget);
  if (nt == 0)
    return fail(GSN_ENOSUP, "gsn_recurrence_stream: H=%d K_in=%d fused=%d does not fit tensor memory", H, K_in,
                (int)fused);
  RecStreamParams p{};
  p.xproj = xproj; p.in_bits = in_bits; p.w_ih = w_ih;
  p.in_planes = static_cast<const uint8_t*>(in_mode == kInImage ? in_image : in_planes);
  p.planes_ring = (planes_ring <= 0 || planes_ring > T) ? T : planes_ring;
  GSN_REQUIRE((in_mode != kInPlanes && in_mode != kInImage) || p.planes_ring == T || out_cnt != nullptr,
              "gsn_recurrence_stream: a ring shorter than T needs out_cnt (the producer's back-pressure)");
  GSN_REQUIRE(in_mode != kInImage || nt == 16, "gsn_recurrence_stream: in_image needs the 16-row tile (R=%d H=%d)", R, H);
  GSN_REQUIRE(img_out == nullptr || nt == 16, "gsn_recurrence_stream: img_out needs the 16-row tile (R=%d H=%d)", R, H);
  p.img_out = static_cast<uint8_t*>(img_out);
  p.img_ring = (img_ring <= 0 || img_ring > T) ? T : img_ring;
  GSN_REQUIRE(img_out == nullptr || p.img_ring == T || bp_cnt != nullptr,
              "gsn_recurrence_stream: an image ring shorter than T needs bp_cnt (the consumer's out_cnt)");
  p.bp_cnt = img_out != nullptr && p.img_ring < T ? bp_cnt : nullptr;
  p.bp_target = bp_target;
  p.w_hh = w_hh; p.bias = bias; p.bn_scale = bn_scale;
  p.bn_shift = bn_shift; p.h_bits = h_bits; p.h_out = h_out; p.c_out = c_out; p.hT = hT; p.cT = cT;
  p.in_cnt = in_cnt; p.in_target = in_target; p.out_cnt = out_cnt; p.spike_count = spike_count;
  static const unsigned int poll_ns = getenv("GSN_POLL_NS") ? (unsigned int)atoi(getenv("GSN_POLL_NS")) : 100u;
  p.poll_ns = poll_ns;
  static const int direct = getenv("GSN_STREAM_DIRECT") ? atoi(getenv("GSN_STREAM_DIRECT")) : 1;
  p.direct = (direct != 0 && nt == 16) ? 1 : 0;
  static const int nofence = getenv("GSN_PUB_NOFENCE") ? atoi(getenv("GSN_PUB_NOFENCE")) : 0;
  p.pub_nofence = nofence;
  static const int dbg = getenv("GSN_TC_DBG") ? atoi(getenv("GSN_TC_DBG")) : 0;
  p.dbg = dbg;
  p.prof = reinterpret_cast<unsigned long long*>(workspace);
  p.T = T; p.R = R; p.H = H; p.Kmma = (H + 15) / 16 * 16;
  p.K_in = fused ? K_in : 0; p.Kin_mma = fused ? (K_in + 15) / 16 * 16 : 0;
  p.trace = trace_buffer();
  const int C = (H + 127) / 128;
  cudaStream_t st = as_stream(stream);
#define GSN_ST_DISPATCH(MODE)                                   \
  switch (nt) {                                                 \
    case 16: return launch_stream<16, MODE>(p, C, st);          \
    case 32: return launch_stream<32, MODE>(p, C, st);          \
    default: return launch_stream<64, MODE>(p, C, st);          \
  }
  if (in_mode == kInImage) return launch_stream<16, kInImage>(p, C, st);
  if (in_mode == kInBits) { GSN_ST_DISPATCH(kInBits) }
  if (in_mode == kInPlanes) { GSN_ST_DISPATCH(kInPlanes) }
  GSN_ST_DISPATCH(kInXproj)
#undef GSN_ST_DISPATCH
}
