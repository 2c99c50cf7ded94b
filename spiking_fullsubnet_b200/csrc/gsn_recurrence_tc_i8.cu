// tcgen05 kind::i8 variant of the GSN recurrence (GSN_BACKEND_TCGEN05_I8).  Same decomposition as
// gsn_recurrence_tc.cu (weights stationary in TMEM, cluster split over neurons, DSMEM spike-bit exchange); only
// the arithmetic of the recurrent product differs:
//   * every weight row is scaled by a power of two 2^(8*NPL-1-e_j) (e_j = exponent of the row's largest |w|) and
//     rounded to an integer of 8*NPL bits, stored as NPL byte planes (top plane signed, the others unsigned):
//     NPL = 4 represents every weight down to 2^-8 of the row maximum EXACTLY and the rest to 2^-32 of it;
//   * spikes are uint8 {0,1}; tcgen05.mma.kind::i8 (K = 32 per instruction, half the instructions of bf16 and
//     ~1.7x cheaper each) accumulates each plane in its own int32 TMEM accumulator -- integer sums are exact and
//     order-independent;
//   * the epilogue recombines the planes in fp32 (lowest first) and undoes the row scale: one rounding per
//     recombination step instead of one per accumulated term.
// H <= 512 fits tensor memory with NPL = 3 (H <= 448 with NPL = 4).
#include <stdlib.h>

#include "gsn_common.cuh"
#include "gsn_tc.cuh"

namespace gsn {

struct RecI8Params {
  const float* xproj;   // [T, R, gH]
  const float* w_hh;    // [gH, H]
  const float* bias;    // [2H]
  const float* bn_scale;
  const float* bn_shift;
  const float* h0;
  const float* c0;
  float* h_out;         // [T, R, H]
  float* c_out;         // [T, R, H] or null
  float* hT;
  float* cT;
  int T, R, H, Kmma;    // Kmma = round_up(H, 32)
  unsigned long long* prof;  // [8] cycle counters of CTA 0 / thread 0 (workspace), see tools/tc_profile.py
  TraceBuf* trace;
};

constexpr uint32_t kTmemCols = 512;

// bit words per row in the staging buffer (+1: odd stride, conflict-free reads); wps = words per CTA slice
__host__ __device__ inline int i8_kw_padded(int C, int wps) { return wps * C + 1; }

template <int NT>
__host__ __device__ inline size_t i8_smem_bytes(int Kmma, int C, bool shared) {
  size_t b = (size_t)NT * Kmma;                           // B operand (uint8)
  b = (b + 127) / 128 * 128;
  b += (size_t)2 * NT * i8_kw_padded(C, shared ? 4 : 2) * 4;  // bit staging, double buffered
  b = (b + 15) / 16 * 16;
  b += 64;                                                // barriers + tmem slot
  if (!shared) b += (size_t)NT * 64 * 4;                  // cell-gate accumulators handed across lanes
  return b;
}

// 0 <= x < 2^23 and |x| < 2^22 integer -> float without the conversion pipe
__device__ __forceinline__ float u2f(uint32_t x) { return __uint_as_float(x | 0x4B000000u) - 8388608.0f; }
__device__ __forceinline__ float s2f(uint32_t x) { return __uint_as_float(x + 0x4B400000u) - 12582912.0f; }

template <int NT, int G, int NPL, bool SHARED>
__global__ void __launch_bounds__(128 * G, 1) k_recurrence_i8(const RecI8Params p) {
#ifdef GSN_I8_PROF
  constexpr bool PROF = true;
#else
  constexpr bool PROF = false;
#endif
  constexpr int NTHREADS = 128 * G;
  constexpr int CPT = NT / G;                 // accumulator columns (= rows of the tile) per thread
  constexpr int CH = SHARED ? (CPT < 8 ? CPT : 8) : CPT;  // columns processed together (ILP)
  // SHARED gates: a CTA owns 128 neurons (TMEM lane = neuron).  Unshared gates (w_hh [2H,H]): a CTA owns 64
  // neurons; lanes 0-63 accumulate their forget-gate rows, lanes 64-127 the cell-gate rows of the SAME neurons,
  // which are handed to lanes 0-63 through shared memory once per frame.
  constexpr int NS = SHARED ? 128 : 64;
  constexpr int WPS = NS / 32;
  constexpr int MAXT = (NT * 32 + NTHREADS - 1) / NTHREADS;  // B-operand rebuild tasks per thread (Kmma <= 512)
  static_assert(CPT == 4 || CPT == 8 || CPT == 16, "NT / G must be 4, 8 or 16");
  extern __shared__ __align__(1024) uint8_t smem[];
  const int tslot = trace_begin(p.trace, 2, p.T, p.R, p.H);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q = warp & 3;    // TMEM lane quarter this warp may access
  const int g = warp >> 2;   // column group: rows [g*CPT, g*CPT + CPT) of the tile
  const uint32_t C = tc::cluster_nctarank(), slice = tc::cluster_ctarank();
  const int row0 = (blockIdx.x / C) * NT;
  const int H = p.H, R = p.R, T = p.T, Kmma = p.Kmma;
  const int tl = q * 32 + lane;                    // this thread's TMEM lane
  const bool isg = !SHARED && tl >= 64;            // lane holds a cell-gate row (unshared only)
  const int j = slice * NS + (SHARED ? tl : (tl & 63));  // this thread's neuron
  const bool jv = j < H;
  const bool comp = jv && !isg;                    // this thread integrates the membrane of neuron j
  const int gH = SHARED ? H : 2 * H;
  const int KWp = i8_kw_padded(C, WPS);

  uint8_t* sB = smem;
  size_t off = ((size_t)NT * Kmma + 127) / 128 * 128;
  uint32_t* bits = reinterpret_cast<uint32_t*>(smem + off);  // [2][NT][KWp]
  off += (size_t)2 * NT * KWp * 4;
  off = (off + 15) / 16 * 16;
  uint64_t* bar_mma = reinterpret_cast<uint64_t*>(smem + off);
  uint64_t* bar_bits = bar_mma + 1;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_mma + 3);
  float* zg = reinterpret_cast<float*>(smem + off + 64);  // [NT][64], unshared only

  if (tid == 0) {
    tc::mbar_init(bar_mma, 1);
    tc::mbar_init(&bar_bits[0], 1);  // one local arrive (expect_tx) + the bytes of every slice's bit words
    tc::mbar_init(&bar_bits[1], 1);
    tc::fence_mbar_init();
  }
  if (warp == 0) tc::tmem_alloc<kTmemCols>(tmem_slot);

  // B-operand rebuild tasks of this thread: (row n, 16 consecutive k) -> one 16-byte store (uint8 spikes).
  //   byte(n, k) = (n/8)*SBO + (k/16)*128 + (n%8)*16 + k%16 ;  8 consecutive threads fill one core matrix
  const uint32_t SBO = 8u * Kmma;
  const int k8n = Kmma / 16;
  uint32_t task_dst[MAXT], task_src[MAXT];
#pragma unroll
  for (int it = 0; it < MAXT; ++it) {
    const int i = tid + NTHREADS * it;
    const int nlo = i & 7, k8 = (i >> 3) % k8n, nhi = (i >> 3) / k8n;
    const int n = nhi * 8 + nlo;
    task_dst[it] = i < NT * k8n ? (uint32_t)(nhi * SBO + k8 * 128 + nlo * 16) : 0xFFFFFFFFu;
    task_src[it] = (uint32_t)(n * KWp + (k8 >> 1)) | ((uint32_t)(16 * (k8 & 1)) << 24);
  }
  // frame 0: the initial spikes h0 (zeros when null)
#pragma unroll
  for (int it = 0; it < MAXT; ++it) {
    if (task_dst[it] == 0xFFFFFFFFu) continue;
    const int i = tid + NTHREADS * it;
    const int nlo = i & 7, k8 = (i >> 3) % k8n, nhi = (i >> 3) / k8n;
    const int row = row0 + nhi * 8 + nlo;
    uint32_t v[4] = {0, 0, 0, 0};
    if (p.h0 && row < R) {
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        const int k = k8 * 16 + e;
        if (k < H && p.h0[(size_t)row * H + k] != 0.f) v[e >> 2] |= 1u << (8 * (e & 3));
      }
    }
    *reinterpret_cast<uint4*>(sB + task_dst[it]) = make_uint4(v[0], v[1], v[2], v[3]);
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t lane_base = (uint32_t)(q * 32) << 16;
  const uint32_t tmem_d = tmem;                       // accumulator of plane pl: columns [pl*NT, pl*NT + NT)
  const uint32_t tmem_a = tmem + NPL * NT;            // plane pl: columns [NPL*NT + pl*Kmma/4, ...)
  const uint32_t plane_cols = Kmma / 4;

  // recurrent weights of this thread's TMEM lane -> NPL byte planes of the row-scaled fixed-point value;
  // the G warps that share a lane quarter split the K range between them
  float zscale = 0.f;  // 2^(e_j - (8*NPL-1)): undoes the row scale in the epilogue
  {
    const float* wrow = p.w_hh + (size_t)(jv ? (isg ? H + j : j) : 0) * H;
    float wmax = 0.f;
    if (jv)
      for (int k = 0; k < H; ++k) wmax = fmaxf(wmax, fabsf(__ldg(wrow + k)));
    int e = 0;
    frexpf(wmax, &e);                       // wmax = m * 2^e, m in [0.5, 1)  ->  |w| < 2^e
    if (wmax == 0.f) e = 0;
    const float qscale = ldexpf(1.0f, 8 * NPL - 1 - e);
    zscale = ldexpf(1.0f, e - (8 * NPL - 1));
    for (int c0 = 8 * g; c0 < (int)plane_cols; c0 += 8 * G) {
      uint32_t v[NPL][8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
#pragma unroll
        for (int pl = 0; pl < NPL; ++pl) v[pl][u] = 0;
#pragma unroll
        for (int ee = 0; ee < 4; ++ee) {
          const int k = 4 * (c0 + u) + ee;
          const float w = (jv && k < H) ? __ldg(wrow + k) : 0.f;
          long long q = __float2ll_rn(w * qscale);  // |q| <= 2^(8*NPL-1)
          const long long lim = (1ll << (8 * NPL - 1)) - 1;
          q = q > lim ? lim : q;
#pragma unroll
          for (int pl = 0; pl < NPL; ++pl) v[pl][u] |= (uint32_t)((q >> (8 * pl)) & 0xFF) << (8 * ee);
        }
      }
#pragma unroll
      for (int pl = 0; pl < NPL; ++pl) tc::tmem_st8(tmem_a + lane_base + pl * plane_cols + c0, v[pl]);
    }
    tc::tmem_wait_st();
  }

  const int jj = jv ? j : 0;
  const float bf = p.bias[jj], bc = p.bias[H + jj];
  const float bs = p.bn_scale ? p.bn_scale[jj] : 1.0f;
  const float bt = p.bn_shift ? p.bn_shift[jj] : 0.0f;
  float c[CPT];
  uint32_t boff[CPT];      // byte offset of (row, neuron) inside one frame of a [T, R, H] fp32 tensor
  uint32_t valid = 0;      // bit i: row i of my group exists and my neuron exists
#pragma unroll
  for (int i = 0; i < CPT; ++i) {
    const int row = row0 + g * CPT + i;
    const bool ok = comp && row < R;
    valid |= ok ? 1u << i : 0u;
    boff[i] = ok ? ((uint32_t)row * (uint32_t)H + (uint32_t)j) * 4u : 0u;
    c[i] = (p.c0 && ok) ? *reinterpret_cast<const float*>(reinterpret_cast<const char*>(p.c0) + boff[i]) : 0.f;
  }
  const size_t frame_bytes = (size_t)R * H * sizeof(float);
  const size_t xframe_bytes = (size_t)R * gH * sizeof(float);
  uint32_t xoff[CPT];      // byte offset of my forget-gate input projection inside one [R, gH] frame
#pragma unroll
  for (int i = 0; i < CPT; ++i) {
    const int row = row0 + g * CPT + i;
    xoff[i] = ((valid >> i) & 1u) ? ((uint32_t)row * (uint32_t)gH + (uint32_t)j) * 4u : 0u;
  }

  // spike-bit exchange: lane l < CPT*C of every warp sends word (l % CPT) of its warp to CTA (l / CPT);
  // the remote staging cell and the remote barrier are fixed per frame parity
  uint32_t snd_cell0 = 0, snd_cell1 = 0, snd_bar0 = 0, snd_bar1 = 0;
  const bool sender = q < WPS && lane < CPT * (int)C;
  if (sender) {
    const int i = lane % CPT;
    const uint32_t r = lane / CPT;
    snd_cell0 = tc::map_to_rank(bits + ((size_t)0 * NT + g * CPT + i) * KWp + slice * WPS + q, r);
    snd_cell1 = tc::map_to_rank(bits + ((size_t)1 * NT + g * CPT + i) * KWp + slice * WPS + q, r);
    snd_bar0 = tc::map_to_rank(&bar_bits[0], r);
    snd_bar1 = tc::map_to_rank(&bar_bits[1], r);
  }
  // all barriers of the cluster are initialised before anybody stores remotely
  tc::tc_fence_before();
  tc::cluster_sync_all();
  tc::tc_fence_after();

  const uint32_t idesc_u = tc::make_idesc_i8(128, NT, false, false);  // unsigned digit planes
  const uint32_t idesc_s = tc::make_idesc_i8(128, NT, true, false);   // top plane: signed
  const uint64_t desc_b0 = tc::make_smem_desc(tc::smem_u32(sB), 128, SBO);
  const int ksteps = Kmma / 32;
  const uint32_t bits_bytes = (uint32_t)NT * WPS * C * 4u;  // every slice sends WPS words per row
  bool alive = true;
  float hval[CPT];
#pragma unroll
  for (int i = 0; i < CPT; ++i) hval[i] = 0.f;

  // trace of frame t (spikes, membrane) -> global, coalesced over neurons
  auto store_frame = [&](int t) {
    char* hf = reinterpret_cast<char*>(p.h_out) + (size_t)t * frame_bytes;
    char* cf = p.c_out ? reinterpret_cast<char*>(p.c_out) + (size_t)t * frame_bytes : nullptr;
#pragma unroll
    for (int i = 0; i < CPT; ++i) {
      if ((valid >> i) & 1u) {
        *reinterpret_cast<float*>(hf + boff[i]) = hval[i];
        if (cf) *reinterpret_cast<float*>(cf + boff[i]) = c[i];
      }
    }
  };

  long long pc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int t = 0; t < T; ++t) {
    const int par = t & 1;
    const long long q0 = PROF ? clock64() : 0;
    // ---- recurrent product of frame t: D = W_hh[slice] . h_{t-1}^T --------------------------------
    tc::fence_proxy_async_smem();  // the B operand was written through the generic proxy
    tc::tc_fence_before();
    __syncthreads();
    const long long q1 = PROF ? clock64() : 0;
    if (warp == 0) {
      tc::tc_fence_after();
      if (tc::elect_one()) {
        tc::mbar_arrive_expect_tx(&bar_bits[par], bits_bytes);  // arm this frame's spike-bit exchange
        // straight-line issue (operands = base + immediate), one int32 accumulator per plane
        tc::mma_i8_planes<NPL>(ksteps, tmem_d, NT, tmem_a, desc_b0, idesc_u, idesc_s);
        tc::mma_commit(bar_mma);
      }
      __syncwarp();
    }
    const long long q2 = PROF ? clock64() : 0;
    // ---- work hidden under the MMAs: trace of frame t-1 out, input projection of frame t in ----------
    if (t > 0) store_frame(t - 1);
    float xf_[CPT], xg_[CPT];
    {
      const char* xf = reinterpret_cast<const char*>(p.xproj) + (size_t)t * xframe_bytes;
      float xp[CPT], xq[CPT];
#pragma unroll
      for (int i = 0; i < CPT; ++i) {
        const bool ok = (valid >> i) & 1u;
        xp[i] = ok ? __ldg(reinterpret_cast<const float*>(xf + xoff[i])) : 0.f;
        xq[i] = SHARED ? xp[i] : (ok ? __ldg(reinterpret_cast<const float*>(xf + xoff[i]) + H) : 0.f);
      }
#pragma unroll
      for (int i = 0; i < CPT; ++i) {  // reference order: (x W_ih^T + bias) + h W_hh^T   (ESN:140-145)
        xf_[i] = __fadd_rn(xp[i], bf);
        xg_[i] = __fadd_rn(xq[i], bc);
      }
    }
    if (!tc::mbar_wait(bar_mma, t & 1)) { alive = false; break; }
    tc::tc_fence_after();
    const long long q3 = PROF ? clock64() : 0;

    // ---- leak / BatchNorm / threshold, CH columns at a time; spikes -> one bit each -----------------
    uint32_t myw = 0;
#pragma unroll
    for (int i0 = 0; i0 < CPT; i0 += CH) {
      float zf[CH];
      {
        uint32_t d0[CH];
        tc::tmem_ld<CH>(tmem_d + lane_base + g * CPT + i0, d0);
        tc::tmem_wait_ld();
#pragma unroll
        for (int u = 0; u < CH; ++u) zf[u] = NPL == 1 ? s2f(d0[u]) : u2f(d0[u]);
#pragma unroll
        for (int pl = 1; pl < NPL; ++pl) {
          uint32_t dp[CH];
          tc::tmem_ld<CH>(tmem_d + pl * NT + lane_base + g * CPT + i0, dp);
          tc::tmem_wait_ld();
          const float wgt = (float)(1u << (8 * pl));
#pragma unroll
          for (int u = 0; u < CH; ++u) zf[u] = fmaf(pl == NPL - 1 ? s2f(dp[u]) : u2f(dp[u]), wgt, zf[u]);
        }
#pragma unroll
        for (int u = 0; u < CH; ++u) zf[u] *= zscale;
      }
      float sg[CH], gh[CH];
      if (!SHARED) {  // cell-gate accumulators move from lanes 64-127 to the neuron's lane (0-63)
        if (isg) {
#pragma unroll
          for (int u = 0; u < CH; ++u) zg[(g * CPT + i0 + u) * 64 + (tl & 63)] = zf[u];
        }
        __syncthreads();
      }
#pragma unroll
      for (int u = 0; u < CH; ++u) {
        const float z = zf[u];
        sg[u] = sigmoid_f32(__fadd_rn(xf_[i0 + u], z));
        gh[u] = __fadd_rn(xg_[i0 + u], SHARED ? z : (isg ? 0.f : zg[(g * CPT + i0 + u) * 64 + (tl & 63)]));
      }
#pragma unroll
      for (int u = 0; u < CH; ++u) {
        float cn = __fadd_rn(__fmul_rn(sg[u], c[i0 + u]), __fmul_rn(__fsub_rn(1.0f, sg[u]), gh[u]));
        cn = __fadd_rn(__fmul_rn(cn, bs), bt);
        c[i0 + u] = cn;
        const bool spike = ((valid >> (i0 + u)) & 1u) && cn >= 0.f;
        hval[i0 + u] = spike ? 1.0f : 0.0f;
        const uint32_t w = __ballot_sync(0xffffffffu, spike);
        myw = (lane % CPT) == (i0 + u) ? w : myw;
      }
    }
    const long long q4 = PROF ? clock64() : 0;
    // ---- exchange: ONE asynchronous DSMEM store per sending lane into the staging buffer of a CTA of the
    //      cluster; the bytes are counted on the receiver's mbarrier (no fence, no arrive on this side) ----
    if (sender) tc::st_async_u32(par ? snd_cell1 : snd_cell0, myw, par ? snd_bar1 : snd_bar0);
    const long long q5 = PROF ? clock64() : 0;
    if (!tc::mbar_wait(&bar_bits[par], (t >> 1) & 1)) { alive = false; break; }
    const long long q6 = PROF ? clock64() : 0;
    // ---- rebuild the bf16 B operand (spikes of frame t, all H neurons of my rows) from the bits -------
    {
      const uint32_t* src = bits + (size_t)par * NT * KWp;
#pragma unroll
      for (int it = 0; it < MAXT; ++it) {
        if (task_dst[it] == 0xFFFFFFFFu) continue;
        const uint32_t b16 = (src[task_src[it] & 0xFFFFFFu] >> (task_src[it] >> 24)) & 0xFFFFu;
        uint32_t v[4];
#pragma unroll
        for (int e = 0; e < 4; ++e)  // 4 bits -> 4 bytes of 0/1
          v[e] = (((b16 >> (4 * e)) & 0xFu) * 0x00204081u) & 0x01010101u;
        *reinterpret_cast<uint4*>(sB + task_dst[it]) = make_uint4(v[0], v[1], v[2], v[3]);
      }
    }
    const long long q7 = PROF ? clock64() : 0;
    if (PROF) {
      pc[0] += q1 - q0; pc[1] += q2 - q1; pc[2] += q3 - q2; pc[3] += q4 - q3;
      pc[4] += q5 - q4; pc[5] += q6 - q5; pc[6] += q7 - q6; pc[7] += q7 - q0;
    }
  }
  if (!alive) __trap();  // a broken pipeline fails loudly instead of hanging the device
  if (PROF && p.prof && blockIdx.x == 0 && tid == 0)
    for (int i = 0; i < 8; ++i) p.prof[i] = (unsigned long long)pc[i];

  store_frame(T - 1);
#pragma unroll
  for (int i = 0; i < CPT; ++i) {
    if ((valid >> i) & 1u) {
      if (p.cT) *reinterpret_cast<float*>(reinterpret_cast<char*>(p.cT) + boff[i]) = c[i];
      if (p.hT) *reinterpret_cast<float*>(reinterpret_cast<char*>(p.hT) + boff[i]) = hval[i];
    }
  }
  tc::tc_fence_before();
  tc::cluster_sync_all();  // nobody leaves while a peer may still store into its staging buffer
  if (warp == 0) tc::tmem_dealloc<kTmemCols>(tmem);
  trace_end(p.trace, tslot);
}

// ------------------------------------------------------------------------------------------------
static int i8_planes(int H) { return 4 * ((H + 31) / 32 * 32) / 4 + 4 * 16 <= (int)kTmemCols ? 4 : 3; }

static int i8_pick_nt(int R, int H, int shared, int sm_count) {
  const int C = shared ? (H + 127) / 128 : (H + 63) / 64;
  const int Kmma = (H + 31) / 32 * 32;
  const int npl = i8_planes(H);
  int best = 0;
  for (int nt : {16, 32, 64}) {
    if (npl * Kmma / 4 + npl * nt > (int)kTmemCols) break;
    if ((nt / 4) * C > 32) break;  // one sending lane per (row of the group, destination CTA)
    best = nt;
    if ((long long)((R + nt - 1) / nt) * C <= sm_count) break;  // whole problem co-resident
  }
  return best;
}

int recurrence_i8_tile(int R, int H, int shared, int sms) { return i8_pick_nt(R, H, shared, sms); }

bool recurrence_i8_supported(int R, int H, int shared) {
  if (R <= 0 || H < 16) return false;
  const int C = shared ? (H + 127) / 128 : (H + 63) / 64;
  if (C > 8) return false;
  return i8_pick_nt(R, H, shared, 148) > 0;
}

template <int NT, int NPL, bool SHARED>
static int launch_i8(const RecI8Params& p, int C, cudaStream_t st) {
  constexpr int G = 4;
  size_t smem = i8_smem_bytes<NT>(p.Kmma, C, SHARED);
  if (smem < tc::kTmemExclusiveSmem) smem = tc::kTmemExclusiveSmem;
  GSN_CUDA(cudaFuncSetAttribute(k_recurrence_i8<NT, G, NPL, SHARED>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)smem));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(((p.R + NT - 1) / NT) * C));
  cfg.blockDim = dim3(128 * G);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = C;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  GSN_CUDA(cudaLaunchKernelEx(&cfg, k_recurrence_i8<NT, G, NPL, SHARED>, p));
  return GSN_OK;
}

template <int NPL, bool SHARED>
static int launch_i8_nt(int nt, const RecI8Params& p, int C, cudaStream_t st) {
  switch (nt) {
    case 16: return launch_i8<16, NPL, SHARED>(p, C, st);
    case 32: return launch_i8<32, NPL, SHARED>(p, C, st);
    case 64: return launch_i8<64, NPL, SHARED>(p, C, st);
    default: return fail(GSN_ENOSUP, "gsn_layer_recurrence(TCGEN05_I8): H=%d does not fit tensor memory", p.H);
  }
}

int launch_recurrence_i8(const float* xproj, const float* w_hh, const float* bias, const float* bn_scale,
                         const float* bn_shift, const float* h0, const float* c0, float* h_out, float* c_out,
                         float* hT, float* cT, int T, int R, int H, int shared, int sm_budget, void* workspace,
                         cudaStream_t st) {
  int dev = 0, sms = 148;
  GSN_CUDA(cudaGetDevice(&dev));
  GSN_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  RecI8Params p{xproj, w_hh, bias, bn_scale, bn_shift, h0, c0, h_out, c_out, hT, cT, T, R, H, (H + 31) / 32 * 32,
                reinterpret_cast<unsigned long long*>(workspace), trace_buffer()};
  const int C = shared ? (H + 127) / 128 : (H + 63) / 64;
  if (sm_budget > 0 && sm_budget < sms) sms = sm_budget;
  const int nt = i8_pick_nt(R, H, shared, sms);
  const int npl = i8_planes(H);
  if (shared) return npl == 4 ? launch_i8_nt<4, true>(nt, p, C, st) : launch_i8_nt<3, true>(nt, p, C, st);
  return npl == 4 ? launch_i8_nt<4, false>(nt, p, C, st) : launch_i8_nt<3, false>(nt, p, C, st);
}

}  // namespace gsn
