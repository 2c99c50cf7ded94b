// tcgen05 / TMEM recurrence (GSN_BACKEND_TCGEN05): GSULayer.forward ESN:75-81 over GSUCell.forward ESN:132-153.
//
// Work decomposition ("weights stationary, swap-AB"):
//   * a thread-block CLUSTER owns one tile of NT rows (independent recurrences) for all T frames;
//   * CTA `s` of the cluster owns the 128-neuron slice [128 s, 128 s + 128) of the H hidden units:
//       D_s[128 neurons x NT rows] = W_hh[slice, :] (A operand)  x  h_{t-1}[rows, :]^T (B operand)
//   * A = the recurrent weights, split into three bf16 planes  w = hi + mid + lo  (EXACT: fp32 has 24
//     significand bits = 3 x 8), written ONCE into tensor memory (TMEM) and never moved again;
//   * B = the spikes of the previous frame, exactly {0,1} in bf16, K-major in shared memory -> every
//     product is exact and the three planes accumulate into one fp32 TMEM accumulator, lo plane first so
//     the accumulator's running sum only loses bits that fp32 could not hold anyway;
//   * epilogue: thread j of the CTA = neuron j of the slice (= TMEM lane j): it reads its NT accumulators
//     with tcgen05.ld, applies leak / BatchNorm / threshold with the membrane potential c[NT] held in
//     REGISTERS for the whole sequence, writes the spike trace (coalesced over neurons), and the warp
//     ballots the new spikes into 1 bit each;
//   * exchange: the bit words go to the staging buffer of EVERY CTA of the cluster through distributed
//     shared memory with st.async (bytes counted on the receiver's mbarrier: no fence, no arrive); each
//     CTA then expands the bits of its NT rows x H neurons back into the bf16 B operand for frame t+1.
// 4 x G warps: warp w works on TMEM lane quarter w%4 (its 32 neurons) and on rows [ (w/4)*NT/G, ... ) of
// the tile; one elected lane of warp 0 issues the MMAs (single-thread tcgen05.mma).
#include <stdlib.h>

#include "gsn_common.cuh"
#include "gsn_tc.cuh"

namespace gsn {

struct RecTcParams {
  const float* xproj;   // [T, R, gH]
  const float* w_hh;    // [gH, H]
  const float* bias;    // [2H]
  const float* bn_scale;
  const float* bn_shift;
  const float* h0;
  const float* c0;
  float* h_out;         // [T, R, H]
  float* c_out;         // [T, R, H] or null
  uint32_t* h_bits;     // [T, R, ceil(H/32)] bit-packed spike trace or null (neuron n = bit n%32 of word n/32)
  float* hT;
  float* cT;
  int T, R, H, Kmma;    // Kmma = round_up(H, 16)
  int dbg;              // PROF builds only (GSN_TC_DBG, timing experiments): 1 = no trace stores, 2 = no xproj loads
  int wpitch;           // > 0: stage the CTA's weight rows in shared memory (row pitch in floats), else direct loads
  unsigned long long* prof;  // [8] cycle counters of CTA 0 / thread 0 (workspace), see tools/tc_profile.py
  TraceBuf* trace;
  // training path only (TRAIN kernels; bn_scale / bn_shift are unused there)
  const float* bn_w;         // [H] BatchNorm affine or null (no BatchNorm)
  const float* bn_b;
  float* run_mean;           // [H] running statistics: read (eval) / updated every frame (batch statistics)
  float* run_var;
  float* f_out;              // [T,R,H] saved for BPTT: sigmoid(forget gate)
  float* g_out;              // [T,R,H] cell-gate pre-activation
  float* xhat_out;           // [T,R,H] normalised pre-BN membrane (batch statistics only)
  float* invstd_out;         // [T,H]
  float* partial;            // [2][tiles][2][C*128] per-tile sums of the BatchNorm statistics
  unsigned int* counter;     // grid barrier
  float momentum, eps;
  int batch_stats;
};

constexpr int kTcPlanes = 3;
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kOneBf16 = 0x3F80u;

// bit words per row in the staging buffer (+1: odd stride, conflict-free reads); wps = words per CTA slice
__host__ __device__ inline int tc_kw_padded(int C, int wps) { return wps * C + 1; }

template <int NT>
__host__ __device__ inline size_t tc_smem_bytes(int Kmma, int C, bool shared) {
  size_t b = (size_t)NT * Kmma * 2;                       // B operand
  b = (b + 127) / 128 * 128;
  b += (size_t)2 * NT * tc_kw_padded(C, shared ? 4 : 2) * 4;  // bit staging, double buffered
  b = (b + 15) / 16 * 16;
  b += 64;                                                // barriers + tmem slot
  if (!shared) b += (size_t)NT * 64 * 4;                  // cell-gate accumulators handed across lanes
  b += 4 * 128 * 2 * 4;                                   // training: cross-group reduction scratch
  return (b + 127) / 128 * 128;
}

// Row pitch (floats) of the weight staging area: rows are 16-byte aligned (bulk copies) and pitch/4 is odd, so the
// 8 lanes of one shared-memory phase read their 16 bytes from 8 different bank groups.
__host__ __device__ inline int tc_wpitch(int H) { return ((H / 4) & 1) ? H : H + 4; }

// exact 3-way split of an fp32 value into bf16 planes by truncation: w == hi + mid + lo
__device__ __forceinline__ void split3(float w, uint32_t& hi, uint32_t& mid, uint32_t& lo) {
  const uint32_t wb = __float_as_uint(w);
  hi = wb >> 16;
  const float r1 = w - __uint_as_float(wb & 0xFFFF0000u);   // exact
  const uint32_t r1b = __float_as_uint(r1);
  mid = r1b >> 16;
  const float r2 = r1 - __uint_as_float(r1b & 0xFFFF0000u);  // exact, <= 8 significant bits
  lo = __float_as_uint(r2) >> 16;
}

template <int NT, int G, bool PROF, bool SHARED, bool TRAIN>
__global__ void __launch_bounds__(128 * G, 1) k_recurrence_tc(const RecTcParams p) {
  static_assert(!TRAIN || SHARED, "the tcgen05 training forward handles shared gate weights only");
  constexpr int NTHREADS = 128 * G;
  constexpr int CPT = NT / G;                 // accumulator columns (= rows of the tile) per thread
  constexpr int CH = (SHARED && !TRAIN) ? (CPT < 8 ? CPT : 8) : CPT;  // columns processed together (ILP)
  // SHARED gates: a CTA owns 128 neurons (TMEM lane = neuron).  Unshared gates (w_hh [2H,H]): a CTA owns 64
  // neurons; lanes 0-63 accumulate their forget-gate rows, lanes 64-127 the cell-gate rows of the SAME neurons,
  // which are handed to lanes 0-63 through shared memory once per frame.
  constexpr int NS = SHARED ? 128 : 64;
  constexpr int WPS = NS / 32;
  constexpr int MAXT = (NT * 40 + NTHREADS - 1) / NTHREADS;  // B-operand rebuild tasks per thread (Kmma <= 320)
  static_assert(CPT == 4 || CPT == 8 || CPT == 16, "NT / G must be 4, 8 or 16");
  constexpr bool PF = CPT <= 8 || (SHARED && !TRAIN);  // prefetch xproj one frame ahead where registers allow
  extern __shared__ __align__(1024) uint8_t smem[];
  int tslot = trace_begin(p.trace, 2, p.T, p.R, p.H);
  if (blockIdx.x != 0) tslot = trace_begin_cta(p.trace, 5, (int)blockIdx.x, p.R, p.H);  // per-CTA records (dev aid)
  const long long e0 = PROF ? clock64() : 0;
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
  const int KWp = tc_kw_padded(C, WPS);
  const int Wb = (H + 31) / 32;

  uint8_t* sB = smem;
  size_t off = ((size_t)NT * Kmma * 2 + 127) / 128 * 128;
  uint32_t* bits = reinterpret_cast<uint32_t*>(smem + off);  // [2][NT][KWp]
  off += (size_t)2 * NT * KWp * 4;
  off = (off + 15) / 16 * 16;
  uint64_t* bar_mma = reinterpret_cast<uint64_t*>(smem + off);
  uint64_t* bar_bits = bar_mma + 1;  // [2]
  uint64_t* bar_w = bar_mma + 3;     // weight staging (bulk copies)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_mma + 4);
  float* zg = reinterpret_cast<float*>(smem + off + 64);  // [NT][64], unshared only
  float* red = reinterpret_cast<float*>(smem + off + 64);  // [G][128][2], TRAIN only (SHARED: zg unused)
  const float* wst = reinterpret_cast<const float*>(smem + tc_smem_bytes<NT>(Kmma, C, SHARED));  // [128][wpitch]

  // this thread's recurrent weight row (TMEM lane tl): staged with ONE 1-D bulk copy per row (coalesced, TMA unit)
  const float* wrow = p.w_hh + (size_t)(jv ? (isg ? H + j : j) : 0) * H;
  if (tid == 0) {
    tc::mbar_init(bar_mma, 1);
    tc::mbar_init(&bar_bits[0], 1);  // one local arrive (expect_tx) + the bytes of every slice's bit words
    tc::mbar_init(&bar_bits[1], 1);
    tc::mbar_init(bar_w, 1);
    tc::fence_mbar_init();
    if (p.wpitch > 0) {
      const int first = (int)slice * NS;
      const int nrows = (H - first < NS ? H - first : NS) * (SHARED ? 1 : 2);
      tc::mbar_arrive_expect_tx(bar_w, (uint32_t)nrows * (uint32_t)H * 4u);
    }
  }
  if (warp == 0) tc::tmem_alloc<kTmemCols>(tmem_slot);
  if (p.wpitch > 0) {
    __syncthreads();  // the barrier is initialised and armed before any copy can complete on it
    if (g == 0 && jv)
      tc::bulk_g2s(const_cast<float*>(wst) + (size_t)tl * p.wpitch, wrow, (uint32_t)H * 4u, bar_w);
  }

  // B-operand rebuild tasks of this thread: (row n, 8 consecutive k) -> one 16-byte store.
  //   byte(n, k) = (n/8)*SBO + (k/8)*128 + (n%8)*16 ;  8 consecutive threads fill one 128-byte core matrix
  const uint32_t SBO = 16u * Kmma;
  const int k8n = Kmma / 8;
  uint32_t task_dst[MAXT], task_src[MAXT];
#pragma unroll
  for (int it = 0; it < MAXT; ++it) {
    const int i = tid + NTHREADS * it;
    const int nlo = i & 7, k8 = (i >> 3) % k8n, nhi = (i >> 3) / k8n;
    const int n = nhi * 8 + nlo;
    task_dst[it] = i < NT * k8n ? (uint32_t)(nhi * SBO + k8 * 128 + nlo * 16) : 0xFFFFFFFFu;
    task_src[it] = (uint32_t)(n * KWp + (k8 >> 2)) | ((uint32_t)(8 * (k8 & 3)) << 24);
  }
  // Programmatic dependent launch (GSN_OPT_PDL): this grid may have been scheduled before the previous kernel of its
  // stream -- the previous frame chunk of the same layer -- has finished; everything it produced (h0, c0) and every
  // other input is only read after this point.  A no-op for ordinary launches.
  asm volatile("griddepcontrol.wait;" ::: "memory");
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
      for (int e = 0; e < 8; ++e) {
        const int k = k8 * 8 + e;
        if (k < H && p.h0[(size_t)row * H + k] != 0.f) v[e >> 1] |= kOneBf16 << (16 * (e & 1));
      }
    }
    *reinterpret_cast<uint4*>(sB + task_dst[it]) = make_uint4(v[0], v[1], v[2], v[3]);
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const long long e1 = PROF ? clock64() : 0;
  const uint32_t tmem = *tmem_slot;
  const uint32_t lane_base = (uint32_t)(q * 32) << 16;
  const uint32_t tmem_d = tmem;                       // accumulator: columns [0, NT)
  const uint32_t tmem_a = tmem + NT;                  // plane pl: columns [NT + pl*Kmma/2, ...)
  const uint32_t plane_cols = Kmma / 2;

  // recurrent weights of this thread's neuron -> three exact bf16 planes in TMEM (lane = neuron);
  // the G warps that share a lane quarter split the K range between them
  {
    const bool staged = p.wpitch > 0;
    if (staged && !tc::mbar_wait_cta(bar_w, 0)) __trap();
    const float* srow = wst + (size_t)tl * p.wpitch;
    for (int c0 = 8 * g; c0 < (int)plane_cols; c0 += 8 * G) {
      float wv[16];
      if (staged) {
#pragma unroll
        for (int v4 = 0; v4 < 4; ++v4) {
          const int k = 2 * c0 + 4 * v4;
          const float4 x = (jv && k < H) ? *reinterpret_cast<const float4*>(srow + k) : make_float4(0.f, 0.f, 0.f, 0.f);
          wv[4 * v4 + 0] = x.x; wv[4 * v4 + 1] = x.y; wv[4 * v4 + 2] = x.z; wv[4 * v4 + 3] = x.w;
        }
      } else {
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          const int k = 2 * c0 + e;
          wv[e] = (jv && k < H) ? __ldg(wrow + k) : 0.f;
        }
      }
      uint32_t vh[8], vm[8], vl[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        uint32_t h2[2], m2[2], l2[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) split3(wv[2 * u + e], h2[e], m2[e], l2[e]);
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
  const long long e2 = PROF ? clock64() : 0;

  const int jj = jv ? j : 0;
  const float bf = p.bias[jj], bc = p.bias[H + jj];
  float bs = p.bn_scale ? p.bn_scale[jj] : 1.0f;   // BatchNorm as y = x * bs + bt
  float bt = p.bn_shift ? p.bn_shift[jj] : 0.0f;
  // training path: affine / statistics handled here instead of the pre-folded bn_scale / bn_shift
  const bool has_bn = TRAIN && p.bn_w != nullptr;
  const bool batch_stats = has_bn && p.batch_stats;
  const float gam = has_bn ? p.bn_w[jj] : 1.f, bet = has_bn ? p.bn_b[jj] : 0.f;
  float rmean = has_bn ? p.run_mean[jj] : 0.f, rvar = has_bn ? p.run_var[jj] : 1.f;
  if (TRAIN) {
    bs = 1.f;
    bt = 0.f;
    if (has_bn && !batch_stats) {  // eval-mode fold, as torch's CPU kernel does it
      bs = gam * (1.0f / sqrtf(rvar + p.eps));
      bt = bet - rmean * bs;
    }
  }
  const unsigned int ntiles = gridDim.x / C, tile = blockIdx.x / C;
  const int Hp = (int)C * 128;
  unsigned int epoch = 0;
  float shift = 0.f;  // one-pass variance is taken around the previous frame's mean
  float c[CPT];
  // My CPT rows are consecutive, so the valid ones are a PREFIX [0, nv) (nv = 0 when my neuron does not exist) and
  // the byte offset of (row i, my neuron) inside one frame is boff0 + i * hstride: no per-column offset registers.
  const int rfirst = row0 + g * CPT;
  const int nv = comp ? (R - rfirst < CPT ? (R - rfirst > 0 ? R - rfirst : 0) : CPT) : 0;
  const bool full = nv == CPT;
  const uint32_t hstride = (uint32_t)H * 4u, xstride = (uint32_t)gH * 4u;
  const uint32_t boff0 = ((uint32_t)rfirst * (uint32_t)H + (uint32_t)j) * 4u;   // [T, R, H] fp32 tensors
  const uint32_t xoff0 = ((uint32_t)rfirst * (uint32_t)gH + (uint32_t)j) * 4u;  // xproj [T, R, gH]
#pragma unroll
  for (int i = 0; i < CPT; ++i)
    c[i] = (p.c0 && i < nv) ? *reinterpret_cast<const float*>(reinterpret_cast<const char*>(p.c0) + boff0 + i * hstride)
                            : 0.f;
  const size_t frame_bytes = (size_t)R * H * sizeof(float);
  const size_t xframe_bytes = (size_t)R * gH * sizeof(float);

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

  const uint32_t idesc = tc::make_idesc_f16(128, NT, true);
  const uint64_t desc_b0 = tc::make_smem_desc(tc::smem_u32(sB), 128, SBO);
  const int ksteps = Kmma / 16;
  const uint32_t bits_bytes = (uint32_t)NT * WPS * C * 4u;  // every slice sends WPS words per row
  bool alive = true;
  float hval[CPT], fsave[TRAIN ? CPT : 1], gsave[TRAIN ? CPT : 1], xsave[TRAIN ? CPT : 1];
#pragma unroll
  for (int i = 0; i < CPT; ++i) hval[i] = 0.f;

  // trace of frame t (spikes, membrane) -> global, coalesced over neurons
  auto store_frame = [&](int t) {
    char* hf = reinterpret_cast<char*>(p.h_out) + (size_t)t * frame_bytes + boff0;
    char* cf = p.c_out ? reinterpret_cast<char*>(p.c_out) + (size_t)t * frame_bytes + boff0 : nullptr;
    if (full && !TRAIN) {  // the common case: a straight run of stores
#pragma unroll
      for (int i = 0; i < CPT; ++i) *reinterpret_cast<float*>(hf + i * hstride) = hval[i];
      if (cf) {
#pragma unroll
        for (int i = 0; i < CPT; ++i) *reinterpret_cast<float*>(cf + i * hstride) = c[i];
      }
      return;
    }
#pragma unroll
    for (int i = 0; i < CPT; ++i) {
      if (i < nv) {
        *reinterpret_cast<float*>(hf + i * hstride) = hval[i];
        if (cf) *reinterpret_cast<float*>(cf + i * hstride) = c[i];
        if (TRAIN && p.f_out) {
          const size_t o = (size_t)t * frame_bytes + boff0 + i * hstride;
          *reinterpret_cast<float*>(reinterpret_cast<char*>(p.f_out) + o) = fsave[i];
          *reinterpret_cast<float*>(reinterpret_cast<char*>(p.g_out) + o) = gsave[i];
          if (batch_stats) *reinterpret_cast<float*>(reinterpret_cast<char*>(p.xhat_out) + o) = xsave[i];
        }
      }
    }
  };

  // input projection of one frame -> registers, always fetched ONE FRAME AHEAD of its use
  float xn[CPT], xqn[SHARED ? 1 : CPT];
  auto load_xproj = [&](int t) {
    const char* xf = reinterpret_cast<const char*>(p.xproj) + (size_t)t * xframe_bytes + xoff0;
#pragma unroll
    for (int i = 0; i < CPT; ++i) {
      const bool ok = full || i < nv;
      xn[i] = ok ? __ldg(reinterpret_cast<const float*>(xf + i * xstride)) : 0.f;
      if (!SHARED) xqn[i] = ok ? __ldg(reinterpret_cast<const float*>(xf + i * xstride) + H) : 0.f;
    }
  };
  if (PF) load_xproj(0);

  long long pc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const long long e3 = PROF ? clock64() : 0;
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
        // plane 0 = lo first; straight-line issue (see mma_planes_unrolled): ~8 instead of ~45 cycles per MMA
        tc::mma_planes<kTcPlanes>(ksteps, tmem_d, tmem_a, desc_b0, idesc);
        tc::mma_commit(bar_mma);
      }
      __syncwarp();
    }
    const long long q2 = PROF ? clock64() : 0;
    // ---- work hidden under the MMAs: trace of frame t-1 out, input projection of frame t in ----------
    if (t > 0 && !(PROF && (p.dbg & 1))) store_frame(t - 1);
    float xf_[CPT], xg_[CPT];
    if (!PF) load_xproj(t);
#pragma unroll
    for (int i = 0; i < CPT; ++i) {  // reference order: (x W_ih^T + bias) + h W_hh^T   (ESN:140-145)
      xf_[i] = __fadd_rn(xn[i], bf);
      xg_[i] = __fadd_rn(SHARED ? xn[i] : xqn[i], bc);
    }
    if (PF && t + 1 < T && !(PROF && (p.dbg & 2))) load_xproj(t + 1);  // consumed one frame later: a full frame of latency tolerance
    if (!tc::mbar_wait_cta(bar_mma, t & 1)) { alive = false; break; }
    tc::tc_fence_after();
    const long long q3 = PROF ? clock64() : 0;

    // ---- leak / BatchNorm / threshold, CH columns at a time; spikes -> one bit each -----------------
    uint32_t myw = 0;
#pragma unroll
    for (int i0 = 0; i0 < CPT; i0 += CH) {
      uint32_t zr[CH];
      tc::tmem_ld<CH>(tmem_d + lane_base + g * CPT + i0, zr);
      tc::tmem_wait_ld();
      float sg[CH], gh[CH];
      if (!SHARED) {  // cell-gate accumulators move from lanes 64-127 to the neuron's lane (0-63)
        if (isg) {
#pragma unroll
          for (int u = 0; u < CH; ++u) zg[(g * CPT + i0 + u) * 64 + (tl & 63)] = __uint_as_float(zr[u]);
        }
        __syncthreads();
      }
#pragma unroll
      for (int u = 0; u < CH; ++u) {
        const float z = __uint_as_float(zr[u]);
        sg[u] = sigmoid_f32(__fadd_rn(xf_[i0 + u], z));
        gh[u] = __fadd_rn(xg_[i0 + u], SHARED ? z : (isg ? 0.f : zg[(g * CPT + i0 + u) * 64 + (tl & 63)]));
      }
      float ctil[CH];
#pragma unroll
      for (int u = 0; u < CH; ++u)
        ctil[u] = __fadd_rn(__fmul_rn(sg[u], c[i0 + u]), __fmul_rn(__fsub_rn(1.0f, sg[u]), gh[u]));
      float mean = 0.f, invstd = 1.f;
      if (TRAIN && batch_stats) {
        // per-frame BatchNorm statistics over ALL rows of the sequence model: my rows -> the 4 row groups of the
        // CTA (shared memory) -> all row tiles (global partials + grid barrier), summed in a fixed order
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int u = 0; u < CH; ++u) {
          if (i0 + u < nv) {
            const float d = ctil[u] - shift;
            s1 += d;
            s2 += d * d;
          }
        }
        red[(g * 128 + tl) * 2 + 0] = s1;
        red[(g * 128 + tl) * 2 + 1] = s2;
        __syncthreads();
        float* part = p.partial + (size_t)par * ntiles * 2 * Hp;
        if (g == 0) {
          float t1 = 0.f, t2 = 0.f;
#pragma unroll
          for (int gg = 0; gg < G; ++gg) {
            t1 += red[(gg * 128 + tl) * 2 + 0];
            t2 += red[(gg * 128 + tl) * 2 + 1];
          }
          part[((size_t)tile * 2 + 0) * Hp + slice * 128 + tl] = t1;
          part[((size_t)tile * 2 + 1) * Hp + slice * 128 + tl] = t2;
        }
        if (!grid_barrier(p.counter, gridDim.x, epoch)) __trap();
        float a1 = 0.f, a2 = 0.f;
        for (unsigned int b0 = 0; b0 < ntiles; b0 += 8) {
          float v1[8], v2[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const bool ok = b0 + u < ntiles;
            v1[u] = ok ? __ldcg(part + ((size_t)(b0 + u) * 2 + 0) * Hp + slice * 128 + tl) : 0.f;
            v2[u] = ok ? __ldcg(part + ((size_t)(b0 + u) * 2 + 1) * Hp + slice * 128 + tl) : 0.f;
          }
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            a1 += v1[u];
            a2 += v2[u];
          }
        }
        const float m1 = a1 / (float)R;
        const float var = fmaxf(a2 / (float)R - m1 * m1, 0.f);  // biased variance
        mean = shift + m1;
        invstd = 1.0f / sqrtf(var + p.eps);
        bs = gam * invstd;
        bt = bet - mean * bs;
        rmean = (1.f - p.momentum) * rmean + p.momentum * mean;
        rvar = (1.f - p.momentum) * rvar + p.momentum * (var * (float)R / (float)(R - 1));
        shift = mean;
        if (tile == 0 && g == 0 && jv) p.invstd_out[(size_t)t * H + j] = invstd;
      }
#pragma unroll
      for (int u = 0; u < CH; ++u) {
        if (TRAIN) {
          fsave[i0 + u] = sg[u];
          gsave[i0 + u] = gh[u];
          xsave[i0 + u] = (ctil[u] - mean) * invstd;
        }
        float cn = __fadd_rn(__fmul_rn(ctil[u], bs), bt);
        c[i0 + u] = cn;
        const bool spike = (i0 + u < nv) && cn >= 0.f;
        hval[i0 + u] = spike ? 1.0f : 0.0f;
        const uint32_t w = __ballot_sync(0xffffffffu, spike);
        myw = (lane % CPT) == (i0 + u) ? w : myw;
      }
    }
    if (p.h_bits && q < WPS && lane < CPT && (int)slice * WPS + q < Wb && row0 + g * CPT + lane < R)
      p.h_bits[((size_t)t * R + row0 + g * CPT + lane) * Wb + slice * WPS + q] = myw;  // the ballots ARE the packed trace
    const long long q4 = PROF ? clock64() : 0;
    // ---- exchange: ONE asynchronous DSMEM store per sending lane into the staging buffer of a CTA of the
    //      cluster; the bytes are counted on the receiver's mbarrier (no fence, no arrive on this side) ----
    if (sender) tc::st_async_u32(par ? snd_cell1 : snd_cell0, myw, par ? snd_bar1 : snd_bar0);
    const long long q5 = PROF ? clock64() : 0;
    if (!tc::mbar_wait_cta(&bar_bits[par], (t >> 1) & 1)) { alive = false; break; }
    const long long q6 = PROF ? clock64() : 0;
    // ---- rebuild the bf16 B operand (spikes of frame t, all H neurons of my rows) from the bits -------
    {
      const uint32_t* src = bits + (size_t)par * NT * KWp;
#pragma unroll
      for (int it = 0; it < MAXT; ++it) {
        if (task_dst[it] == 0xFFFFFFFFu) continue;
        const uint32_t b8 = (src[task_src[it] & 0xFFFFFFu] >> (task_src[it] >> 24)) & 0xFFu;
        uint32_t v[4];
#pragma unroll
        for (int e = 0; e < 4; ++e)
          v[e] = ((b8 >> (2 * e)) & 1u ? kOneBf16 : 0u) | ((b8 >> (2 * e + 1)) & 1u ? (kOneBf16 << 16) : 0u);
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
  // the next chunk's grid (if it was launched as a programmatic dependent) may be scheduled from here on: its CTAs
  // are then already queued, ahead of lower-priority work, when this grid's SMs become free
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const long long e4 = PROF ? clock64() : 0;
  if (PROF && p.prof && blockIdx.x == 0 && tid == 0)
    for (int i = 0; i < 8; ++i) p.prof[i] = (unsigned long long)pc[i];

  store_frame(T - 1);
  if (TRAIN && batch_stats && tile == 0 && g == 0 && jv) {
    p.run_mean[j] = rmean;
    p.run_var[j] = rvar;
  }
#pragma unroll
  for (int i = 0; i < CPT; ++i) {
    if (i < nv) {
      if (p.cT) *reinterpret_cast<float*>(reinterpret_cast<char*>(p.cT) + boff0 + i * hstride) = c[i];
      if (p.hT) *reinterpret_cast<float*>(reinterpret_cast<char*>(p.hT) + boff0 + i * hstride) = hval[i];
    }
  }
  tc::tc_fence_before();
  tc::cluster_sync_all();  // nobody leaves while a peer may still store into its staging buffer
  if (warp == 0) tc::tmem_dealloc<kTmemCols>(tmem);
  if (PROF && p.prof && blockIdx.x == 0 && tid == 0) {  // launch-constant phases: alloc + h0, weights, state, loop, exit
    const long long e5 = clock64();
    p.prof[8] = e1 - e0; p.prof[9] = e2 - e1; p.prof[10] = e3 - e2; p.prof[11] = e4 - e3; p.prof[12] = e5 - e4;
  }
  trace_end(p.trace, tslot);
}

// ------------------------------------------------------------------------------------------------
static int tc_pick_nt(int R, int H, int shared, int sm_count) {
  const int C = shared ? (H + 127) / 128 : (H + 63) / 64;
  const int Kmma = (H + 15) / 16 * 16;
  const int cols_a = kTcPlanes * Kmma / 2;
  int best = 0;
  for (int nt : {16, 32, 64}) {
    if (cols_a + nt > (int)kTmemCols) break;
    if ((nt / 4) * C > 32) break;  // one sending lane per (row of the group, destination CTA)
    best = nt;
    if ((long long)((R + nt - 1) / nt) * C <= sm_count) break;  // whole problem co-resident
  }
  return best;
}

bool recurrence_tc_supported(int R, int H, int shared) {
  if (R <= 0 || H < 16) return false;
  const int C = shared ? (H + 127) / 128 : (H + 63) / 64;
  if (C > 8) return false;  // portable cluster size
  return tc_pick_nt(R, H, shared, 148) > 0;
}

int recurrence_tc_tile(int R, int H, int shared, int sms) { return tc_pick_nt(R, H, shared, sms); }

size_t recurrence_tc_workspace(int, int, int) { return 256; }

template <int NT, int G, bool PROF, bool SHARED, bool TRAIN = false>
static int launch_nt(const RecTcParams& p_in, int C, cudaStream_t st) {
  RecTcParams p = p_in;
  size_t smem = tc_smem_bytes<NT>(p.Kmma, C, SHARED);
  // weight rows staged through shared memory with bulk copies when they are 16-byte aligned and the area fits
  p.wpitch = 0;
  if (p.H % 4 == 0 && (reinterpret_cast<uintptr_t>(p.w_hh) & 15) == 0) {
    const size_t stage = (size_t)128 * tc_wpitch(p.H) * sizeof(float);
    if (smem + stage <= tc::kMaxDynamicSmem) {
      p.wpitch = tc_wpitch(p.H);
      smem += stage;
    }
  }
  if (smem < tc::kTmemExclusiveSmem) smem = tc::kTmemExclusiveSmem;
  GSN_CUDA(cudaFuncSetAttribute(k_recurrence_tc<NT, G, PROF, SHARED, TRAIN>,
                                cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(((p.R + NT - 1) / NT) * C));
  cfg.blockDim = dim3(128 * G);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[3];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = C;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeCooperative;  // training: the per-frame grid barrier needs co-residency
  attr[1].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = TRAIN ? 2 : 1;
  if (!TRAIN && launch_option(GSN_OPT_PDL)) {
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.numAttrs = 2;
  }
  GSN_CUDA(cudaLaunchKernelEx(&cfg, k_recurrence_tc<NT, G, PROF, SHARED, TRAIN>, p));
  return GSN_OK;
}

int launch_recurrence_tc(const float* xproj, const float* w_hh, const float* bias, const float* bn_scale,
                         const float* bn_shift, const float* h0, const float* c0, float* h_out, float* c_out,
                         float* hT, float* cT, int T, int R, int H, int shared, int sm_budget, void* workspace,
                         uint32_t* h_bits, cudaStream_t st) {
  int dev = 0, sms = 148;
  GSN_CUDA(cudaGetDevice(&dev));
  GSN_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  RecTcParams p{};
  p.xproj = xproj; p.w_hh = w_hh; p.bias = bias; p.bn_scale = bn_scale; p.bn_shift = bn_shift; p.h0 = h0; p.c0 = c0;
  p.h_out = h_out; p.c_out = c_out; p.hT = hT; p.cT = cT; p.T = T; p.R = R; p.H = H; p.Kmma = (H + 15) / 16 * 16;
  p.h_bits = h_bits;
  p.prof = reinterpret_cast<unsigned long long*>(workspace); p.trace = trace_buffer();
  const int C = shared ? (H + 127) / 128 : (H + 63) / 64;
  if (sm_budget > 0 && sm_budget < sms) sms = sm_budget;
  const int nt = tc_pick_nt(R, H, shared, sms);
  static const bool prof = getenv("GSN_TC_PROF") != nullptr;  // dev knob: per-phase cycle counters
  p.dbg = getenv("GSN_TC_DBG") ? atoi(getenv("GSN_TC_DBG")) : 0;
  if (!shared) {
    switch (nt) {
      case 16: return launch_nt<16, 4, false, false>(p, C, st);
      case 32: return launch_nt<32, 4, false, false>(p, C, st);
      case 64: return launch_nt<64, 4, false, false>(p, C, st);
      default: return fail(GSN_ENOSUP, "gsn_layer_recurrence(TCGEN05): H=%d does not fit tensor memory", H);
    }
  }
  switch (nt) {
    case 16: return prof ? launch_nt<16, 4, true, true>(p, C, st) : launch_nt<16, 4, false, true>(p, C, st);
    case 32: return prof ? launch_nt<32, 4, true, true>(p, C, st) : launch_nt<32, 4, false, true>(p, C, st);
    case 64: return prof ? launch_nt<64, 4, true, true>(p, C, st) : launch_nt<64, 4, false, true>(p, C, st);
    default: return fail(GSN_ENOSUP, "gsn_layer_recurrence(TCGEN05): H=%d does not fit tensor memory", H);
  }
}

// ---- training forward on tcgen05 (shared gate weights): same kernel with TRAIN = true, cooperative launch ----
bool recurrence_tc_train_supported(int R, int H, int shared) {
  return shared && H % 4 == 0 && recurrence_tc_supported(R, H, shared);
}

size_t recurrence_tc_train_workspace(int R, int H) {
  const size_t C = (H + 127) / 128, tiles = (R + 15) / 16;
  return (2 * tiles * 2 * C * 128) * sizeof(float) + 256;
}

int launch_recurrence_tc_train(const float* xproj, const float* w_hh, const float* bias, const float* bn_w,
                               const float* bn_b, float* run_mean, float* run_var, float* h_out, float* c_out,
                               float* f_out, float* g_out, float* xhat_out, float* invstd_out, int T, int R, int H,
                               int training, float momentum, float eps, int sm_budget, void* workspace,
                               cudaStream_t st) {
  int dev = 0, sms = 148;
  GSN_CUDA(cudaGetDevice(&dev));
  GSN_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int C = (H + 127) / 128;
  const int all_sms = sms;
  if (sm_budget > 0 && sm_budget < sms) sms = sm_budget;
  int nt = tc_pick_nt(R, H, 1, sms);
  // one CTA per SM (TMEM-exclusive shared memory request): the cooperative grid must fit the device
  while (nt > 0 && nt < 64 && (long long)((R + nt - 1) / nt) * C > all_sms) nt *= 2;
  if (nt == 0 || (long long)((R + nt - 1) / nt) * C > all_sms)
    return fail(GSN_ENOSUP, "gsn_layer_train_forward_tc: R=%d H=%d does not fit one cooperative grid", R, H);
  RecTcParams p{};
  p.xproj = xproj; p.w_hh = w_hh; p.bias = bias; p.h_out = h_out; p.c_out = c_out; p.T = T; p.R = R; p.H = H;
  p.Kmma = (H + 15) / 16 * 16; p.trace = trace_buffer();
  p.bn_w = bn_w; p.bn_b = bn_b; p.run_mean = run_mean; p.run_var = run_var; p.f_out = f_out; p.g_out = g_out;
  p.xhat_out = xhat_out; p.invstd_out = invstd_out; p.momentum = momentum; p.eps = eps;
  p.batch_stats = (bn_w != nullptr && training) ? 1 : 0;
  const size_t tiles = (R + nt - 1) / nt;
  p.partial = reinterpret_cast<float*>(workspace);
  p.counter = reinterpret_cast<unsigned int*>(p.partial + 2 * tiles * 2 * C * 128);
  GSN_CUDA(cudaMemsetAsync(p.counter, 0, 64, st));
  switch (nt) {
    case 16: return launch_nt<16, 4, false, true, true>(p, C, st);
    case 32: return launch_nt<32, 4, false, true, true>(p, C, st);
    case 64: return launch_nt<64, 4, false, true, true>(p, C, st);
    default: return fail(GSN_ENOSUP, "gsn_layer_train_forward_tc: unsupported tile");
  }
}

}  // namespace gsn
