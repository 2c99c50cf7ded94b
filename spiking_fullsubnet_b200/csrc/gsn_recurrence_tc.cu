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
//     shared memory (st.shared::cluster) + a remote mbarrier arrive; each CTA then expands the bits of
//     its NT rows x H neurons back into the bf16 B operand for frame t+1.
// One 128-thread warpgroup does everything; thread 0 issues the MMAs (single-thread tcgen05.mma).
#include "gsn_common.cuh"
#include "gsn_tc.cuh"

namespace gsn {

struct RecTcParams {
  const float* xproj;   // [T, R, H]
  const float* w_hh;    // [H, H]
  const float* bias;    // [2H]
  const float* bn_scale;
  const float* bn_shift;
  const float* h0;
  const float* c0;
  float* h_out;         // [T, R, H]
  float* c_out;         // [T, R, H] or null
  float* hT;
  float* cT;
  int T, R, H, Kmma;    // Kmma = round_up(H, 16)
};

constexpr int kTcPlanes = 3;
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kOneBf16 = 0x3F80u;

__host__ __device__ inline int tc_kw_padded(int C) { return 4 * C + 1; }  // odd stride: conflict-free reads

template <int NT>
__host__ __device__ inline size_t tc_smem_bytes(int Kmma, int C) {
  size_t b = (size_t)NT * Kmma * 2;                       // B operand
  b = (b + 127) / 128 * 128;
  b += (size_t)2 * NT * tc_kw_padded(C) * 4;              // bit staging, double buffered
  b = (b + 15) / 16 * 16;
  b += 64;                                                // barriers + tmem slot
  return b;
}

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

template <int NT>
__global__ void __launch_bounds__(128, 1) k_recurrence_tc(const RecTcParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t C = tc::cluster_nctarank(), slice = tc::cluster_ctarank();
  const int row0 = (blockIdx.x / C) * NT;
  const int H = p.H, R = p.R, T = p.T, Kmma = p.Kmma;
  const int j = slice * 128 + tid;  // this thread's neuron
  const bool jv = j < H;
  const int KWp = tc_kw_padded(C);

  uint8_t* sB = smem;
  size_t off = ((size_t)NT * Kmma * 2 + 127) / 128 * 128;
  uint32_t* bits = reinterpret_cast<uint32_t*>(smem + off);  // [2][NT][KWp]
  off += (size_t)2 * NT * KWp * 4;
  off = (off + 15) / 16 * 16;
  uint64_t* bar_mma = reinterpret_cast<uint64_t*>(smem + off);
  uint64_t* bar_bits = bar_mma + 1;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_mma + 3);

  if (tid == 0) {
    tc::mbar_init(bar_mma, 1);
    tc::mbar_init(&bar_bits[0], 4 * C);
    tc::mbar_init(&bar_bits[1], 4 * C);
    tc::fence_mbar_init();
  }
  if (warp == 0) tc::tmem_alloc<kTmemCols>(tmem_slot);

  // B operand of frame 0: the initial spikes h0 (zeros when null).  byte(n,k) = (n/8)*SBO + (k/8)*128 + (n%8)*16
  const uint32_t SBO = 16u * Kmma;
  const int k8n = Kmma / 8;
  for (int i = tid; i < NT * k8n; i += 128) {
    const int nlo = i & 7, k8 = (i >> 3) % k8n, nhi = (i >> 3) / k8n;
    const int n = nhi * 8 + nlo, row = row0 + n;
    uint32_t v[4] = {0, 0, 0, 0};
    if (p.h0 && row < R) {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int k = k8 * 8 + e;
        if (k < H && p.h0[(size_t)row * H + k] != 0.f) v[e >> 1] |= kOneBf16 << (16 * (e & 1));
      }
    }
    *reinterpret_cast<uint4*>(sB + (size_t)nhi * SBO + k8 * 128 + nlo * 16) = make_uint4(v[0], v[1], v[2], v[3]);
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
  const uint32_t tmem_d = tmem;                       // accumulator: columns [0, NT)
  const uint32_t tmem_a = tmem + NT;                  // plane pl: columns [NT + pl*Kmma/2, ...)
  const uint32_t plane_cols = Kmma / 2;

  // recurrent weights of this thread's neuron -> three exact bf16 planes in TMEM (lane = neuron)
  {
    const float* wrow = p.w_hh + (size_t)(jv ? j : 0) * H;
    for (int c0 = 0; c0 < (int)plane_cols; c0 += 8) {
      uint32_t vh[8], vm[8], vl[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        uint32_t h2[2], m2[2], l2[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int k = 2 * (c0 + q) + e;
          const float w = (jv && k < H) ? __ldg(wrow + k) : 0.f;
          split3(w, h2[e], m2[e], l2[e]);
        }
        vh[q] = h2[0] | (h2[1] << 16);
        vm[q] = m2[0] | (m2[1] << 16);
        vl[q] = l2[0] | (l2[1] << 16);
      }
      tc::tmem_st8(tmem_a + lane_base + 0 * plane_cols + c0, vl);  // plane 0 = lo (issued first)
      tc::tmem_st8(tmem_a + lane_base + 1 * plane_cols + c0, vm);
      tc::tmem_st8(tmem_a + lane_base + 2 * plane_cols + c0, vh);
    }
    tc::tmem_wait_st();
  }

  const int jj = jv ? j : 0;
  const float bf = p.bias[jj], bc = p.bias[H + jj];
  const float bs = p.bn_scale ? p.bn_scale[jj] : 1.0f;
  const float bt = p.bn_shift ? p.bn_shift[jj] : 0.0f;
  float c[NT];
#pragma unroll
  for (int n = 0; n < NT; ++n) {
    const int row = row0 + n;
    c[n] = (p.c0 && jv && row < R) ? p.c0[(size_t)row * H + j] : 0.f;
  }
  // all barriers of the cluster are initialised before anybody arrives remotely
  tc::cluster_sync_all();

  const uint32_t idesc = tc::make_idesc_f16(128, NT, true);
  const uint64_t desc_b0 = tc::make_smem_desc(tc::smem_u32(sB), 128, SBO);
  const int ksteps = Kmma / 16;
  bool alive = true;

  for (int t = 0; t < T; ++t) {
    // ---- recurrent product of frame t: D = W_hh[slice] . h_{t-1}^T --------------------------------
    tc::fence_proxy_async_smem();  // B operand was written through the generic proxy
    tc::tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc::tc_fence_after();
      uint32_t acc = 0;
#pragma unroll 1
      for (int pl = 0; pl < kTcPlanes; ++pl) {
        const uint32_t a0 = tmem_a + pl * plane_cols;
#pragma unroll 4
        for (int ks = 0; ks < ksteps; ++ks) {
          tc::mma_ts(tmem_d, a0 + ks * 8, desc_b0 + (uint64_t)(ks * 16), idesc, acc);
          acc = 1;
        }
      }
      tc::mma_commit(bar_mma);
    }
    // input projection of frame t for my neuron, all rows (coalesced over neurons); overlaps the MMAs
    float xp[NT];
#pragma unroll
    for (int n = 0; n < NT; ++n) {
      const int row = row0 + n;
      xp[n] = (jv && row < R) ? __ldg(p.xproj + ((size_t)t * R + row) * H + j) : 0.f;
    }
    if (!tc::mbar_wait(bar_mma, t & 1)) { alive = false; break; }
    tc::tc_fence_after();

    // ---- leak / BatchNorm / threshold; spikes -> bit words ----------------------------------------
    const int par = t & 1;
    uint32_t myword[(NT + 31) / 32];
#pragma unroll
    for (int q = 0; q < (NT + 31) / 32; ++q) myword[q] = 0;
#pragma unroll
    for (int n0 = 0; n0 < NT; n0 += 16) {
      uint32_t zr[16];
      tc::tmem_ld16(tmem_d + lane_base + n0, zr);
      tc::tmem_wait_ld();
#pragma unroll
      for (int q = 0; q < 16; ++q) {
        const int n = n0 + q, row = row0 + n;
        const float z = __uint_as_float(zr[q]);
        // reference order: (x W_ih^T + bias) + h W_hh^T   (ESN:140-145)
        const float f_hat = __fadd_rn(__fadd_rn(xp[n], bf), z);
        const float g_hat = __fadd_rn(__fadd_rn(xp[n], bc), z);
        const float cn = gsu_membrane(f_hat, g_hat, c[n], bs, bt);
        c[n] = cn;
        const bool ok = jv && row < R;
        const bool spike = ok && cn >= 0.f;
        if (ok) {
          const size_t o = ((size_t)t * R + row) * H + j;
          p.h_out[o] = spike ? 1.0f : 0.0f;
          if (p.c_out) p.c_out[o] = cn;
        }
        const uint32_t w = __ballot_sync(0xffffffffu, spike);
        if (lane == (n & 31)) myword[n >> 5] = w;
      }
    }
    // ---- exchange the bits with every CTA of the cluster (DSMEM), then rebuild the B operand -------
    {
      uint32_t* dst = bits + (size_t)par * NT * KWp;
#pragma unroll
      for (int q = 0; q < (NT + 31) / 32; ++q) {
        const int n = q * 32 + lane;
        if (n < NT) {
          uint32_t* cell = dst + n * KWp + slice * 4 + warp;
          for (uint32_t r = 0; r < C; ++r) tc::st_cluster_u32(tc::map_to_rank(cell, r), myword[q]);
        }
      }
      __syncwarp();
      if (lane == 0)
        for (uint32_t r = 0; r < C; ++r) tc::mbar_arrive_cluster(&bar_bits[par], r);
    }
    if (!tc::mbar_wait(&bar_bits[par], (t >> 1) & 1)) { alive = false; break; }
    {
      const uint32_t* src = bits + (size_t)par * NT * KWp;
      for (int i = tid; i < NT * k8n; i += 128) {
        const int nlo = i & 7, k8 = (i >> 3) % k8n, nhi = (i >> 3) / k8n;
        const int n = nhi * 8 + nlo;
        const uint32_t b8 = (src[n * KWp + (k8 >> 2)] >> (8 * (k8 & 3))) & 0xFFu;
        uint32_t v[4];
#pragma unroll
        for (int e = 0; e < 4; ++e)
          v[e] = ((b8 >> (2 * e)) & 1u ? kOneBf16 : 0u) | ((b8 >> (2 * e + 1)) & 1u ? (kOneBf16 << 16) : 0u);
        *reinterpret_cast<uint4*>(sB + (size_t)nhi * SBO + k8 * 128 + nlo * 16) = make_uint4(v[0], v[1], v[2], v[3]);
      }
    }
  }
  if (!alive) __trap();  // a broken pipeline fails loudly instead of hanging the device

  if (jv) {
#pragma unroll
    for (int n = 0; n < NT; ++n) {
      const int row = row0 + n;
      if (row < R) {
        if (p.cT) p.cT[(size_t)row * H + j] = c[n];
        if (p.hT) p.hT[(size_t)row * H + j] = p.h_out[((size_t)(T - 1) * R + row) * H + j];
      }
    }
  }
  tc::tc_fence_before();
  tc::cluster_sync_all();  // nobody leaves while a peer may still write into its staging buffer
  if (warp == 0) tc::tmem_dealloc<kTmemCols>(tmem);
}

// ------------------------------------------------------------------------------------------------
static int tc_pick_nt(int R, int H, int sm_count) {
  const int C = (H + 127) / 128;
  const int Kmma = (H + 15) / 16 * 16;
  const int cols_a = kTcPlanes * Kmma / 2;
  int best = 0;
  for (int nt : {16, 32, 64}) {
    if (cols_a + nt > (int)kTmemCols) break;
    best = nt;
    if ((long long)((R + nt - 1) / nt) * C <= sm_count) break;  // whole problem co-resident
  }
  return best;
}

bool recurrence_tc_supported(int R, int H, int shared) {
  if (!shared || R <= 0 || H < 16) return false;
  const int C = (H + 127) / 128;
  if (C > 8) return false;
  return tc_pick_nt(R, H, 148) > 0;
}

size_t recurrence_tc_workspace(int, int, int) { return 256; }

template <int NT>
static int launch_nt(const RecTcParams& p, int C, cudaStream_t st) {
  const size_t smem = tc_smem_bytes<NT>(p.Kmma, C);
  GSN_CUDA(cudaFuncSetAttribute(k_recurrence_tc<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(((p.R + NT - 1) / NT) * C));
  cfg.blockDim = dim3(128);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = C;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  GSN_CUDA(cudaLaunchKernelEx(&cfg, k_recurrence_tc<NT>, p));
  return GSN_OK;
}

int launch_recurrence_tc(const float* xproj, const float* w_hh, const float* bias, const float* bn_scale,
                         const float* bn_shift, const float* h0, const float* c0, float* h_out, float* c_out,
                         float* hT, float* cT, int T, int R, int H, int shared, void*, cudaStream_t st) {
  GSN_REQUIRE(shared, "gsn_layer_recurrence(TCGEN05): unshared gate weights are not supported");
  int dev = 0, sms = 148;
  GSN_CUDA(cudaGetDevice(&dev));
  GSN_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  RecTcParams p{xproj, w_hh, bias, bn_scale, bn_shift, h0, c0, h_out, c_out, hT, cT, T, R, H, (H + 15) / 16 * 16};
  const int C = (H + 127) / 128;
  const int nt = tc_pick_nt(R, H, sms);
  switch (nt) {
    case 16: return launch_nt<16>(p, C, st);
    case 32: return launch_nt<32>(p, C, st);
    case 64: return launch_nt<64>(p, C, st);
    default: return fail(GSN_ENOSUP, "gsn_layer_recurrence(TCGEN05): H=%d does not fit tensor memory", H);
  }
}

}  // namespace gsn
