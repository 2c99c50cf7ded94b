// Training path of the GSN recurrence (fp32, CUDA cores, cooperative launch):
//   forward  : GSUCell.forward ESN:132-153 with nn.BatchNorm1d in TRAINING mode (per-frame batch statistics over
//              all R rows of the sequence model, running statistics updated every frame, ESN:149-150), saving
//              what BPTT needs;
//   backward : BPTT through the frames with the Triangle surrogate gradient max(0, 1-|c|) (ESN:95-101) and the
//              batch-statistics BatchNorm backward; emits dL/d(gate pre-activations) per frame, from which the
//              host forms dW_hh, dW_ih, dx with three large GEMMs.
// Batch statistics couple every row of a frame, so each frame needs one grid-wide reduction: CTAs write their
// partial sums, cross a grid barrier (cooperative launch guarantees co-residency) and every CTA reduces the
// partials in the same fixed order (deterministic, identical on all CTAs).
// Layout as k_recurrence_simt: one CTA owns RT = 8 rows, thread j owns neuron j of those rows.
#include <stdlib.h>

#include "gsn_common.cuh"

namespace gsn {

constexpr int TR_RT = 8;

// acc[r] += sum_k s[k][r] * W[k][col]  (and acc2 with column col2) for the RT rows of this CTA, thread = column.
// W [Kdim][ld] fp32 in global memory is either RESIDENT in shared memory (`wsm` holds all of it) or streamed
// through `wsm` in 16-row slabs with cp.async double buffering (coalesced 16-byte copies, L2 -> smem), instead
// of one 4-byte read-only load per thread per k.
constexpr int TR_KB = 16;
constexpr int TR_NST = 4;
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)),
               "l"(gsrc)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void stage_rows(float* dst, const float* W, int k0, int nrows, int ld) {
  const int chunks = nrows * (ld / 4);  // 16-byte chunks, rows are contiguous in both places
  for (int i = threadIdx.x; i < chunks; i += blockDim.x) cp_async16(dst + 4 * i, W + (size_t)k0 * ld + 4 * i);
}

template <int RT, bool TWO>
__device__ __forceinline__ void rows_times_w(const float* __restrict__ W, int Kdim, int ld, int col, int col2,
                                             bool active, const float* s, float* wsm, bool resident,
                                             float (&acc)[RT], float (&acc2)[RT]) {
  auto fma_rows = [&](const float* wrow, int k) {
    const float w1 = wrow[col];
    const float w2 = TWO ? wrow[col2] : 0.f;
    float sv[RT];
#pragma unroll
    for (int q4 = 0; q4 < RT / 4; ++q4) {
      const float4 s4 = *reinterpret_cast<const float4*>(s + k * RT + 4 * q4);
      sv[4 * q4] = s4.x; sv[4 * q4 + 1] = s4.y; sv[4 * q4 + 2] = s4.z; sv[4 * q4 + 3] = s4.w;
    }
#pragma unroll
    for (int r = 0; r < RT; ++r) {
      acc[r] = fmaf(sv[r], w1, acc[r]);
      if (TWO) acc2[r] = fmaf(sv[r], w2, acc2[r]);
    }
  };
  if (resident) {
    if (active) {
#pragma unroll 8
      for (int k = 0; k < Kdim; ++k) fma_rows(wsm + (size_t)k * ld, k);
    }
    return;
  }
  // TR_NST-deep ring of TR_KB-row slabs: TR_NST-1 slabs (L2 round trips) are in flight while one is multiplied
  const int nslab = (Kdim + TR_KB - 1) / TR_KB;
#pragma unroll
  for (int st = 0; st < TR_NST - 1; ++st) {
    if (st < nslab) stage_rows(wsm + (size_t)st * TR_KB * ld, W, st * TR_KB, min(TR_KB, Kdim - st * TR_KB), ld);
    cp_async_commit();
  }
  for (int sidx = 0; sidx < nslab; ++sidx) {
    const int k0 = sidx * TR_KB;
    const int pre = sidx + TR_NST - 1;
    if (pre < nslab)
      stage_rows(wsm + (size_t)(pre % TR_NST) * TR_KB * ld, W, pre * TR_KB, min(TR_KB, Kdim - pre * TR_KB), ld);
    cp_async_commit();
    cp_async_wait<TR_NST - 1>();
    __syncthreads();  // slab sidx has landed for every thread
    if (active) {
      const float* cur = wsm + (size_t)(sidx % TR_NST) * TR_KB * ld;
      const int kn = min(TR_KB, Kdim - k0);
#pragma unroll 8
      for (int kk = 0; kk < kn; ++kk) fma_rows(cur + (size_t)kk * ld, k0 + kk);
    }
    __syncthreads();  // everybody is done with this buffer before it is refilled
  }
  cp_async_wait<0>();
}

// deterministic cross-CTA sum of the per-CTA partials of one frame: loads are issued 8 at a time, the adds keep
// the fixed order b = 0, 1, 2, ...
__device__ __forceinline__ void reduce_partials(const float* part, unsigned int nblocks, int H, int j, bool active,
                                                float& a1, float& a2) {
  a1 = 0.f;
  a2 = 0.f;
  if (!active) return;
  for (unsigned int b0 = 0; b0 < nblocks; b0 += 8) {
    float v1[8], v2[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const bool ok = b0 + u < nblocks;
      v1[u] = ok ? __ldcg(part + ((size_t)(b0 + u) * 2 + 0) * H + j) : 0.f;
      v2[u] = ok ? __ldcg(part + ((size_t)(b0 + u) * 2 + 1) * H + j) : 0.f;
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      a1 += v1[u];
      a2 += v2[u];
    }
  }
}

struct TrainFwdParams {
  const float* xproj;   // [T,R,gH]
  const float* wt;      // [H,gH]  transposed recurrent weights (workspace)
  const float* bias;    // [2H]
  const float* bn_w;    // [H] or null (no BatchNorm)
  const float* bn_b;
  float* run_mean;      // [H] updated in place (training) / read (eval)
  float* run_var;
  float* h_out;         // [T,R,H]
  float* c_out;         // [T,R,H] membrane potential after BatchNorm (the carried state)
  float* f_out;         // [T,R,H] sigmoid(forget gate)
  float* g_out;         // [T,R,H] cell-gate pre-activation
  float* xhat_out;      // [T,R,H] normalised pre-BN membrane (training) / unused
  float* invstd_out;    // [T,H]   (training)
  float* partial;       // [2][nblocks][2][H] scratch
  unsigned int* counter;
  int T, R, H, training;
  float momentum, eps;
  int resident;         // the transposed weights fit shared memory next to the spikes
  long long* prof;      // [6] cycle counters of CTA 0 (development aid, GSN_TRAIN_PROF) or null
};

template <bool SHARED>
__global__ void __launch_bounds__(512) k_rec_train_fwd(const TrainFwdParams p) {
  extern __shared__ __align__(16) float sh[];  // [2][H][RT] spikes, then the weight slabs / resident weights
  constexpr int RT = TR_RT;
  const int H = p.H, R = p.R, T = p.T;
  const int gH = SHARED ? H : 2 * H;
  const int j = threadIdx.x;
  const int row0 = blockIdx.x * RT;
  const bool active = j < H;
  const int jj = active ? j : 0;
  const unsigned int nblocks = gridDim.x;
  float* wsm = sh + (size_t)2 * H * RT;
  const bool resident = p.resident != 0;
  if (resident) {  // the whole transposed weight matrix fits shared memory: load it once
    stage_rows(wsm, p.wt, 0, H, gH);
    cp_async_commit();
    cp_async_wait<0>();
  }
  const float bf = p.bias[jj], bc = p.bias[H + jj];
  const bool bn = p.bn_w != nullptr;
  const float gam = bn ? p.bn_w[jj] : 1.f, bet = bn ? p.bn_b[jj] : 0.f;
  float rmean = bn ? p.run_mean[jj] : 0.f, rvar = bn ? p.run_var[jj] : 1.f;
  const bool batch_stats = bn && p.training;
  float alpha = 1.f, beta = 0.f;  // eval-mode fold (torch's CPU kernel: alpha = w * invstd, beta = b - mean*alpha)
  if (bn && !p.training) {
    alpha = gam * (1.0f / sqrtf(rvar + p.eps));
    beta = bet - rmean * alpha;
  }
  float c[RT];
#pragma unroll
  for (int r = 0; r < RT; ++r) {
    c[r] = 0.f;
    if (active) sh[j * RT + r] = 0.f;
  }
  unsigned int epoch = 0;
  float shift = 0.f;  // shift for the one-pass variance: the previous frame's mean
  __syncthreads();

  long long pc[6] = {0, 0, 0, 0, 0, 0};
  for (int t = 0; t < T; ++t) {
    const long long q0 = p.prof ? clock64() : 0;
    const float* cur = sh + (size_t)(t & 1) * H * RT;
    float* nxt = sh + (size_t)((t & 1) ^ 1) * H * RT;
    float xf[RT], xg[RT];
#pragma unroll
    for (int r = 0; r < RT; ++r) {
      const int row = row0 + r;
      const bool ok = active && row < R;
      const size_t base = ((size_t)t * R + (ok ? row : 0)) * gH;
      xf[r] = ok ? p.xproj[base + j] : 0.f;
      xg[r] = SHARED ? xf[r] : (ok ? p.xproj[base + H + j] : 0.f);
    }
    float af[RT], ag[RT];
#pragma unroll
    for (int r = 0; r < RT; ++r) { af[r] = 0.f; ag[r] = 0.f; }
    rows_times_w<RT, !SHARED>(p.wt, H, gH, jj, H + jj, active, cur, wsm, resident, af, ag);
    const long long q1 = p.prof ? clock64() : 0;
    float fv[RT], gv[RT], ct[RT];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int r = 0; r < RT; ++r) {
      const float f_hat = __fadd_rn(__fadd_rn(xf[r], bf), af[r]);
      gv[r] = __fadd_rn(__fadd_rn(xg[r], bc), SHARED ? af[r] : ag[r]);
      fv[r] = sigmoid_f32(f_hat);
      ct[r] = __fadd_rn(__fmul_rn(fv[r], c[r]), __fmul_rn(__fsub_rn(1.0f, fv[r]), gv[r]));
      if (active && row0 + r < R) {
        const float d = ct[r] - shift;
        s1 += d;
        s2 += d * d;
      }
    }
    float mean = 0.f, invstd = 1.f;
    long long q2 = 0, q3 = 0;
    if (batch_stats) {
      float* part = p.partial + (size_t)(t & 1) * nblocks * 2 * H;
      if (active) {
        part[((size_t)blockIdx.x * 2 + 0) * H + j] = s1;
        part[((size_t)blockIdx.x * 2 + 1) * H + j] = s2;
      }
      q2 = p.prof ? clock64() : 0;
      if (!grid_barrier(p.counter, nblocks, epoch)) __trap();
      q3 = p.prof ? clock64() : 0;
      float a1, a2;
      reduce_partials(part, nblocks, H, j, active, a1, a2);
      const float m1 = a1 / (float)R;                 // E[x - shift]
      const float var = fmaxf(a2 / (float)R - m1 * m1, 0.f);  // biased variance
      mean = shift + m1;
      invstd = 1.0f / sqrtf(var + p.eps);
      alpha = gam * invstd;
      beta = bet - mean * alpha;
      // running statistics (unbiased variance), every frame (ESN:149-150 calls BatchNorm once per frame)
      rmean = (1.f - p.momentum) * rmean + p.momentum * mean;
      rvar = (1.f - p.momentum) * rvar + p.momentum * (var * (float)R / (float)(R - 1));
      shift = mean;
      if (blockIdx.x == 0 && active) p.invstd_out[(size_t)t * H + j] = invstd;
    }
    const long long q4 = p.prof ? clock64() : 0;
#pragma unroll
    for (int r = 0; r < RT; ++r) {
      const int row = row0 + r;
      const float cn = __fadd_rn(__fmul_rn(ct[r], alpha), beta);
      c[r] = cn;
      const float h = cn >= 0.f ? 1.0f : 0.0f;
      if (active) {
        nxt[j * RT + r] = (row < R) ? h : 0.f;
        if (row < R) {
          const size_t o = ((size_t)t * R + row) * H + j;
          p.h_out[o] = h;
          p.c_out[o] = cn;
          if (p.f_out) {
            p.f_out[o] = fv[r];
            p.g_out[o] = gv[r];
            if (batch_stats) p.xhat_out[o] = (ct[r] - mean) * invstd;
          }
        }
      }
    }
    __syncthreads();
    if (p.prof) {
      const long long q5 = clock64();
      pc[0] += q1 - q0; pc[1] += q2 - q1; pc[2] += q3 - q2; pc[3] += q4 - q3; pc[4] += q5 - q4; pc[5] += q5 - q0;
    }
  }
  if (p.prof && blockIdx.x == 0 && j == 0)
    for (int i = 0; i < 6; ++i) p.prof[i] = pc[i];
  if (batch_stats && blockIdx.x == 0 && active) {
    p.run_mean[j] = rmean;
    p.run_var[j] = rvar;
  }
}

struct TrainBwdParams {
  const float* dh_out;   // [T,R,H] dL/dh_t from above
  const float* w_hh;     // [gH,H]
  const float* c;        // [T,R,H] post-BN membrane (forward output)
  const float* f;        // [T,R,H]
  const float* g;        // [T,R,H]
  const float* xhat;     // [T,R,H] (batch statistics) or null
  const float* invstd;   // [T,H]   (batch statistics) or null
  const float* bn_w;     // [H] or null
  const float* run_var;  // [H] (eval-mode BN) or null
  float* dz;             // [T,R,gH] dL/d(gate pre-activation) = dL/dxproj
  float* dbias_part;     // [nblocks][2H]
  float* dgamma;         // [H] (written by CTA 0; batch statistics only)
  float* dbeta;          // [H]
  float* partial;        // [2][nblocks][2][H]
  unsigned int* counter;
  int T, R, H, training;
  float eps;
  int resident;
};

// RT = 8 rows per CTA (any H <= 512) or RT = 16 (H <= 256: half the CTAs -> half the weight streaming and a
// quarter of the all-to-all partial traffic, which is what bounds large row counts)
template <bool SHARED, int RT>
__global__ void __launch_bounds__(RT == 8 ? 512 : 256) k_rec_train_bwd(const TrainBwdParams p) {
  extern __shared__ __align__(16) float sh[];  // [2][gH][RT] dz of the later frame, then weight slabs / weights
  const int H = p.H, R = p.R, T = p.T;
  const int gH = SHARED ? H : 2 * H;
  const int j = threadIdx.x;
  const int row0 = blockIdx.x * RT;
  const bool active = j < H;
  const int jj = active ? j : 0;
  const unsigned int nblocks = gridDim.x;
  float* wsm = sh + (size_t)2 * gH * RT;
  const bool resident = p.resident != 0;
  if (resident) {
    stage_rows(wsm, p.w_hh, 0, gH, H);
    cp_async_commit();
    cp_async_wait<0>();
  }
  const bool bn = p.bn_w != nullptr;
  const bool batch_stats = bn && p.training;
  const float gam = bn ? p.bn_w[jj] : 1.f;
  const float eval_scale = (bn && !p.training) ? gam * (1.0f / sqrtf(p.run_var[jj] + p.eps)) : 1.f;
  float dcn[RT];
#pragma unroll
  for (int r = 0; r < RT; ++r) dcn[r] = 0.f;
  for (int i = j; i < 2 * gH * RT; i += blockDim.x) sh[i] = 0.f;
  float acc_bf = 0.f, acc_bc = 0.f, acc_dg = 0.f, acc_db = 0.f;
  unsigned int epoch = 0;
  __syncthreads();

  for (int t = T - 1; t >= 0; --t) {
    const float* later = sh + (size_t)(t & 1) * gH * RT;        // dz_{t+1} (zeros for t = T-1)
    float* mine = sh + (size_t)((t & 1) ^ 1) * gH * RT;         // dz_t for frame t-1
    // dh_t = dL/dh_t (from above) + dz_{t+1} @ W_hh   (thread j = column j of W_hh, coalesced)
    float dh[RT], dh_unused[RT];
#pragma unroll
    for (int r = 0; r < RT; ++r) dh[r] = 0.f;
    rows_times_w<RT, false>(p.w_hh, gH, H, jj, 0, active, later, wsm, resident, dh, dh_unused);
    float dc[RT], xh[RT], fv[RT], gv[RT], cp[RT];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int r = 0; r < RT; ++r) {
      const int row = row0 + r;
      const bool ok = active && row < R;
      const size_t o = ((size_t)t * R + (ok ? row : 0)) * H + jj;
      const float ct = ok ? p.c[o] : 0.f;
      fv[r] = ok ? p.f[o] : 0.f;
      gv[r] = ok ? p.g[o] : 0.f;
      cp[r] = (ok && t > 0) ? p.c[o - (size_t)R * H] : 0.f;
      xh[r] = (ok && batch_stats) ? p.xhat[o] : 0.f;
      const float up = ok ? p.dh_out[o] : 0.f;
      dc[r] = ok ? dcn[r] + (up + dh[r]) * fmaxf(0.f, 1.0f - fabsf(ct)) : 0.f;  // Triangle surrogate
      s1 += dc[r];
      s2 += dc[r] * xh[r];
    }
    float m1 = 0.f, m2 = 0.f, scale = eval_scale;
    if (batch_stats) {
      float* part = p.partial + (size_t)(t & 1) * nblocks * 2 * H;
      if (active) {
        part[((size_t)blockIdx.x * 2 + 0) * H + j] = s1;
        part[((size_t)blockIdx.x * 2 + 1) * H + j] = s2;
      }
      if (!grid_barrier(p.counter, nblocks, epoch)) __trap();
      float a1, a2;
      reduce_partials(part, nblocks, H, j, active, a1, a2);
      acc_db += a1;
      acc_dg += a2;
      m1 = a1 / (float)R;
      m2 = a2 / (float)R;
      scale = gam * (active ? p.invstd[(size_t)t * H + j] : 1.f);
    }
#pragma unroll
    for (int r = 0; r < RT; ++r) {
      const int row = row0 + r;
      const bool ok = active && row < R;
      const float dct = ok ? scale * (dc[r] - m1 - xh[r] * m2) : 0.f;  // through BatchNorm
      const float one_f = 1.0f - fv[r];
      const float df = dct * (cp[r] - gv[r]) * fv[r] * one_f;
      const float dg = dct * one_f;
      dcn[r] = dct * fv[r];
      acc_bf += df;
      acc_bc += dg;
      if (active) {
        if (SHARED) {
          mine[j * RT + r] = df + dg;
          if (ok) p.dz[((size_t)t * R + row) * gH + j] = df + dg;
        } else {
          mine[j * RT + r] = df;
          mine[(H + j) * RT + r] = dg;
          if (ok) {
            p.dz[((size_t)t * R + row) * gH + j] = df;
            p.dz[((size_t)t * R + row) * gH + H + j] = dg;
          }
        }
      }
    }
    __syncthreads();
  }
  if (active) {
    p.dbias_part[(size_t)blockIdx.x * 2 * H + j] = acc_bf;
    p.dbias_part[(size_t)blockIdx.x * 2 * H + H + j] = acc_bc;
    if (batch_stats && blockIdx.x == 0) {
      p.dgamma[j] = acc_dg;
      p.dbeta[j] = acc_db;
    }
  }
}

__global__ void k_transpose_f32(const float* __restrict__ w, float* __restrict__ wt, int rows, int cols) {
  __shared__ float tile[32][33];
  const int x0 = blockIdx.x * 32, y0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = y0 + i, c = x0 + threadIdx.x;
    tile[i][threadIdx.x] = (r < rows && c < cols) ? w[(size_t)r * cols + c] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = x0 + i, c = y0 + threadIdx.x;
    if (r < cols && c < rows) wt[(size_t)r * rows + c] = tile[threadIdx.x][i];
  }
}

template <typename K, typename P>
static int coop_launch(K kernel, const P& params, int blocks, int threads, size_t smem, cudaStream_t st,
                       const char* name) {
  GSN_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int dev = 0, sms = 0, per_sm = 0;
  GSN_CUDA(cudaGetDevice(&dev));
  GSN_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  GSN_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem));
  if ((long long)per_sm * sms < blocks)
    return fail(GSN_ENOSUP, "%s: %d CTAs cannot be co-resident (%d per SM x %d SMs); reduce the batch", name,
                blocks, per_sm, sms);
  void* args[] = {const_cast<P*>(&params)};
  GSN_CUDA(cudaLaunchCooperativeKernel(reinterpret_cast<void*>(kernel), dim3(blocks), dim3(threads), args, smem, st));
  return GSN_OK;
}

}  // namespace gsn

namespace gsn {
bool recurrence_tc_train_supported(int R, int H, int shared);
size_t recurrence_tc_train_workspace(int R, int H);
int launch_recurrence_tc_train(const float*, const float*, const float*, const float*, const float*, float*, float*,
                               float*, float*, float*, float*, float*, float*, int, int, int, int, float, float, int,
                               void*, cudaStream_t);
}  // namespace gsn

extern "C" int gsn_layer_train_tc_supported(int R, int H, int shared) {
  return gsn::recurrence_tc_train_supported(R, H, shared) ? 1 : 0;
}

extern "C" size_t gsn_layer_train_tc_workspace_bytes(int R, int H) {
  return R > 0 && H > 0 ? gsn::recurrence_tc_train_workspace(R, H) : 0;
}

extern "C" int gsn_layer_train_forward_tc(const float* xproj, const float* w_hh, const float* bias,
                                          const float* bn_weight, const float* bn_bias, float* running_mean,
                                          float* running_var, float* h_out, float* c_out, float* f_out, float* g_out,
                                          float* xhat_out, float* invstd_out, int T, int R, int H, int training,
                                          float momentum, float eps, int sm_budget, void* workspace,
                                          gsn_stream_t stream) {
  GSN_REQUIRE(xproj && w_hh && bias && h_out && c_out && f_out && g_out && workspace,
              "gsn_layer_train_forward_tc: null pointer");
  GSN_REQUIRE(T > 0 && R > 0 && H > 0, "gsn_layer_train_forward_tc: bad shape T=%d R=%d H=%d", T, R, H);
  GSN_REQUIRE((bn_weight == nullptr) == (bn_bias == nullptr), "gsn_layer_train_forward_tc: bn params");
  GSN_REQUIRE(!bn_weight || (running_mean && running_var), "gsn_layer_train_forward_tc: running statistics missing");
  GSN_REQUIRE(!(bn_weight && training) || R > 1, "Expected more than 1 value per channel when training (rows=%d)", R);
  GSN_REQUIRE(!(bn_weight && training) || (xhat_out && invstd_out),
              "gsn_layer_train_forward_tc: training needs the saved-tensor outputs");
  GSN_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "gsn_layer_train_forward_tc: workspace alignment");
  if (!gsn::recurrence_tc_train_supported(R, H, 1))
    return gsn::fail(GSN_ENOSUP, "gsn_layer_train_forward_tc: shape R=%d H=%d not supported", R, H);
  return gsn::launch_recurrence_tc_train(xproj, w_hh, bias, bn_weight, bn_bias, running_mean, running_var, h_out,
                                         c_out, f_out, g_out, xhat_out, invstd_out, T, R, H, training, momentum, eps,
                                         sm_budget, workspace, gsn::as_stream(stream));
}

// workspace layout (floats): [wt: H*gH] [partial: 2*nblocks*2*H] [counter: 64 bytes]
extern "C" size_t gsn_layer_train_workspace_bytes(int R, int H, int shared) {
  if (R <= 0 || H <= 0) return 0;
  const size_t gH = shared ? H : 2 * (size_t)H;
  const size_t nblocks = (R + gsn::TR_RT - 1) / gsn::TR_RT;
  return ((size_t)H * gH + 2 * nblocks * 2 * H + 2 * nblocks * 2 * (size_t)H) * sizeof(float) + 256;
}

extern "C" int gsn_layer_train_forward(const float* xproj, const float* w_hh, const float* bias, const float* bn_weight,
                                       const float* bn_bias, float* running_mean, float* running_var, float* h_out,
                                       float* c_out, float* f_out, float* g_out, float* xhat_out, float* invstd_out,
                                       int T, int R, int H, int shared, int training, float momentum, float eps,
                                       void* workspace, gsn_stream_t stream) {
  GSN_REQUIRE(xproj && w_hh && bias && h_out && c_out && workspace, "gsn_layer_train_forward: null pointer");
  GSN_REQUIRE(T > 0 && R > 0 && H > 0 && H <= 512, "gsn_layer_train_forward: bad shape T=%d R=%d H=%d", T, R, H);
  GSN_REQUIRE((bn_weight == nullptr) == (bn_bias == nullptr), "gsn_layer_train_forward: bn params");
  GSN_REQUIRE(!bn_weight || (running_mean && running_var), "gsn_layer_train_forward: running statistics missing");
  GSN_REQUIRE(!(bn_weight && training) || R > 1,
              "Expected more than 1 value per channel when training (rows=%d)", R);
  GSN_REQUIRE(!(bn_weight && training) || (f_out && g_out && xhat_out && invstd_out),
              "gsn_layer_train_forward: training needs the saved-tensor outputs");
  GSN_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "gsn_layer_train_forward: workspace alignment");
  cudaStream_t st = gsn::as_stream(stream);
  const int gH = shared ? H : 2 * H;
  const int nblocks = (R + gsn::TR_RT - 1) / gsn::TR_RT;
  float* wt = reinterpret_cast<float*>(workspace);
  float* partial = wt + (size_t)H * gH;
  unsigned int* counter = reinterpret_cast<unsigned int*>(partial + (size_t)4 * nblocks * H + (size_t)4 * nblocks * H);
  GSN_CUDA(cudaMemsetAsync(counter, 0, 64, st));
  dim3 tb(32, 8), tg((H + 31) / 32, (gH + 31) / 32);
  gsn::k_transpose_f32<<<tg, tb, 0, st>>>(w_hh, wt, gH, H);  // [gH,H] -> [H,gH]
  GSN_LAUNCH_CHECK("k_transpose_f32");
  GSN_REQUIRE(H % 4 == 0, "gsn_layer_train_forward: H=%d must be a multiple of 4", H);
  const size_t spikes = (size_t)2 * H * gsn::TR_RT * sizeof(float);
  const size_t all_w = (size_t)H * gH * sizeof(float);
  const int resident = spikes + all_w <= 200 * 1024;
  const size_t smem = spikes + (resident ? all_w : (size_t)gsn::TR_NST * gsn::TR_KB * gH * sizeof(float));
  gsn::TrainFwdParams p{xproj, wt, bias, bn_weight, bn_bias, running_mean, running_var, h_out, c_out, f_out, g_out,
                        xhat_out, invstd_out, partial, counter, T, R, H, training, momentum, eps, resident,
                        getenv("GSN_TRAIN_PROF") ? reinterpret_cast<long long*>(counter) + 2 : nullptr};
  const int threads = ((H + 31) / 32) * 32;
  return shared ? gsn::coop_launch(gsn::k_rec_train_fwd<true>, p, nblocks, threads, smem, st, "gsn_layer_train_forward")
                : gsn::coop_launch(gsn::k_rec_train_fwd<false>, p, nblocks, threads, smem, st, "gsn_layer_train_forward");
}

extern "C" int gsn_layer_train_backward(const float* dh_out, const float* w_hh, const float* c, const float* f,
                                        const float* g, const float* xhat, const float* invstd, const float* bn_weight,
                                        const float* running_var, float* dz, float* dbias_part, float* dgamma,
                                        float* dbeta, int T, int R, int H, int shared, int training, float eps,
                                        void* workspace, gsn_stream_t stream) {
  GSN_REQUIRE(dh_out && w_hh && c && f && g && dz && dbias_part && workspace, "gsn_layer_train_backward: null pointer");
  GSN_REQUIRE(T > 0 && R > 0 && H > 0 && H <= 512, "gsn_layer_train_backward: bad shape T=%d R=%d H=%d", T, R, H);
  GSN_REQUIRE(!(bn_weight && training) || (xhat && invstd && dgamma && dbeta),
              "gsn_layer_train_backward: batch-statistics tensors missing");
  GSN_REQUIRE(!(bn_weight && !training) || running_var, "gsn_layer_train_backward: running_var missing");
  GSN_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "gsn_layer_train_backward: workspace alignment");
  cudaStream_t st = gsn::as_stream(stream);
  const int gH = shared ? H : 2 * H;
  const int nb8 = (R + gsn::TR_RT - 1) / gsn::TR_RT;   // workspace / dbias_part are sized for 8-row CTAs
  // 16-row CTAs only where the all-to-all partial traffic clearly dominates (measured: no gain at R = 768)
  const int rt = (H <= 256 && nb8 > 128) ? 16 : 8;
  const int nblocks = (R + rt - 1) / rt;
  float* wt = reinterpret_cast<float*>(workspace);
  float* partial = wt + (size_t)H * gH + (size_t)4 * nb8 * H;  // second scratch region
  unsigned int* counter = reinterpret_cast<unsigned int*>(wt + (size_t)H * gH + (size_t)8 * nb8 * H);
  GSN_CUDA(cudaMemsetAsync(counter, 0, 64, st));
  GSN_CUDA(cudaMemsetAsync(dbias_part, 0, (size_t)nb8 * 2 * H * sizeof(float), st));  // rows >= nblocks stay zero
  GSN_REQUIRE(H % 4 == 0, "gsn_layer_train_backward: H=%d must be a multiple of 4", H);
  const size_t dzs = (size_t)2 * gH * rt * sizeof(float);
  const size_t all_w = (size_t)H * gH * sizeof(float);
  const int resident = dzs + all_w <= 200 * 1024;
  const size_t smem = dzs + (resident ? all_w : (size_t)gsn::TR_NST * gsn::TR_KB * H * sizeof(float));
  gsn::TrainBwdParams p{dh_out, w_hh, c, f, g, xhat, invstd, bn_weight, running_var, dz, dbias_part, dgamma, dbeta,
                        partial, counter, T, R, H, training, eps, resident};
  const int threads = ((H + 31) / 32) * 32;
  const char* nm = "gsn_layer_train_backward";
  if (rt == 16)
    return shared ? gsn::coop_launch(gsn::k_rec_train_bwd<true, 16>, p, nblocks, threads, smem, st, nm)
                  : gsn::coop_launch(gsn::k_rec_train_bwd<false, 16>, p, nblocks, threads, smem, st, nm);
  return shared ? gsn::coop_launch(gsn::k_rec_train_bwd<true, 8>, p, nblocks, threads, smem, st, nm)
                : gsn::coop_launch(gsn::k_rec_train_bwd<false, 8>, p, nblocks, threads, smem, st, nm);
}
