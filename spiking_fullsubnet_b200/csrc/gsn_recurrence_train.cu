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
#include "gsn_common.cuh"

namespace gsn {

constexpr int TR_RT = 8;

// sense-free monotonic grid barrier; `counter` is zeroed by the host before the launch
__device__ __forceinline__ bool grid_barrier(unsigned int* counter, unsigned int nblocks, unsigned int& epoch) {
  __syncthreads();
  bool ok = true;
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned int target = (epoch + 1u) * nblocks;
    atomicAdd(counter, 1u);
    unsigned int polls = 0;
    while (true) {
      unsigned int v;
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
      if (v >= target) break;
      if (++polls > (1u << 26)) { ok = false; break; }
    }
    __threadfence();
  }
  ok = __syncthreads_and(ok);
  ++epoch;
  return ok;
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
};

template <bool SHARED>
__global__ void __launch_bounds__(512) k_rec_train_fwd(const TrainFwdParams p) {
  extern __shared__ __align__(16) float sh[];  // [2][H][RT] spikes
  constexpr int RT = TR_RT;
  const int H = p.H, R = p.R, T = p.T;
  const int gH = SHARED ? H : 2 * H;
  const int j = threadIdx.x;
  const int row0 = blockIdx.x * RT;
  const bool active = j < H;
  const int jj = active ? j : 0;
  const unsigned int nblocks = gridDim.x;
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

  for (int t = 0; t < T; ++t) {
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
    if (active) {
#pragma unroll 4
      for (int k = 0; k < H; ++k) {
        const float wf = __ldg(p.wt + (size_t)k * gH + j);
        const float wg = SHARED ? 0.f : __ldg(p.wt + (size_t)k * gH + H + j);
        const float4 s0 = *reinterpret_cast<const float4*>(cur + k * RT);
        const float4 s1 = *reinterpret_cast<const float4*>(cur + k * RT + 4);
        const float s[RT] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
#pragma unroll
        for (int r = 0; r < RT; ++r) {
          af[r] = fmaf(s[r], wf, af[r]);
          if (!SHARED) ag[r] = fmaf(s[r], wg, ag[r]);
        }
      }
    }
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
    if (batch_stats) {
      float* part = p.partial + (size_t)(t & 1) * nblocks * 2 * H;
      if (active) {
        part[((size_t)blockIdx.x * 2 + 0) * H + j] = s1;
        part[((size_t)blockIdx.x * 2 + 1) * H + j] = s2;
      }
      if (!grid_barrier(p.counter, nblocks, epoch)) __trap();
      float a1 = 0.f, a2 = 0.f;
      if (active) {
        for (unsigned int b = 0; b < nblocks; ++b) {
          a1 += __ldcg(part + ((size_t)b * 2 + 0) * H + j);
          a2 += __ldcg(part + ((size_t)b * 2 + 1) * H + j);
        }
      }
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
  }
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
};

template <bool SHARED>
__global__ void __launch_bounds__(512) k_rec_train_bwd(const TrainBwdParams p) {
  extern __shared__ __align__(16) float sh[];  // [2][gH][RT] dz of the later frame
  constexpr int RT = TR_RT;
  const int H = p.H, R = p.R, T = p.T;
  const int gH = SHARED ? H : 2 * H;
  const int j = threadIdx.x;
  const int row0 = blockIdx.x * RT;
  const bool active = j < H;
  const int jj = active ? j : 0;
  const unsigned int nblocks = gridDim.x;
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
    float dh[RT];
#pragma unroll
    for (int r = 0; r < RT; ++r) dh[r] = 0.f;
    if (active) {
#pragma unroll 4
      for (int m = 0; m < gH; ++m) {
        const float w = __ldg(p.w_hh + (size_t)m * H + j);
        const float4 s0 = *reinterpret_cast<const float4*>(later + m * RT);
        const float4 s1 = *reinterpret_cast<const float4*>(later + m * RT + 4);
        const float s[RT] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
#pragma unroll
        for (int r = 0; r < RT; ++r) dh[r] = fmaf(s[r], w, dh[r]);
      }
    }
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
      float a1 = 0.f, a2 = 0.f;
      if (active) {
        for (unsigned int b = 0; b < nblocks; ++b) {
          a1 += __ldcg(part + ((size_t)b * 2 + 0) * H + j);
          a2 += __ldcg(part + ((size_t)b * 2 + 1) * H + j);
        }
      }
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
  gsn::TrainFwdParams p{xproj, wt, bias, bn_weight, bn_bias, running_mean, running_var, h_out, c_out, f_out, g_out,
                        xhat_out, invstd_out, partial, counter, T, R, H, training, momentum, eps};
  const int threads = ((H + 31) / 32) * 32;
  const size_t smem = (size_t)2 * H * gsn::TR_RT * sizeof(float);
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
  const int nblocks = (R + gsn::TR_RT - 1) / gsn::TR_RT;
  float* wt = reinterpret_cast<float*>(workspace);
  float* partial = wt + (size_t)H * gH + (size_t)4 * nblocks * H;  // second scratch region
  unsigned int* counter = reinterpret_cast<unsigned int*>(wt + (size_t)H * gH + (size_t)8 * nblocks * H);
  GSN_CUDA(cudaMemsetAsync(counter, 0, 64, st));
  gsn::TrainBwdParams p{dh_out, w_hh, c, f, g, xhat, invstd, bn_weight, running_var, dz, dbias_part, dgamma, dbeta,
                        partial, counter, T, R, H, training, eps};
  const int threads = ((H + 31) / 32) * 32;
  const size_t smem = (size_t)2 * gH * gsn::TR_RT * sizeof(float);
  return shared ? gsn::coop_launch(gsn::k_rec_train_bwd<true>, p, nblocks, threads, smem, st, "gsn_layer_train_backward")
                : gsn::coop_launch(gsn::k_rec_train_bwd<false>, p, nblocks, threads, smem, st, "gsn_layer_train_backward");
}
