// fp32 CUDA-core recurrence (GSN_BACKEND_SIMT): the on-device fp32 arbiter and the fallback for shapes
// the tcgen05 kernel does not take.  GSULayer.forward ESN:75-81 over GSUCell.forward ESN:132-153.
//
// One CTA owns RT consecutive rows for all T frames; thread j owns neuron j of those rows, keeps the
// membrane potential c[RT] in registers across frames, and reads the recurrent weights transposed
// (w_hh_t [H, gH], coalesced over j; L1/L2 resident).  Spikes of frame t-1 sit in shared memory as
// floats [H][RT], double buffered, so one __syncthreads per frame suffices.
#include "gsn_common.cuh"

namespace gsn {

constexpr int SIMT_RT = 8;

__global__ void k_transpose_whh(const float* __restrict__ w, float* __restrict__ wt, int gH, int H) {
  // w [gH, H] -> wt [H, gH]
  __shared__ float tile[32][33];
  const int x0 = blockIdx.x * 32, y0 = blockIdx.y * 32;  // x over H (cols of w), y over gH
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = y0 + i, c = x0 + threadIdx.x;
    tile[i][threadIdx.x] = (r < gH && c < H) ? w[(size_t)r * H + c] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = x0 + i, c = y0 + threadIdx.x;  // wt row r (k), col c (out neuron)
    if (r < H && c < gH) wt[(size_t)r * gH + c] = tile[threadIdx.x][i];
  }
}

template <bool SHARED>
__global__ void __launch_bounds__(512)
    k_recurrence_simt(const float* __restrict__ xproj, const float* __restrict__ wt,
                      const float* __restrict__ bias, const float* __restrict__ bn_scale,
                      const float* __restrict__ bn_shift, const float* __restrict__ h0,
                      const float* __restrict__ c0, float* __restrict__ h_out,
                      float* __restrict__ c_out, float* __restrict__ hT, float* __restrict__ cT,
                      int T, int R, int H) {
  extern __shared__ __align__(16) float sh[];  // [2][H][RT]
  constexpr int RT = SIMT_RT;
  const int gH = SHARED ? H : 2 * H;
  const int j = threadIdx.x;
  const int row0 = blockIdx.x * RT;
  const bool active = j < H;
  const int jj = active ? j : 0;
  const float bf = bias[jj], bc = bias[H + jj];
  const float bs = bn_scale ? bn_scale[jj] : 1.0f;
  const float bt = bn_shift ? bn_shift[jj] : 0.0f;

  float c[RT];
#pragma unroll
  for (int r = 0; r < RT; ++r) {
    const int row = row0 + r;
    const bool ok = active && row < R;
    c[r] = (ok && c0) ? c0[(size_t)row * H + j] : 0.f;
    if (active) sh[(0 * H + j) * RT + r] = (ok && h0) ? h0[(size_t)row * H + j] : 0.f;
  }
  __syncthreads();

  for (int t = 0; t < T; ++t) {
    const float* cur = sh + (size_t)(t & 1) * H * RT;
    float* nxt = sh + (size_t)((t & 1) ^ 1) * H * RT;
    // issue the xproj loads early; they are consumed after the k loop
    float xf[RT], xg[RT];
#pragma unroll
    for (int r = 0; r < RT; ++r) {
      const int row = row0 + r;
      const bool ok = active && row < R;
      const size_t base = ((size_t)t * R + (ok ? row : 0)) * gH;
      xf[r] = ok ? xproj[base + j] : 0.f;
      xg[r] = SHARED ? xf[r] : (ok ? xproj[base + H + j] : 0.f);
    }
    float af[RT], ag[RT];
#pragma unroll
    for (int r = 0; r < RT; ++r) { af[r] = 0.f; ag[r] = 0.f; }
    if (active) {
#pragma unroll 4
      for (int k = 0; k < H; ++k) {
        const float wf = __ldg(wt + (size_t)k * gH + j);
        const float wg = SHARED ? 0.f : __ldg(wt + (size_t)k * gH + H + j);
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
#pragma unroll
    for (int r = 0; r < RT; ++r) {
      const int row = row0 + r;
      // the reference adds in the order (x W_ih^T + bias) + h W_hh^T  (ESN:140-145)
      const float f_hat = __fadd_rn(__fadd_rn(xf[r], bf), af[r]);
      const float g_hat = __fadd_rn(__fadd_rn(xg[r], bc), SHARED ? af[r] : ag[r]);
      const float cn = gsu_membrane(f_hat, g_hat, c[r], bs, bt);
      c[r] = cn;
      const float h = cn >= 0.f ? 1.0f : 0.0f;
      if (active) {
        nxt[j * RT + r] = h;
        if (row < R) {
          const size_t o = ((size_t)t * R + row) * H + j;
          h_out[o] = h;
          if (c_out) c_out[o] = cn;
        }
      }
    }
    __syncthreads();
  }
  if (active) {
    const float* last = sh + (size_t)(T & 1) * H * RT;
#pragma unroll
    for (int r = 0; r < RT; ++r) {
      const int row = row0 + r;
      if (row < R) {
        if (hT) hT[(size_t)row * H + j] = last[j * RT + r];
        if (cT) cT[(size_t)row * H + j] = c[r];
      }
    }
  }
}

int launch_recurrence_simt(const float* xproj, const float* w_hh, const float* bias,
                           const float* bn_scale, const float* bn_shift, const float* h0,
                           const float* c0, float* h_out, float* c_out, float* hT, float* cT, int T,
                           int R, int H, int shared, void* workspace, cudaStream_t st) {
  const int gH = shared ? H : 2 * H;
  GSN_REQUIRE(H <= 512, "gsn_layer_recurrence(SIMT): H=%d > 512", H);
  float* wt = reinterpret_cast<float*>(workspace);
  dim3 tb(32, 8), tg((H + 31) / 32, (gH + 31) / 32);
  k_transpose_whh<<<tg, tb, 0, st>>>(w_hh, wt, gH, H);
  GSN_LAUNCH_CHECK("k_transpose_whh");
  const int threads = ((H + 31) / 32) * 32;
  const int blocks = (R + SIMT_RT - 1) / SIMT_RT;
  const size_t smem = (size_t)2 * H * SIMT_RT * sizeof(float);
  if (shared) {
    GSN_CUDA(cudaFuncSetAttribute(k_recurrence_simt<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)smem));
    k_recurrence_simt<true><<<blocks, threads, smem, st>>>(xproj, wt, bias, bn_scale, bn_shift, h0, c0,
                                                            h_out, c_out, hT, cT, T, R, H);
  } else {
    GSN_CUDA(cudaFuncSetAttribute(k_recurrence_simt<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)smem));
    k_recurrence_simt<false><<<blocks, threads, smem, st>>>(xproj, wt, bias, bn_scale, bn_shift, h0,
                                                             c0, h_out, c_out, hT, cT, T, R, H);
  }
  GSN_LAUNCH_CHECK("k_recurrence_simt");
  return GSN_OK;
}

size_t recurrence_simt_workspace(int H, int shared) {
  return (size_t)H * (shared ? H : 2 * H) * sizeof(float);
}

}  // namespace gsn
