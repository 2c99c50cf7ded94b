// Shared helpers for libgsn_b200 (sm_100a).  Internal header -- the public ABI is include/gsn_b200.h.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "gsn_b200.h"

namespace gsn {

// thread-local last-error message (gsn_last_error)
char* err_buf();
int fail(int code, const char* fmt, ...);

#define GSN_REQUIRE(cond, ...)                          \
  do {                                                  \
    if (!(cond)) return gsn::fail(GSN_EINVAL, __VA_ARGS__); \
  } while (0)

#define GSN_CUDA(call)                                                                      \
  do {                                                                                      \
    cudaError_t e__ = (call);                                                               \
    if (e__ != cudaSuccess)                                                                 \
      return gsn::fail(GSN_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__),  \
                       __FILE__, __LINE__);                                                 \
  } while (0)

#define GSN_LAUNCH_CHECK(name)                                                               \
  do {                                                                                       \
    cudaError_t e__ = cudaGetLastError();                                                    \
    if (e__ != cudaSuccess)                                                                  \
      return gsn::fail(GSN_ECUDA, "launch of %s failed: %s", name, cudaGetErrorString(e__)); \
  } while (0)

static inline cudaStream_t as_stream(gsn_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

// ---- optional device-side launch trace (gsn_trace_set): CTA 0 of a kernel stamps %globaltimer at entry/exit
struct TraceRec {
  unsigned long long t0, t1;
  int kind, a, b, c;
};
struct TraceBuf {  // lives in device memory; header then records
  unsigned int count, capacity;
  unsigned int pad[14];
  TraceRec rec[1];
};
TraceBuf* trace_buffer();  // host side: current buffer or nullptr (gsn_api.cu)
int launch_option(int option);  // gsn_set_option value of the calling thread (gsn_api.cu)

__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ int trace_begin(TraceBuf* tb, int kind, int a, int b, int c) {
  if (tb == nullptr || blockIdx.x != 0 || blockIdx.y != 0 || blockIdx.z != 0 || threadIdx.x != 0 || threadIdx.y != 0)
    return -1;
  const unsigned int slot = atomicAdd(&tb->count, 1u);
  if (slot >= tb->capacity) return -1;
  tb->rec[slot].kind = kind; tb->rec[slot].a = a; tb->rec[slot].b = b; tb->rec[slot].c = c;
  tb->rec[slot].t1 = 0;
  tb->rec[slot].t0 = global_ns();
  return (int)slot;
}
// every CTA of a kernel (not only CTA 0) when the capacity field's top bit is set (gsn_trace_set flags: per-CTA)
__device__ __forceinline__ int trace_begin_cta(TraceBuf* tb, int kind, int a, int b, int c) {
  if (tb == nullptr || threadIdx.x != 0 || threadIdx.y != 0 || blockIdx.x == 0) return -1;
  if (!(tb->pad[13] & 1u)) return -1;
  const unsigned int slot = atomicAdd(&tb->count, 1u);
  if (slot >= tb->capacity) return -1;
  tb->rec[slot].kind = kind; tb->rec[slot].a = a; tb->rec[slot].b = b; tb->rec[slot].c = c;
  tb->rec[slot].t1 = 0;
  tb->rec[slot].t0 = global_ns();
  return (int)slot;
}
__device__ __forceinline__ void trace_end(TraceBuf* tb, int slot) {
  if (slot >= 0) tb->rec[slot].t1 = global_ns();
}

__device__ __forceinline__ float sigmoid_f32(float x) {
  // the reference's torch.sigmoid formula 1 / (1 + exp(-x)) with full-precision expf; the division is
  // MUFU.RCP + one Newton step (<= 1 ulp, branch-free so several evaluations interleave).  The clamp keeps
  // 1 + exp(-x) finite; sigmoid(-87) ~ 1.6e-38 is already below fp32's normal range.
  const float d = 1.0f + expf(-fmaxf(x, -87.0f));
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(d));
  return fmaf(r, fmaf(-d, r, 1.0f), r);
}

// c_t and h_t of one neuron from the gate pre-activations (ESN:146-151)
__device__ __forceinline__ float gsu_membrane(float f_hat, float g_hat, float c_prev, float bn_scale,
                                              float bn_shift) {
  const float f = sigmoid_f32(f_hat);
  // written exactly as the reference evaluates it: f*c + (1-f)*g, then the folded BN affine
  float c = __fadd_rn(__fmul_rn(f, c_prev), __fmul_rn(__fsub_rn(1.0f, f), g_hat));
  c = __fadd_rn(__fmul_rn(c, bn_scale), bn_shift);
  return c;
}

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

}  // namespace gsn
