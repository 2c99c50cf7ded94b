// C-ABI glue of libgsn_b200: error reporting, device binding, recurrence back-end dispatch.
#include <stdarg.h>

#include "gsn_common.cuh"

namespace gsn {

char* err_buf() {
  static thread_local char buf[512] = "";
  return buf;
}

static TraceBuf* g_trace = nullptr;
TraceBuf* trace_buffer() { return g_trace; }

static thread_local int g_options[4] = {0, 0, 0, 0};
int launch_option(int option) { return option >= 0 && option < 4 ? g_options[option] : 0; }

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(err_buf(), 512, fmt, ap);
  va_end(ap);
  return code;
}

// gsn_recurrence_simt.cu
int launch_recurrence_simt(const float*, const float*, const float*, const float*, const float*,
                           const float*, const float*, float*, float*, float*, float*, int, int, int,
                           int, void*, cudaStream_t);
size_t recurrence_simt_workspace(int H, int shared);
// gsn_recurrence_tc.cu
int launch_recurrence_tc(const float*, const float*, const float*, const float*, const float*,
                         const float*, const float*, float*, float*, float*, float*, int, int, int, int, int,
                         void*, uint32_t*, cudaStream_t);
size_t recurrence_tc_workspace(int R, int H, int shared);
bool recurrence_tc_supported(int R, int H, int shared);
int recurrence_tc_tile(int R, int H, int shared, int sms);
int recurrence_i8_tile(int R, int H, int shared, int sms);
// gsn_recurrence_tc_i8.cu
int launch_recurrence_i8(const float*, const float*, const float*, const float*, const float*,
                         const float*, const float*, float*, float*, float*, float*, int, int, int, int, int,
                         void*, cudaStream_t);
bool recurrence_i8_supported(int R, int H, int shared);

}  // namespace gsn

extern "C" int gsn_abi_version(void) { return GSN_ABI_VERSION; }

extern "C" const char* gsn_last_error(void) { return gsn::err_buf(); }

extern "C" int gsn_trace_set(void* device_buffer, size_t bytes) {
  if (device_buffer == nullptr) { gsn::g_trace = nullptr; return GSN_OK; }
  GSN_REQUIRE(bytes >= 64 + sizeof(gsn::TraceRec), "gsn_trace_set: buffer too small");
  unsigned int hdr[16] = {0};
  hdr[1] = (unsigned int)((bytes - 64) / sizeof(gsn::TraceRec));
  GSN_CUDA(cudaMemcpy(device_buffer, hdr, sizeof(hdr), cudaMemcpyHostToDevice));
  gsn::g_trace = reinterpret_cast<gsn::TraceBuf*>(device_buffer);
  return GSN_OK;
}

extern "C" int gsn_set_option(int option, int value) {
  GSN_REQUIRE(option == GSN_OPT_PDL || option == GSN_OPT_F32_MAX_CTAS, "gsn_set_option: unknown option %d", option);
  gsn::g_options[option] = value;
  return GSN_OK;
}

extern "C" int gsn_bind_device(int device) {
  GSN_CUDA(cudaSetDevice(device));
  return GSN_OK;
}

extern "C" int gsn_device_info(int* sm_count, int* cc_major, int* cc_minor, int* smem_optin_bytes) {
  int dev = 0;
  GSN_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp p;
  GSN_CUDA(cudaGetDeviceProperties(&p, dev));
  if (sm_count) *sm_count = p.multiProcessorCount;
  if (cc_major) *cc_major = p.major;
  if (cc_minor) *cc_minor = p.minor;
  if (smem_optin_bytes) *smem_optin_bytes = (int)p.sharedMemPerBlockOptin;
  return GSN_OK;
}

extern "C" int gsn_layer_recurrence_pick_backend(int R, int H, int shared) {
  if (gsn::recurrence_tc_supported(R, H, shared)) return GSN_BACKEND_TCGEN05;
  if (gsn::recurrence_i8_supported(R, H, shared)) return GSN_BACKEND_TCGEN05_I8;
  return GSN_BACKEND_SIMT;
}

extern "C" int gsn_layer_recurrence_tile(int R, int H, int shared, int backend, int sm_budget) {
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (sm_budget > 0 && sm_budget < sms) sms = sm_budget;
  if (backend == GSN_BACKEND_AUTO) backend = gsn_layer_recurrence_pick_backend(R, H, shared);
  if (backend == GSN_BACKEND_TCGEN05) return gsn::recurrence_tc_tile(R, H, shared, sms);
  if (backend == GSN_BACKEND_TCGEN05_I8) return gsn::recurrence_i8_tile(R, H, shared, sms);
  return 0;
}

extern "C" size_t gsn_layer_recurrence_workspace_bytes(int R, int H, int shared, int backend) {
  if (R <= 0 || H <= 0) return 0;
  size_t a = gsn::recurrence_simt_workspace(H, shared);
  size_t b = gsn::recurrence_tc_supported(R, H, shared) ? gsn::recurrence_tc_workspace(R, H, shared) : 0;
  if (backend == GSN_BACKEND_SIMT) return a;
  if (backend == GSN_BACKEND_TCGEN05 || backend == GSN_BACKEND_TCGEN05_I8) return 256;
  return a > b ? a : b;
}

extern "C" int gsn_layer_recurrence_bits(const float* xproj, const float* w_hh, const float* bias,
                                         const float* bn_scale, const float* bn_shift, const float* h0,
                                         const float* c0, float* h_out, float* c_out, float* hT, float* cT,
                                         uint32_t* h_bits, int T, int R, int H, int shared, int backend,
                                         int sm_budget, void* workspace, gsn_stream_t stream) {
  GSN_REQUIRE(xproj && w_hh && bias && h_out && workspace, "gsn_layer_recurrence: null pointer");
  GSN_REQUIRE(T > 0 && R > 0 && H > 0, "gsn_layer_recurrence: bad shape T=%d R=%d H=%d", T, R, H);
  GSN_REQUIRE((bn_scale == nullptr) == (bn_shift == nullptr), "gsn_layer_recurrence: bn params");
  GSN_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0,
              "gsn_layer_recurrence: workspace must be 256-byte aligned");
  if (backend == GSN_BACKEND_AUTO) backend = gsn_layer_recurrence_pick_backend(R, H, shared);
  cudaStream_t st = gsn::as_stream(stream);
  int rc;
  if (backend == GSN_BACKEND_SIMT) {
    rc = gsn::launch_recurrence_simt(xproj, w_hh, bias, bn_scale, bn_shift, h0, c0, h_out, c_out, hT,
                                     cT, T, R, H, shared, workspace, st);
  } else if (backend == GSN_BACKEND_TCGEN05) {
    if (!gsn::recurrence_tc_supported(R, H, shared))
      return gsn::fail(GSN_ENOSUP, "gsn_layer_recurrence(TCGEN05): shape R=%d H=%d shared=%d not supported",
                       R, H, shared);
    return gsn::launch_recurrence_tc(xproj, w_hh, bias, bn_scale, bn_shift, h0, c0, h_out, c_out, hT,
                                     cT, T, R, H, shared, sm_budget, workspace, h_bits, st);  // packs in-kernel
  } else if (backend == GSN_BACKEND_TCGEN05_I8) {
    if (!gsn::recurrence_i8_supported(R, H, shared))
      return gsn::fail(GSN_ENOSUP, "gsn_layer_recurrence(TCGEN05_I8): shape R=%d H=%d shared=%d not supported",
                       R, H, shared);
    rc = gsn::launch_recurrence_i8(xproj, w_hh, bias, bn_scale, bn_shift, h0, c0, h_out, c_out, hT, cT, T, R, H,
                                   shared, sm_budget, workspace, st);
  } else {
    return gsn::fail(GSN_EINVAL, "gsn_layer_recurrence: unknown backend %d", backend);
  }
  if (rc == GSN_OK && h_bits) rc = gsn_pack_spikes(h_out, h_bits, (int64_t)T * R, H, stream);
  return rc;
}

extern "C" int gsn_layer_recurrence(const float* xproj, const float* w_hh, const float* bias,
                                    const float* bn_scale, const float* bn_shift, const float* h0,
                                    const float* c0, float* h_out, float* c_out, float* hT, float* cT,
                                    int T, int R, int H, int shared, int backend, int sm_budget,
                                    void* workspace, gsn_stream_t stream) {
  return gsn_layer_recurrence_bits(xproj, w_hh, bias, bn_scale, bn_shift, h0, c0, h_out, c_out, hT, cT, nullptr,
                                   T, R, H, shared, backend, sm_budget, workspace, stream);
}
