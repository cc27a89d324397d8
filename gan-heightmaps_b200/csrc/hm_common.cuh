// Shared helpers for the hmgan kernel library (sm_100a only).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/hmgan.h"

namespace hm {

void set_error(const char* fmt, ...);

#define HM_CHECK_ARG(cond, ...)            \
  do {                                     \
    if (!(cond)) {                         \
      hm::set_error(__VA_ARGS__);          \
      return HM_ERR_BAD_ARG;               \
    }                                      \
  } while (0)

#define HM_CHECK_LAUNCH(name)                                                  \
  do {                                                                         \
    cudaError_t e__ = cudaGetLastError();                                      \
    if (e__ != cudaSuccess) {                                                  \
      hm::set_error("%s: launch failed: %s", name, cudaGetErrorString(e__));   \
      return HM_ERR_CUDA;                                                      \
    }                                                                          \
  } while (0)

template <typename T> __device__ __forceinline__ float ldf(const T* p);
template <> __device__ __forceinline__ float ldf<float>(const float* p) { return *p; }
template <> __device__ __forceinline__ float ldf<__half>(const __half* p) { return __half2float(*p); }
template <typename T> __device__ __forceinline__ void stf(T* p, float v);
template <> __device__ __forceinline__ void stf<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void stf<__half>(__half* p, float v) { *p = __float2half_rn(v); }

__device__ __forceinline__ float act_fwd(float v, int act, float slope) {
  switch (act) {
    case HM_ACT_LRELU: return v >= 0.f ? v : v * slope;
    case HM_ACT_RELU: return v > 0.f ? v : 0.f;
    case HM_ACT_SIGMOID: return 1.f / (1.f + expf(-v));
    case HM_ACT_TANH: return tanhf(v);
    default: return v;
  }
}
// derivative expressed through the OUTPUT value y = act(x)
__device__ __forceinline__ float act_grad_from_out(float y, int act, float slope) {
  switch (act) {
    case HM_ACT_LRELU: return y >= 0.f ? 1.f : slope;
    case HM_ACT_RELU: return y > 0.f ? 1.f : 0.f;
    case HM_ACT_SIGMOID: return y * (1.f - y);
    case HM_ACT_TANH: return 1.f - y * y;
    default: return 1.f;
  }
}

// 8 consecutive elements as one (fp16) or two (fp32) 16-byte accesses; pointers must be 16-byte aligned
__device__ __forceinline__ void load8(const __half* p, float* v) {
  uint4 u = *reinterpret_cast<const uint4*>(p);
  const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
  for (int i = 0; i < 4; i++) {
    float2 f = __half22float2(h[i]);
    v[2 * i] = f.x;
    v[2 * i + 1] = f.y;
  }
}
__device__ __forceinline__ void load8(const float* p, float* v) {
  float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void store8(__half* p, const float* v) {
  uint4 u;
  __half2* h = reinterpret_cast<__half2*>(&u);
#pragma unroll
  for (int i = 0; i < 4; i++) h[i] = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
  *reinterpret_cast<uint4*>(p) = u;
}
__device__ __forceinline__ void store8(float* p, const float* v) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// thin_conv.cu: HBM/L1-bound specialisations behind hm_conv_gather / hm_conv_wgrad; return false if the
// descriptor does not qualify (the caller then runs the generic tiled kernel).
bool thin_in_conv_launch(const HmConvDesc* d, const void* x1, const void* x2, const void* w, const float* bias,
                         void* y, void* y2, cudaStream_t st);
bool thin_wgrad_launch(const HmConvDesc* d, const void* x1, const void* x2, const void* dy, float* dw,
                       cudaStream_t st);

inline int num_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

}  // namespace hm
