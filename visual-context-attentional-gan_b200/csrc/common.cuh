// Shared helpers for the vcagan_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdint>
#include <cstdio>

#define VCA_OK 0
#define VCA_ERR_ARG (-1)
#define VCA_ERR_CUDA (-2)
#define VCA_ERR_UNSUPPORTED (-3)

// dtype codes used across the C ABI
#define VCA_F32 0
#define VCA_BF16 1

extern "C" void vca_set_error(const char* fmt, ...);

#define VCA_CHECK_ARG(cond)                                                        \
  do {                                                                             \
    if (!(cond)) {                                                                 \
      vca_set_error("%s:%d: bad argument: %s", __FILE__, __LINE__, #cond);         \
      return VCA_ERR_ARG;                                                          \
    }                                                                              \
  } while (0)

#define VCA_LAUNCH_CHECK()                                                         \
  do {                                                                             \
    cudaError_t e__ = cudaGetLastError();                                          \
    if (e__ != cudaSuccess) {                                                      \
      vca_set_error("%s:%d: CUDA launch failed: %s", __FILE__, __LINE__,           \
                    cudaGetErrorString(e__));                                      \
      return VCA_ERR_CUDA;                                                         \
    }                                                                              \
  } while (0)

typedef __nv_bfloat16 bf16;

__device__ __forceinline__ float to_f(float v) { return v; }
__device__ __forceinline__ float to_f(bf16 v) { return __bfloat162float(v); }
template <class T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f<bf16>(float v) { return __float2bfloat16_rn(v); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Block-wide sum; all threads get the result. blockDim.x multiple of 32, <= 1024.
__device__ __forceinline__ float block_sum(float v, float* sh /*>=33 floats*/) {
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) sh[w] = v;
  __syncthreads();
  int nw = (blockDim.x + 31) >> 5;
  float r = (threadIdx.x < nw) ? sh[threadIdx.x] : 0.f;
  if (w == 0) {
    r = warp_sum(r);
    if (lane == 0) sh[32] = r;
  }
  __syncthreads();
  return sh[32];
}

static inline int vca_num_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

static inline unsigned vca_grid_1d(long long n, int block, int per_thread = 1) {
  long long g = (n + (long long)block * per_thread - 1) / ((long long)block * per_thread);
  long long cap = (long long)vca_num_sms() * 32;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (unsigned)g;
}

// Geometry of a (up to) 3-D convolution on channels-last tensors [N, D, H, W, C].
struct ConvGeom {
  int N, ID, IH, IW, Cin;
  int OD, OH, OW, Cout;
  int KD, KH, KW;
  int sd, sh, sw;
  int pd, ph, pw;
};
