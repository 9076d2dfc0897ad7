// 128-bit vector access helpers for the HBM-bound kernels: 4 x fp32 or 8 x bf16 per thread per access.
#pragma once
#include "common.cuh"

template <class T> struct Vec;
template <> struct Vec<float> {
  typedef float Elem;
  static constexpr int N = 4;
  typedef float4 Raw;
  __device__ static __forceinline__ Raw ldraw(const float* p) { return *reinterpret_cast<const float4*>(p); }
  __device__ static __forceinline__ void unpack(const Raw& t, float* v) { v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
  __device__ static __forceinline__ void load(const float* p, float* v) {
    const float4 t = *reinterpret_cast<const float4*>(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  }
  __device__ static __forceinline__ void store(float* p, const float* v) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  }
};
template <> struct Vec<bf16> {
  typedef bf16 Elem;
  static constexpr int N = 8;
  typedef uint4 Raw;
  __device__ static __forceinline__ Raw ldraw(const bf16* p) { return *reinterpret_cast<const uint4*>(p); }
  __device__ static __forceinline__ void unpack(const Raw& t, float* v) {
    const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) { v[2 * i] = __uint_as_float(w[i] << 16); v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u); }
  }
  __device__ static __forceinline__ void load(const bf16* p, float* v) {
    const uint4 t = *reinterpret_cast<const uint4*>(p);
    const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      v[2 * i] = __uint_as_float(w[i] << 16);
      v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
  }
  __device__ static __forceinline__ void store(bf16* p, const float* v) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
      w[i] = *reinterpret_cast<uint32_t*>(&h);
    }
    *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
  }
};

// 4 x bf16 per access (64-bit): halves the per-thread channel state of the register-heavy BatchNorm backward kernels
struct VecH4 {
  typedef bf16 Elem;
  static constexpr int N = 4;
  typedef uint2 Raw;
  __device__ static __forceinline__ Raw ldraw(const bf16* p) { return *reinterpret_cast<const uint2*>(p); }
  __device__ static __forceinline__ void unpack(const Raw& t, float* v) {
    v[0] = __uint_as_float(t.x << 16); v[1] = __uint_as_float(t.x & 0xffff0000u);
    v[2] = __uint_as_float(t.y << 16); v[3] = __uint_as_float(t.y & 0xffff0000u);
  }
  __device__ static __forceinline__ void load(const bf16* p, float* v) { unpack(ldraw(p), v); }
  __device__ static __forceinline__ void store(bf16* p, const float* v) {
    __nv_bfloat162 h0 = __floats2bfloat162_rn(v[0], v[1]), h1 = __floats2bfloat162_rn(v[2], v[3]);
    *reinterpret_cast<uint2*>(p) = make_uint2(*reinterpret_cast<uint32_t*>(&h0), *reinterpret_cast<uint32_t*>(&h1));
  }
};

static inline bool vca_aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// Launch shape for [R, C] row-major matrices processed as column-vectors of Vec<T>::N channels:
// blockDim = (tx, ty): tx threads cover the CV = C / N column vectors (tiled by gridDim.y when CV > tx), ty rows.
struct RowColGrid {
  dim3 grid, block;
};
static inline RowColGrid row_col_grid(long long R, int CV, int max_row_blocks) {
  int tx = 1;
  while (tx < CV && tx < 64) tx <<= 1;
  int ty = 256 / tx;
  long long gx = (R + ty - 1) / ty;
  if (gx > max_row_blocks) gx = max_row_blocks;
  if (gx < 1) gx = 1;
  RowColGrid g;
  g.block = dim3(tx, ty);
  g.grid = dim3((unsigned)gx, (unsigned)((CV + tx - 1) / tx));
  return g;
}

// scalar "vector" so the same kernel template serves channel counts that are not a multiple of 4/8 (C = 1 mels)
template <class T> struct Vec1 {
  static constexpr int N = 1;
  __device__ static __forceinline__ void load(const T* p, float* v) { v[0] = to_f(*p); }
  __device__ static __forceinline__ void store(T* p, const float* v) { *p = from_f<T>(v[0]); }
};
