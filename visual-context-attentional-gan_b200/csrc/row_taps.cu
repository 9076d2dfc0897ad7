// Row-tap combine for convolutions over an input that is CONSTANT along the H axis (the phoneme features tiled 20x
// along the mel-frequency axis, generator.py:249-250; SURVEY appendix A #14 i).  For such channels a KH x KW conv is
//     y[f] = sum_{kh : 0 <= f+kh-ph < F}  R[kh],        R[kh] = 1-D conv of the single distinct row with w[:, :, kh, :]
// so the 2-D conv collapses to ONE row (R, computed by the ordinary conv kernels with kh folded into the output
// channels: 20x fewer MACs for F = 20) plus this HBM-bound combine; only the zero padding makes rows differ.
//   fwd: y[b,f,t,c] = yn[b,f,t,c] + sum_{valid kh} R[b,t,kh*C + c]
//   bwd: dR[b,t,kh*C + c] = sum_{f : 0 <= f+kh-ph < F} dy[b,f,t,c]            (d yn = dy)
#include "vec.cuh"

namespace {

template <class T, class VT>
__global__ void row_taps_fwd_kernel(const T* __restrict__ yn, const T* __restrict__ R, T* __restrict__ y, int B, int F, int Tn, int C,
                                    int KH, int ph) {
  constexpr int V = VT::N;
  const int CV = C / V;
  const long long total = (long long)B * F * Tn * CV;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int cv = (int)(i % CV); long long r = i / CV;
    const int t = (int)(r % Tn); r /= Tn;
    const int f = (int)(r % F); const int b = (int)(r / F);
    float acc[V];
    VT::load(yn + i * V, acc);
    const T* rrow = R + ((long long)b * Tn + t) * (KH * C) + cv * V;
    for (int kh = 0; kh < KH; ++kh) {
      if ((unsigned)(f + kh - ph) >= (unsigned)F) continue;
      float v[V];
      VT::load(rrow + (long long)kh * C, v);
#pragma unroll
      for (int k = 0; k < V; ++k) acc[k] += v[k];
    }
    VT::store(y + i * V, acc);
  }
}

template <class T, class VT>
__global__ void row_taps_bwd_kernel(const T* __restrict__ dy, T* __restrict__ dR, int B, int F, int Tn, int C, int KH, int ph) {
  constexpr int V = VT::N;
  const int CV = C / V;
  const long long total = (long long)B * Tn * KH * CV;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int cv = (int)(i % CV); long long r = i / CV;
    const int kh = (int)(r % KH); r /= KH;
    const int t = (int)(r % Tn); const int b = (int)(r / Tn);
    float acc[V];
#pragma unroll
    for (int k = 0; k < V; ++k) acc[k] = 0.f;
    int f0 = ph - kh; if (f0 < 0) f0 = 0;
    int f1 = F + ph - kh; if (f1 > F) f1 = F;
    for (int f = f0; f < f1; ++f) {
      float v[V];
      VT::load(dy + ((((long long)b * F + f) * Tn + t) * C) + cv * V, v);
#pragma unroll
      for (int k = 0; k < V; ++k) acc[k] += v[k];
    }
    VT::store(dR + i * V, acc);
  }
}

}  // namespace

extern "C" {

int vca_row_taps_fwd(int dtype, const void* yn, const void* R, void* y, int B, int F, int T, int C, int KH, int ph, cudaStream_t s) {
  VCA_CHECK_ARG(yn && R && y && B > 0 && F > 0 && T > 0 && C > 0 && KH > 0 && ph >= 0);
  const int V = dtype == VCA_F32 ? 4 : 8;
  if (C % V || !vca_aligned16(yn) || !vca_aligned16(R) || !vca_aligned16(y)) {
    vca_set_error("vca_row_taps_fwd: C must be a multiple of %d and tensors 16-byte aligned", V);
    return VCA_ERR_UNSUPPORTED;
  }
  const unsigned grid = vca_grid_1d((long long)B * F * T * (C / V), 256);
  if (dtype == VCA_F32) row_taps_fwd_kernel<float, Vec<float>><<<grid, 256, 0, s>>>((const float*)yn, (const float*)R, (float*)y, B, F, T, C, KH, ph);
  else row_taps_fwd_kernel<bf16, Vec<bf16>><<<grid, 256, 0, s>>>((const bf16*)yn, (const bf16*)R, (bf16*)y, B, F, T, C, KH, ph);
  VCA_LAUNCH_CHECK();
  return VCA_OK;
}

int vca_row_taps_bwd(int dtype, const void* dy, void* dR, int B, int F, int T, int C, int KH, int ph, cudaStream_t s) {
  VCA_CHECK_ARG(dy && dR && B > 0 && F > 0 && T > 0 && C > 0 && KH > 0 && ph >= 0);
  const int V = dtype == VCA_F32 ? 4 : 8;
  if (C % V || !vca_aligned16(dy) || !vca_aligned16(dR)) {
    vca_set_error("vca_row_taps_bwd: C must be a multiple of %d and tensors 16-byte aligned", V);
    return VCA_ERR_UNSUPPORTED;
  }
  const unsigned grid = vca_grid_1d((long long)B * T * KH * (C / V), 256);
  if (dtype == VCA_F32) row_taps_bwd_kernel<float, Vec<float>><<<grid, 256, 0, s>>>((const float*)dy, (float*)dR, B, F, T, C, KH, ph);
  else row_taps_bwd_kernel<bf16, Vec<bf16>><<<grid, 256, 0, s>>>((const bf16*)dy, (bf16*)dR, B, F, T, C, KH, ph);
  VCA_LAUNCH_CHECK();
  return VCA_OK;
}

}  // extern "C"
