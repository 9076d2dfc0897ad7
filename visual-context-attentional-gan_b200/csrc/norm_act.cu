// BatchNorm (+ residual) (+ activation) on channels-last tensors viewed as a [R, C] matrix, plus the
// stand-alone activations.  HBM-bound: every kernel is a single coalesced pass (threads run along C).
// Reference sites: visual_front.py:12-13, resnet.py:34-63, generator.py:105-126,179,209-225,325-329.
#include "common.cuh"

namespace {

enum { ACT_NONE = 0, ACT_LRELU = 1, ACT_PRELU = 2, ACT_RELU = 3 };

// Thread layout for column reductions: blockDim = (CX, RY); thread x walks channels c = bx*CX + x, rows strided.
constexpr int CX = 32, RY = 8, ROWS_PER_CTA = 256;

template <class T>
__global__ void __launch_bounds__(CX* RY) bn_stats_kernel(const T* __restrict__ x, long long R, int C,
                                                          double* __restrict__ sums /*[2][C]*/) {
  __shared__ float s1[RY][CX + 1], s2[RY][CX + 1];
  const int c = blockIdx.x * CX + threadIdx.x;
  const long long r0 = (long long)blockIdx.y * ROWS_PER_CTA;
  float a = 0.f, b = 0.f;
  if (c < C) {
    long long rend = r0 + ROWS_PER_CTA < R ? r0 + ROWS_PER_CTA : R;
    for (long long r = r0 + threadIdx.y; r < rend; r += RY) {
      float v = to_f(x[r * C + c]);
      a += v; b = fmaf(v, v, b);
    }
  }
  s1[threadIdx.y][threadIdx.x] = a; s2[threadIdx.y][threadIdx.x] = b;
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
#pragma unroll
    for (int i = 1; i < RY; ++i) { a += s1[i][threadIdx.x]; b += s2[i][threadIdx.x]; }
    atomicAdd(&sums[c], (double)a);
    atomicAdd(&sums[C + c], (double)b);
  }
}

__global__ void bn_finalize_kernel(const double* __restrict__ sums, long long R, int C, float eps, float momentum,
                                   float* __restrict__ mean, float* __restrict__ invstd,
                                   float* __restrict__ running_mean, float* __restrict__ running_var) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double m = sums[c] / (double)R;
  double var = sums[C + c] / (double)R - m * m;
  if (var < 0) var = 0;
  mean[c] = (float)m;
  invstd[c] = (float)(1.0 / sqrt(var + (double)eps));
  if (running_mean) {
    double unb = R > 1 ? var * (double)R / (double)(R - 1) : var;
    running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)m;
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unb;
  }
}

__global__ void bn_eval_stats_kernel(const float* __restrict__ rm, const float* __restrict__ rv, int C, float eps,
                                     float* __restrict__ mean, float* __restrict__ invstd) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) { mean[c] = rm[c]; invstd[c] = 1.f / sqrtf(rv[c] + eps); }
}

__device__ __forceinline__ float act_fwd(float v, int act, float slope) {
  if (act == ACT_NONE) return v;
  if (act == ACT_RELU) return v > 0.f ? v : 0.f;
  return v > 0.f ? v : v * slope;  // LRELU: constant slope; PRELU: per-channel slope passed in
}

// y = act( (x-mean)*invstd*gamma + beta  [+ res] )
template <class T>
__global__ void bn_act_fwd_kernel(const T* __restrict__ x, const T* __restrict__ res, T* __restrict__ y, long long R, int C,
                                  const float* __restrict__ mean, const float* __restrict__ invstd,
                                  const float* __restrict__ gamma, const float* __restrict__ beta, int act, float slope,
                                  const float* __restrict__ prelu_w) {
  long long total = R * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % C);
    float sc = invstd[c] * gamma[c];
    float v = (to_f(x[i]) - mean[c]) * sc + beta[c];
    if (res) v += to_f(res[i]);
    float s = act == ACT_PRELU ? prelu_w[c] : slope;
    y[i] = from_f<T>(act_fwd(v, act, s));
  }
}

// per-channel sums for the backward: s[0][c] = sum dpre, s[1][c] = sum dpre*xhat, s[2][c] = sum dy*min(pre,0) (PReLU)
template <class T>
__global__ void __launch_bounds__(CX* RY)
    bn_act_bwd_reduce_kernel(const T* __restrict__ dy, const T* __restrict__ x, const T* __restrict__ res, long long R, int C,
                             const float* __restrict__ mean, const float* __restrict__ invstd,
                             const float* __restrict__ gamma, const float* __restrict__ beta, int act, float slope,
                             const float* __restrict__ prelu_w, double* __restrict__ sums /*[3][C]*/) {
  __shared__ float s1[RY][CX + 1], s2[RY][CX + 1], s3[RY][CX + 1];
  const int c = blockIdx.x * CX + threadIdx.x;
  const long long r0 = (long long)blockIdx.y * ROWS_PER_CTA;
  float a = 0.f, b = 0.f, d = 0.f;
  if (c < C) {
    const float mu = mean[c], is = invstd[c], ga = gamma[c], be = beta[c];
    const float s = act == ACT_PRELU ? prelu_w[c] : slope;
    long long rend = r0 + ROWS_PER_CTA < R ? r0 + ROWS_PER_CTA : R;
    for (long long r = r0 + threadIdx.y; r < rend; r += RY) {
      long long i = r * C + c;
      float xh = (to_f(x[i]) - mu) * is;
      float pre = xh * ga + be;
      if (res) pre += to_f(res[i]);
      float g = to_f(dy[i]);
      float dpre = g;
      if (act == ACT_RELU) dpre = pre > 0.f ? g : 0.f;
      else if (act != ACT_NONE) { dpre = pre > 0.f ? g : g * s; if (pre <= 0.f) d = fmaf(g, pre, d); }
      a += dpre; b = fmaf(dpre, xh, b);
    }
  }
  s1[threadIdx.y][threadIdx.x] = a; s2[threadIdx.y][threadIdx.x] = b; s3[threadIdx.y][threadIdx.x] = d;
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
#pragma unroll
    for (int i = 1; i < RY; ++i) { a += s1[i][threadIdx.x]; b += s2[i][threadIdx.x]; d += s3[i][threadIdx.x]; }
    atomicAdd(&sums[c], (double)a);
    atomicAdd(&sums[C + c], (double)b);
    atomicAdd(&sums[2 * C + c], (double)d);
  }
}

// dx = gamma*invstd*(dpre - mean(dpre) - xhat*mean(dpre*xhat))   (train)   |   gamma*invstd*dpre (eval)
// dres = dpre.  Also writes dgamma/dbeta/dprelu (one thread per channel in block 0).
template <class T>
__global__ void bn_act_bwd_apply_kernel(const T* __restrict__ dy, const T* __restrict__ x, const T* __restrict__ res,
                                        T* __restrict__ dx, T* __restrict__ dres, long long R, int C,
                                        const float* __restrict__ mean, const float* __restrict__ invstd,
                                        const float* __restrict__ gamma, const float* __restrict__ beta, int act,
                                        float slope, const float* __restrict__ prelu_w, const double* __restrict__ sums,
                                        int train, float* __restrict__ dgamma, float* __restrict__ dbeta,
                                        float* __restrict__ dprelu) {
  long long total = R * C;
  const double invR = 1.0 / (double)R;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % C);
    const float mu = mean[c], is = invstd[c], ga = gamma[c];
    float xh = (to_f(x[i]) - mu) * is;
    float pre = xh * ga + beta[c];
    if (res) pre += to_f(res[i]);
    float g = to_f(dy[i]);
    float s = act == ACT_PRELU ? prelu_w[c] : slope;
    float dpre = g;
    if (act == ACT_RELU) dpre = pre > 0.f ? g : 0.f;
    else if (act != ACT_NONE) dpre = pre > 0.f ? g : g * s;
    if (dres) dres[i] = from_f<T>(dpre);
    float v = dpre;
    if (train) v = dpre - (float)(sums[c] * invR) - xh * (float)(sums[C + c] * invR);
    dx[i] = from_f<T>(v * ga * is);
  }
  if (blockIdx.x == 0) {
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      if (dgamma) dgamma[c] = (float)sums[C + c];
      if (dbeta) dbeta[c] = (float)sums[c];
      if (dprelu) dprelu[c] = (float)sums[2 * C + c];
    }
  }
}

template <class T>
__global__ void lrelu_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, long long n, float slope) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float v = to_f(x[i]);
    y[i] = from_f<T>(v > 0.f ? v : v * slope);
  }
}
// dx = dy * (x > 0 ? 1 : slope)   (also used for its own double-backward: linear in dy)
template <class T>
__global__ void lrelu_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ x, T* __restrict__ dx, long long n,
                                 float slope) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float g = to_f(dy[i]);
    dx[i] = from_f<T>(to_f(x[i]) > 0.f ? g : g * slope);
  }
}
template <class T>
__global__ void tanh_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    y[i] = from_f<T>(tanhf(to_f(x[i])));
}
template <class T>
__global__ void tanh_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ y, T* __restrict__ dx, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float t = to_f(y[i]);
    dx[i] = from_f<T>(to_f(dy[i]) * (1.f - t * t));
  }
}
// out = alpha*a + beta*b  (b may be null)
template <class T>
__global__ void axpby_kernel(const T* __restrict__ a, const T* __restrict__ b, T* __restrict__ out, long long n, float alpha,
                             float beta) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float v = alpha * to_f(a[i]);
    if (b) v = fmaf(beta, to_f(b[i]), v);
    out[i] = from_f<T>(v);
  }
}
// column sums of a [R, C] matrix into fp32 (bias gradients): out[c] += sum_r x[r,c]  (out zeroed by caller)
template <class T>
__global__ void __launch_bounds__(CX* RY) colsum_kernel(const T* __restrict__ x, long long R, int C, float* __restrict__ out) {
  __shared__ float s1[RY][CX + 1];
  const int c = blockIdx.x * CX + threadIdx.x;
  const long long r0 = (long long)blockIdx.y * ROWS_PER_CTA;
  float a = 0.f;
  if (c < C) {
    long long rend = r0 + ROWS_PER_CTA < R ? r0 + ROWS_PER_CTA : R;
    for (long long r = r0 + threadIdx.y; r < rend; r += RY) a += to_f(x[r * C + c]);
  }
  s1[threadIdx.y][threadIdx.x] = a;
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
#pragma unroll
    for (int i = 1; i < RY; ++i) a += s1[i][threadIdx.x];
    atomicAdd(&out[c], a);
  }
}

template <class TI, class TO>
__global__ void cast_kernel(const TI* __restrict__ x, TO* __restrict__ y, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    y[i] = from_f<TO>(to_f(x[i]));
}

inline dim3 col_grid(long long R, int C) {
  long long gy = (R + ROWS_PER_CTA - 1) / ROWS_PER_CTA;
  return dim3((C + CX - 1) / CX, (unsigned)(gy > 65535 ? 65535 : gy), 1);
}

}  // namespace

#define DISPATCH_T(dtype, CALL_F32, CALL_BF16) \
  do { if ((dtype) == VCA_F32) { CALL_F32; } else { CALL_BF16; } } while (0)

extern "C" {

// sums: device scratch double[2*C], zeroed here.  Writes mean/invstd (biased var) and updates running stats
// (momentum, unbiased var) when running_mean != null.
int vca_bn_stats(int dtype, const void* x, long long R, int C, float eps, float momentum, double* sums, float* mean,
                 float* invstd, float* running_mean, float* running_var, cudaStream_t s) {
  VCA_CHECK_ARG(x && sums && mean && invstd && R > 0 && C > 0);
  VCA_CHECK_ARG((R + ROWS_PER_CTA - 1) / ROWS_PER_CTA <= 65535);
  cudaMemsetAsync(sums, 0, sizeof(double) * 2 * C, s);
  dim3 grid = col_grid(R, C), block(CX, RY);
  DISPATCH_T(dtype, (bn_stats_kernel<float><<<grid, block, 0, s>>>((const float*)x, R, C, sums)),
             (bn_stats_kernel<bf16><<<grid, block, 0, s>>>((const bf16*)x, R, C, sums)));
  VCA_LAUNCH_CHECK();
  bn_finalize_kernel<<<(C + 127) / 128, 128, 0, s>>>(sums, R, C, eps, momentum, mean, invstd, running_mean, running_var);
  VCA_LAUNCH_CHECK();
  return VCA_OK;
}
int vca_bn_eval_stats(const float* running_mean, const float* running_var, int C, float eps, float* mean, float* invstd,
                      cudaStream_t s) {
  VCA_CHECK_ARG(running_mean && running_var && mean && invstd && C > 0);
  bn_eval_stats_kernel<<<(C + 127) / 128, 128, 0, s>>>(running_mean, running_var, C, eps, mean, invstd);
  VCA_LAUNCH_CHECK();
  return VCA_OK;
}
int vca_bn_act_fwd(int dtype, const void* x, const void* res, void* y, long long R, int C, const float* mean,
                   const float* invstd, const float* gamma, const float* beta, int act, float slope, const float* prelu_w,
                   cudaStream_t s) {
  VCA_CHECK_ARG(x && y && mean && invstd && gamma && beta && R > 0 && C > 0 && (act != ACT_PRELU || prelu_w));
  unsigned grid = vca_grid_1d(R * C, 256, 4);
  DISPATCH_T(dtype,
             (bn_act_fwd_kernel<float><<<grid, 256, 0, s>>>((const float*)x, (const float*)res, (float*)y, R, C, mean, invstd,
                                                            gamma, beta, act, slope, prelu_w)),
             (bn_act_fwd_kernel<bf16><<<grid, 256, 0, s>>>((const bf16*)x, (const bf16*)res, (bf16*)y, R, C, mean, invstd,
                                                           gamma, beta, act, slope, prelu_w)));
  VCA_LAUNCH_CHECK();
  return VCA_OK;
}
// sums: device scratch double[3*C] (zeroed here).  dres/dgamma/dbeta/dprelu may be null.
int vca_bn_act_bwd(int dtype, const void* dy, const void* x, const void* res, void* dx, void* dres, long long R, int C,
                   const float* mean, const float* invstd, const float* gamma, const float* beta, int act, float slope,
                   const float* prelu_w, int train, double* sums, float* dgamma, float* dbeta, float* dprelu,
                   cudaStream_t s) {
  VCA_CHECK_ARG(dy && x && dx && mean && invstd && gamma && beta && sums && R > 0 && C > 0);
  VCA_CHECK_ARG((R + ROWS_PER_CTA - 1) / ROWS_PER_CTA <= 65535);
  cudaMemsetAsync(sums, 0, sizeof(double) * 3 * C, s);
  dim3 grid = col_grid(R, C), block(CX, RY);
  DISPATCH_T(dtype,
             (bn_act_bwd_reduce_kernel<float><<<grid, block, 0, s>>>((const float*)dy, (const float*)x, (const float*)res, R, C,
                                                                     mean, invstd, gamma, beta, act, slope, prelu_w, sums)),
             (bn_act_bwd_reduce_kernel<bf16><<<grid, block, 0, s>>>((const bf16*)dy, (const bf16*)x, (const bf16*)res, R, C,
                                                                    mean, invstd, gamma, beta, act, slope, prelu_w, sums)));
  VCA_LAUNCH_CHECK();
  unsigned g1 = vca_grid_1d(R * C, 256, 4);
  DISPATCH_T(dtype,
             (bn_act_bwd_apply_kernel<float><<<g1, 256, 0, s>>>((const float*)dy, (const float*)x, (const float*)res,
                                                                (float*)dx, (float*)dres, R, C, mean, invstd, gamma, beta, act,
                                                                slope, prelu_w, sums, train, dgamma, dbeta, dprelu)),
             (bn_act_bwd_apply_kernel<bf16><<<g1, 256, 0, s>>>((const bf16*)dy, (const bf16*)x, (const bf16*)res, (bf16*)dx,
                                                               (bf16*)dres, R, C, mean, invstd, gamma, beta, act, slope,
                                                               prelu_w, sums, train, dgamma, dbeta, dprelu)));
  VCA_LAUNCH_CHECK();
  return VCA_OK;
}
int vca_lrelu_fwd(int dtype, const void* x, void* y, long long n, float slope, cudaStream_t s) {
  VCA_CHECK_ARG(x && y && n >= 0);
  if (n == 0) return VCA_OK;
  unsigned grid = vca_grid_1d(n, 256, 4);
  DISPATCH_T(dtype, (lrelu_fwd_kernel<float><<<grid, 256, 0, s>>>((const float*)x, (float*)y, n, slope)),
             (lrelu_fwd_kernel<bf16><<<grid, 256, 0, s>>>((const bf16*)x, (bf16*)y, n, slope)));
  VCA_LAUNCH_CHECK();
  return VCA_OK;
}
int vca_lrelu_bwd(int dtype, const void* dy, const void* x, void* dx, long long n, float slope, cudaStream_t s) {
  VCA_CHECK_ARG(dy && x && dx && n >= 0);
  if (n == 0) return VCA_OK;
  unsigned grid = vca_grid_1d(n, 256, 4);
  DISPATCH_T(dtype, (lrelu_bwd_kernel<float><<<grid, 256, 0, s>>>((const float*)dy, (const float*)x, (float*)dx, n, slope)),
             (lrelu_bwd_kernel<bf16><<<grid, 256, 0, s>>>((const bf16*)dy, (const bf16*)x, (bf16*)dx, n, slope)));
  VCA_LAUNCH_CHECK();
  return VCA_OK;
}
int vca_tanh_fwd(int dtype, const void* x, void* y, long long n, cudaStream_t s) {
  VCA_CHECK_ARG(x && y && n >= 0);
  if (n == 0) return VCA_OK;
  unsigned grid = vca_grid_1d(n, 256, 4);
  DISPATCH_T(dtype, (tanh_fwd_kernel<float><<<grid, 256, 0, s>>>((const float*)x, (float*)y, n)),
             (tanh_fwd_kernel<bf16><<<grid, 256, 0, s>>>((const bf16*)x, (bf16*)y, n)));
  VCA_LAUNCH_CHECK();
  return VCA_OK;
}
int vca_tanh_bwd(int dtype, const void* dy, const void* y, void* dx, long long n, cudaStream_t s) {
  VCA_CHECK_ARG(dy && y && dx && n >= 0);
  if (n == 0) return VCA_OK;
  unsigned grid = vca_grid_1d(n, 256, 4);
  DISPATCH_T(dtype, (tanh_bwd_kernel<float><<<grid, 256, 0, s>>>((const float*)dy, (const float*)y, (float*)dx, n)),
             (tanh_bwd_kernel<bf16><<<grid, 256, 0, s>>>((const bf16*)dy, (const bf16*)y, (bf16*)dx, n)));
  VCA_LAUNCH_CHECK();
  return VCA_OK;
}
int vca_axpby(int dtype, const void* a, const void* b, void* out, long long n, float alpha, float beta, cudaStream_t s) {
  VCA_CHECK_ARG(a && out && n >= 0);
  if (n == 0) return VCA_OK;
  unsigned grid = vca_grid_1d(n, 256, 4);
  DISPATCH_T(dtype, (axpby_kernel<float><<<grid, 256, 0, s>>>((const float*)a, (const float*)b, (float*)out, n, alpha, beta)),
             (axpby_kernel<bf16><<<grid, 256, 0, s>>>((const bf16*)a, (const bf16*)b, (bf16*)out, n, alpha, beta)));
  VCA_LAUNCH_CHECK();
  return VCA_OK;
}
// out[c] = sum_r x[r,c]; out is zeroed here.
int vca_colsum(int dtype, const void* x, long long R, int C, float* out, cudaStream_t s) {
  VCA_CHECK_ARG(x && out && R > 0 && C > 0);
  VCA_CHECK_ARG((R + ROWS_PER_CTA - 1) / ROWS_PER_CTA <= 65535);
  cudaMemsetAsync(out, 0, sizeof(float) * C, s);
  dim3 grid = col_grid(R, C), block(CX, RY);
  DISPATCH_T(dtype, (colsum_kernel<float><<<grid, block, 0, s>>>((const float*)x, R, C, out)),
             (colsum_kernel<bf16><<<grid, block, 0, s>>>((const bf16*)x, R, C, out)));
  VCA_LAUNCH_CHECK();
  return VCA_OK;
}
int vca_cast(int dt_in, int dt_out, const void* x, void* y, long long n, cudaStream_t s) {
  VCA_CHECK_ARG(x && y && n >= 0);
  if (n == 0) return VCA_OK;
  unsigned grid = vca_grid_1d(n, 256, 4);
  if (dt_in == VCA_F32 && dt_out == VCA_BF16) cast_kernel<float, bf16><<<grid, 256, 0, s>>>((const float*)x, (bf16*)y, n);
  else if (dt_in == VCA_BF16 && dt_out == VCA_F32) cast_kernel<bf16, float><<<grid, 256, 0, s>>>((const bf16*)x, (float*)y, n);
  else if (dt_in == VCA_F32 && dt_out == VCA_F32) cast_kernel<float, float><<<grid, 256, 0, s>>>((const float*)x, (float*)y, n);
  else cast_kernel<bf16, bf16><<<grid, 256, 0, s>>>((const bf16*)x, (bf16*)y, n);
  VCA_LAUNCH_CHECK();
  return VCA_OK;
}

}  // extern "C"
