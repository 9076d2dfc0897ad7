// BatchNorm (+ residual) (+ activation) on channels-last tensors viewed as a [R, C] matrix, plus the stand-alone
// activations.  HBM-bound: every kernel is a single coalesced pass with 128-bit accesses (threads own a fixed
// group of 4/8 channels, so per-channel constants live in registers and there is no index arithmetic per
// element); per-channel reductions accumulate in fp64 (exact enough to beat the fp32 CPU reference, see
// tests/diag_grad_errors.py) and are combined with one fp64 atomic per channel per CTA.
// Reference sites: visual_front.py:12-13, resnet.py:34-63, generator.py:105-126,179,209-225,325-329.
#include "vec.cuh"

namespace {

enum { ACT_NONE = 0, ACT_LRELU = 1, ACT_PRELU = 2, ACT_RELU = 3 };
constexpr int MAX_ROW_BLOCKS = 148 * 4;

__device__ __forceinline__ float act_fwd(float v, int act, float s) {
  if (act == ACT_NONE) return v;
  if (act == ACT_RELU) return v > 0.f ? v : 0.f;
  return v > 0.f ? v : v * s;
}
__device__ __forceinline__ float act_bwd(float g, float pre, int act, float s) {
  if (act == ACT_NONE) return g;
  if (act == ACT_RELU) return pre > 0.f ? g : 0.f;
  return pre > 0.f ? g : g * s;
}

// Reduce per-thread fp64 partials over threadIdx.y and add them to out[c] (one atomic per channel per CTA).
template <int V>
__device__ __forceinline__ void block_col_reduce(const double (&acc)[V], double* sh, double* out, int c0, bool active) {
  const int tx = threadIdx.x, ty = threadIdx.y, TX = blockDim.x, TY = blockDim.y;
  __syncthreads();
#pragma unroll
  for (int i = 0; i < V; ++i) sh[(ty * TX + tx) * V + i] = acc[i];
  __syncthreads();
  if (ty == 0 && active) {
#pragma unroll
    for (int i = 0; i < V; ++i) {
      double s = 0;
      for (int y = 0; y < TY; ++y) s += sh[(y * TX + tx) * V + i];
      atomicAdd(&out[c0 + i], s);
    }
  }
}


// Row loop with 4 rows of 128-bit loads in flight per thread before any arithmetic (the kernels keep ~50 registers of
// per-channel constants, so occupancy is low and memory-level parallelism has to come from here).
#define BN_ROW_LOOP(NT, LOADS, BODY, FLUSH)                                                                              \
  {                                                                                                                \
    const long long stride__ = (long long)gridDim.x * blockDim.y;                                                 \
    long long r__ = (long long)blockIdx.x * blockDim.y + threadIdx.y;                                             \
    for (; r__ + 3 * stride__ < R; r__ += 4 * stride__) {                                                         \
      typename Vec<T>::Raw raw__[4][NT];                                                                           \
      _Pragma("unroll") for (int u__ = 0; u__ < 4; ++u__) { const long long o = (r__ + u__ * stride__) * C + cv * V; LOADS(raw__[u__]) } \
      _Pragma("unroll") for (int u__ = 0; u__ < 4; ++u__) { const long long o = (r__ + u__ * stride__) * C + cv * V; BODY(raw__[u__]) }  \
      FLUSH                                                                                                        \
    }                                                                                                              \
    for (; r__ < R; r__ += stride__) {                                                                             \
      typename Vec<T>::Raw raw1__[NT];                                                                             \
      const long long o = r__ * C + cv * V;                                                                        \
      LOADS(raw1__) BODY(raw1__) FLUSH                                                                             \
    }                                                                                                              \
  }

template <class T>
__global__ void __launch_bounds__(256) bn_stats_kernel(const T* __restrict__ x, long long R, int C, double* __restrict__ sums) {
  constexpr int V = Vec<T>::N;
  __shared__ double sh[256 * V];
  const int cv = blockIdx.y * blockDim.x + threadIdx.x;
  const bool active = cv < C / V;
  double s1[V], s2[V];
#pragma unroll
  for (int i = 0; i < V; ++i) s1[i] = s2[i] = 0;
  if (active) {
#define ST_LOADS(RAW) RAW[0] = Vec<T>::ldraw(x + o);
#define ST_BODY(RAW)                                                              \
  {                                                                               \
    float v[V];                                                                   \
    Vec<T>::unpack(RAW[0], v);                                                    \
    _Pragma("unroll") for (int i = 0; i < V; ++i) { f1[i] += v[i]; f2[i] = fmaf(v[i], v[i], f2[i]); } \
  }
#define ST_FLUSH _Pragma("unroll") for (int i = 0; i < V; ++i) { s1[i] += (double)f1[i]; s2[i] += (double)f2[i]; f1[i] = f2[i] = 0.f; }
    float f1[V], f2[V];   // fp32 partials over one 4-row group, folded into the fp64 running sums (keeps the FP64 pipe idle)
#pragma unroll
    for (int i = 0; i < V; ++i) f1[i] = f2[i] = 0.f;
    BN_ROW_LOOP(1, ST_LOADS, ST_BODY, ST_FLUSH)
  }
  block_col_reduce<V>(s1, sh, sums, cv * V, active);
  block_col_reduce<V>(s2, sh, sums + C, cv * V, active);
}

__global__ void bn_finalize_kernel(const double* __restrict__ sums, long long R, int C, float eps, float momentum,
                                   float* __restrict__ mean, float* __restrict__ invstd, float* __restrict__ running_mean,
                                   float* __restrict__ running_var) {
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

struct BnParams {
  const float *mean, *invstd, *gamma, *beta, *prelu_w;
  int act;
  float slope;
};

// y = act( (x-mean)*invstd*gamma + beta  [+ res] )
template <class T>
__global__ void __launch_bounds__(256) bn_act_fwd_kernel(const T* __restrict__ x, const T* __restrict__ res, T* __restrict__ y,
                                                         long long R, int C, BnParams p) {
  constexpr int V = Vec<T>::N;
  const int cv = blockIdx.y * blockDim.x + threadIdx.x;
  if (cv >= C / V) return;
  float mu[V], sc[V], be[V], sl[V];
#pragma unroll
  for (int i = 0; i < V; ++i) {
    const int c = cv * V + i;
    mu[i] = p.mean[c]; sc[i] = p.invstd[c] * p.gamma[c]; be[i] = p.beta[c];
    sl[i] = p.act == ACT_PRELU ? p.prelu_w[c] : p.slope;
  }
#define FW_LOADS(RAW) RAW[0] = Vec<T>::ldraw(x + o); if (res) RAW[1] = Vec<T>::ldraw(res + o);
#define FW_BODY(RAW)                                                              \
  {                                                                               \
    float v[V], rr[V];                                                            \
    Vec<T>::unpack(RAW[0], v);                                                    \
    if (res) Vec<T>::unpack(RAW[1], rr);                                          \
    _Pragma("unroll") for (int i = 0; i < V; ++i) {                               \
      float t = (v[i] - mu[i]) * sc[i] + be[i];                                   \
      if (res) t += rr[i];                                                        \
      v[i] = act_fwd(t, p.act, sl[i]);                                            \
    }                                                                             \
    Vec<T>::store(y + o, v);                                                      \
  }
  BN_ROW_LOOP(2, FW_LOADS, FW_BODY, )
}

// per-channel sums for the backward: s[0][c] = sum dpre, s[1][c] = sum dpre*xhat, s[2][c] = sum dy*min(pre,0) (PReLU)
template <class T>
__global__ void __launch_bounds__(256) bn_act_bwd_reduce_kernel(const T* __restrict__ dy, const T* __restrict__ x,
                                                                const T* __restrict__ res, long long R, int C, BnParams p,
                                                                double* __restrict__ sums) {
  constexpr int V = Vec<T>::N;
  __shared__ double sh[256 * V];
  const int cv = blockIdx.y * blockDim.x + threadIdx.x;
  const bool active = cv < C / V;
  double a[V], b[V], d[V];
#pragma unroll
  for (int i = 0; i < V; ++i) a[i] = b[i] = d[i] = 0;
  if (active) {
    float mu[V], is[V], ga[V], be[V], sl[V];
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const int c = cv * V + i;
      mu[i] = p.mean[c]; is[i] = p.invstd[c]; ga[i] = p.gamma[c]; be[i] = p.beta[c];
      sl[i] = p.act == ACT_PRELU ? p.prelu_w[c] : p.slope;
    }
#define RD_LOADS(RAW) RAW[0] = Vec<T>::ldraw(x + o); RAW[1] = Vec<T>::ldraw(dy + o); if (res) RAW[2] = Vec<T>::ldraw(res + o);
#define RD_BODY(RAW)                                                              \
  {                                                                               \
    float xv[V], g[V], rr[V];                                                     \
    Vec<T>::unpack(RAW[0], xv); Vec<T>::unpack(RAW[1], g);                        \
    if (res) Vec<T>::unpack(RAW[2], rr);                                          \
    _Pragma("unroll") for (int i = 0; i < V; ++i) {                               \
      const float xh = (xv[i] - mu[i]) * is[i];                                   \
      float pre = xh * ga[i] + be[i];                                             \
      if (res) pre += rr[i];                                                      \
      const float dpre = act_bwd(g[i], pre, p.act, sl[i]);                        \
      fa[i] += dpre;                                                              \
      fb[i] = fmaf(dpre, xh, fb[i]);                                              \
      if (p.act == ACT_PRELU && pre <= 0.f) fd[i] = fmaf(g[i], pre, fd[i]);       \
    }                                                                             \
  }
#define RD_FLUSH _Pragma("unroll") for (int i = 0; i < V; ++i) { a[i] += (double)fa[i]; b[i] += (double)fb[i]; d[i] += (double)fd[i]; fa[i] = fb[i] = fd[i] = 0.f; }
    float fa[V], fb[V], fd[V];
#pragma unroll
    for (int i = 0; i < V; ++i) fa[i] = fb[i] = fd[i] = 0.f;
    BN_ROW_LOOP(3, RD_LOADS, RD_BODY, RD_FLUSH)
  }
  block_col_reduce<V>(a, sh, sums, cv * V, active);
  block_col_reduce<V>(b, sh, sums + C, cv * V, active);
  if (p.act == ACT_PRELU) block_col_reduce<V>(d, sh, sums + 2 * C, cv * V, active);
}

// dx = gamma*invstd*(dpre - mean(dpre) - xhat*mean(dpre*xhat))   (train)   |   gamma*invstd*dpre (eval);  dres = dpre.
template <class T>
__global__ void __launch_bounds__(256) bn_act_bwd_apply_kernel(const T* __restrict__ dy, const T* __restrict__ x,
                                                               const T* __restrict__ res, T* __restrict__ dx,
                                                               T* __restrict__ dres, long long R, int C, BnParams p,
                                                               const double* __restrict__ sums, int train) {
  constexpr int V = Vec<T>::N;
  const int cv = blockIdx.y * blockDim.x + threadIdx.x;
  if (cv >= C / V) return;
  float mu[V], is[V], ga[V], be[V], sl[V], m1[V], m2[V];
  const double invR = 1.0 / (double)R;
#pragma unroll
  for (int i = 0; i < V; ++i) {
    const int c = cv * V + i;
    mu[i] = p.mean[c]; is[i] = p.invstd[c]; ga[i] = p.gamma[c]; be[i] = p.beta[c];
    sl[i] = p.act == ACT_PRELU ? p.prelu_w[c] : p.slope;
    m1[i] = train ? (float)(sums[c] * invR) : 0.f;
    m2[i] = train ? (float)(sums[C + c] * invR) : 0.f;
  }
#define AP_BODY(RAW)                                                              \
  {                                                                               \
    float xv[V], g[V], rr[V];                                                     \
    Vec<T>::unpack(RAW[0], xv); Vec<T>::unpack(RAW[1], g);                        \
    if (res) Vec<T>::unpack(RAW[2], rr);                                          \
    _Pragma("unroll") for (int i = 0; i < V; ++i) {                               \
      const float xh = (xv[i] - mu[i]) * is[i];                                   \
      float pre = xh * ga[i] + be[i];                                             \
      if (res) pre += rr[i];                                                      \
      const float dpre = act_bwd(g[i], pre, p.act, sl[i]);                        \
      g[i] = dpre;                                                                \
      xv[i] = (dpre - m1[i] - xh * m2[i]) * ga[i] * is[i];                        \
    }                                                                             \
    Vec<T>::store(dx + o, xv);                                                    \
    if (dres) Vec<T>::store(dres + o, g);                                         \
  }
  BN_ROW_LOOP(3, RD_LOADS, AP_BODY, )
}

__global__ void bn_param_grads_kernel(const double* __restrict__ sums, int C, float* __restrict__ dgamma,
                                      float* __restrict__ dbeta, float* __restrict__ dprelu) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  if (dgamma) dgamma[c] = (float)sums[C + c];
  if (dbeta) dbeta[c] = (float)sums[c];
  if (dprelu) dprelu[c] = (float)sums[2 * C + c];
}

// ---- flat element-wise kernels (8 elements per thread per iteration) ------------------------------------------
struct OpLRelu { float s; __device__ float operator()(float a, float) const { return a > 0.f ? a : a * s; } };
struct OpLReluBwd { float s; __device__ float operator()(float g, float x) const { return x > 0.f ? g : g * s; } };
struct OpTanh { __device__ float operator()(float a, float) const { return tanhf(a); } };
struct OpTanhBwd { __device__ float operator()(float g, float y) const { return g * (1.f - y * y); } };
struct OpAxpby { float al, be; __device__ float operator()(float a, float b) const { return fmaf(be, b, al * a); } };
struct OpMul { __device__ float operator()(float a, float b) const { return a * b; } };

template <class TI, class TO, class Op, bool TWO>
__global__ void __launch_bounds__(256) ew_kernel(const TI* __restrict__ a, const TI* __restrict__ b, TO* __restrict__ out,
                                                 long long n, Op op, bool vec_ok) {
  constexpr int E = 8;
  const long long nv = vec_ok ? n / E : 0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nv; i += (long long)gridDim.x * blockDim.x) {
    float x[E], y[E];
    constexpr int VI = Vec<TI>::N, VO = Vec<TO>::N;
#pragma unroll
    for (int k = 0; k < E / VI; ++k) Vec<TI>::load(a + i * E + k * VI, x + k * VI);
    if (TWO) {
#pragma unroll
      for (int k = 0; k < E / VI; ++k) Vec<TI>::load(b + i * E + k * VI, y + k * VI);
    }
#pragma unroll
    for (int k = 0; k < E; ++k) x[k] = op(x[k], TWO ? y[k] : 0.f);
#pragma unroll
    for (int k = 0; k < E / VO; ++k) Vec<TO>::store(out + i * E + k * VO, x + k * VO);
  }
  for (long long i = nv * E + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = from_f<TO>(op(to_f(a[i]), TWO ? to_f(b[i]) : 0.f));
}

template <class TI, class TO, class Op>
int ew_launch(const void* a, const void* b, void* out, long long n, Op op, cudaStream_t s) {
  if (n <= 0) return VCA_OK;
  const bool vec_ok = vca_aligned16(a) && vca_aligned16(out) && (!b || vca_aligned16(b));
  unsigned grid = vca_grid_1d(n, 256, 8);
  if (b) ew_kernel<TI, TO, Op, true><<<grid, 256, 0, s>>>((const TI*)a, (const TI*)b, (TO*)out, n, op, vec_ok);
  else ew_kernel<TI, TO, Op, false><<<grid, 256, 0, s>>>((const TI*)a, nullptr, (TO*)out, n, op, vec_ok);
  VCA_LAUNCH_CHECK();
  return VCA_OK;
}
template <class Op>
int ew_dispatch(int dtype, const void* a, const void* b, void* out, long long n, Op op, cudaStream_t s) {
  return dtype == VCA_F32 ? ew_launch<float, float, Op>(a, b, out, n, op, s) : ew_launch<bf16, bf16, Op>(a, b, out, n, op, s);
}

// column sums of a [R, C] matrix into fp32 (bias gradients); scalar path handles any C
template <class T>
__global__ void __launch_bounds__(256) colsum_vec_kernel(const T* __restrict__ x, long long R, int C, double* __restrict__ out) {
  constexpr int V = Vec<T>::N;
  __shared__ double sh[256 * V];
  const int cv = blockIdx.y * blockDim.x + threadIdx.x;
  const bool active = cv < C / V;
  double a[V];
#pragma unroll
  for (int i = 0; i < V; ++i) a[i] = 0;
  if (active) {
#pragma unroll 2
    for (long long r = (long long)blockIdx.x * blockDim.y + threadIdx.y; r < R; r += (long long)gridDim.x * blockDim.y) {
      float v[V];
      Vec<T>::load(x + r * C + cv * V, v);
#pragma unroll
      for (int i = 0; i < V; ++i) a[i] += (double)v[i];
    }
  }
  block_col_reduce<V>(a, sh, out, cv * V, active);
}
template <class T>
__global__ void colsum_scalar_kernel(const T* __restrict__ x, long long R, int C, double* __restrict__ out) {
  // one warp-row per 32 channels; rows strided over blockIdx.x * blockDim.y
  const int c = blockIdx.y * 32 + threadIdx.x;
  double a = 0;
  if (c < C)
    for (long long r = (long long)blockIdx.x * blockDim.y + threadIdx.y; r < R; r += (long long)gridDim.x * blockDim.y)
      a += (double)to_f(x[r * C + c]);
  __shared__ double sh[8][33];
  sh[threadIdx.y][threadIdx.x] = a;
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
    for (int y = 1; y < 8; ++y) a += sh[y][threadIdx.x];
    atomicAdd(&out[c], a);
  }
}
__global__ void d2f_kernel(const double* __restrict__ in, float* __restrict__ out, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (float)in[i];
}

template <class T>
bool vec_ok(const void* a, const void* b, const void* c, const void* d, const void* e, int C) {
  return C % Vec<T>::N == 0 && vca_aligned16(a) && (!b || vca_aligned16(b)) && (!c || vca_aligned16(c)) &&
         (!d || vca_aligned16(d)) && (!e || vca_aligned16(e));
}

}  // namespace

extern "C" {

// sums: device scratch double[2*C], zeroed here.  Writes mean/invstd (biased var) and updates running stats
// (momentum, unbiased var) when running_mean != null.
int vca_bn_stats(int dtype, const void* x, long long R, int C, float eps, float momentum, double* sums, float* mean,
                 float* invstd, float* running_mean, float* running_var, cudaStream_t s) {
  VCA_CHECK_ARG(x && sums && mean && invstd && R > 0 && C > 0);
  const bool ok = dtype == VCA_F32 ? vec_ok<float>(x, 0, 0, 0, 0, C) : vec_ok<bf16>(x, 0, 0, 0, 0, C);
  if (!ok) { vca_set_error("vca_bn_stats: C must be a multiple of %d and x 16-byte aligned", dtype == VCA_F32 ? 4 : 8); return VCA_ERR_UNSUPPORTED; }
  cudaMemsetAsync(sums, 0, sizeof(double) * 2 * C, s);
  if (dtype == VCA_F32) {
    RowColGrid g = row_col_grid(R, C / 4, MAX_ROW_BLOCKS);
    bn_stats_kernel<float><<<g.grid, g.block, 0, s>>>((const float*)x, R, C, sums);
  } else {
    RowColGrid g = row_col_grid(R, C / 8, MAX_ROW_BLOCKS);
    bn_stats_kernel<bf16><<<g.grid, g.block, 0, s>>>((const bf16*)x, R, C, sums);
  }
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
  const bool ok = dtype == VCA_F32 ? vec_ok<float>(x, res, y, 0, 0, C) : vec_ok<bf16>(x, res, y, 0, 0, C);
  if (!ok) { vca_set_error("vca_bn_act_fwd: C must be a multiple of %d and tensors 16-byte aligned", dtype == VCA_F32 ? 4 : 8); return VCA_ERR_UNSUPPORTED; }
  BnParams p{mean, invstd, gamma, beta, prelu_w, act, slope};
  if (dtype == VCA_F32) {
    RowColGrid g = row_col_grid(R, C / 4, 148 * 8);
    bn_act_fwd_kernel<float><<<g.grid, g.block, 0, s>>>((const float*)x, (const float*)res, (float*)y, R, C, p);
  } else {
    RowColGrid g = row_col_grid(R, C / 8, 148 * 8);
    bn_act_fwd_kernel<bf16><<<g.grid, g.block, 0, s>>>((const bf16*)x, (const bf16*)res, (bf16*)y, R, C, p);
  }
  VCA_LAUNCH_CHECK();
  return VCA_OK;
}
// sums: device scratch double[3*C] (zeroed here).  dres/dgamma/dbeta/dprelu may be null.
int vca_bn_act_bwd(int dtype, const void* dy, const void* x, const void* res, void* dx, void* dres, long long R, int C,
                   const float* mean, const float* invstd, const float* gamma, const float* beta, int act, float slope,
                   const float* prelu_w, int train, double* sums, float* dgamma, float* dbeta, float* dprelu,
                   cudaStream_t s) {
  VCA_CHECK_ARG(dy && x && dx && mean && invstd && gamma && beta && sums && R > 0 && C > 0 && (act != ACT_PRELU || prelu_w));
  const bool ok = dtype == VCA_F32 ? vec_ok<float>(dy, x, res, dx, dres, C) : vec_ok<bf16>(dy, x, res, dx, dres, C);
  if (!ok) { vca_set_error("vca_bn_act_bwd: C must be a multiple of %d and tensors 16-byte aligned", dtype == VCA_F32 ? 4 : 8); return VCA_ERR_UNSUPPORTED; }
  cudaMemsetAsync(sums, 0, sizeof(double) * 3 * C, s);
  BnParams p{mean, invstd, gamma, beta, prelu_w, act, slope};
  if (dtype == VCA_F32) {
    RowColGrid g = row_col_grid(R, C / 4, MAX_ROW_BLOCKS);
    bn_act_bwd_reduce_kernel<float><<<g.grid, g.block, 0, s>>>((const float*)dy, (const float*)x, (const float*)res, R, C, p, sums);
    VCA_LAUNCH_CHECK();
    RowColGrid g2 = row_col_grid(R, C / 4, 148 * 8);
    bn_act_bwd_apply_kernel<float><<<g2.grid, g2.block, 0, s>>>((const float*)dy, (const float*)x, (const float*)res, (float*)dx,
                                                                 (float*)dres, R, C, p, sums, train);
  } else {
    RowColGrid g = row_col_grid(R, C / 8, MAX_ROW_BLOCKS);
    bn_act_bwd_reduce_kernel<bf16><<<g.grid, g.block, 0, s>>>((const bf16*)dy, (const bf16*)x, (const bf16*)res, R, C, p, sums);
    VCA_LAUNCH_CHECK();
    RowColGrid g2 = row_col_grid(R, C / 8, 148 * 8);
    bn_act_bwd_apply_kernel<bf16><<<g2.grid, g2.block, 0, s>>>((const bf16*)dy, (const bf16*)x, (const bf16*)res, (bf16*)dx,
                                                                (bf16*)dres, R, C, p, sums, train);
  }
  VCA_LAUNCH_CHECK();
  bn_param_grads_kernel<<<(C + 127) / 128, 128, 0, s>>>(sums, C, dgamma, dbeta, act == ACT_PRELU ? dprelu : nullptr);
  VCA_LAUNCH_CHECK();
  return VCA_OK;
}
int vca_lrelu_fwd(int dtype, const void* x, void* y, long long n, float slope, cudaStream_t s) {
  VCA_CHECK_ARG(x && y && n >= 0);
  return ew_dispatch(dtype, x, nullptr, y, n, OpLRelu{slope}, s);
}
int vca_lrelu_bwd(int dtype, const void* dy, const void* x, void* dx, long long n, float slope, cudaStream_t s) {
  VCA_CHECK_ARG(dy && x && dx && n >= 0);
  return ew_dispatch(dtype, dy, x, dx, n, OpLReluBwd{slope}, s);
}
int vca_tanh_fwd(int dtype, const void* x, void* y, long long n, cudaStream_t s) {
  VCA_CHECK_ARG(x && y && n >= 0);
  return ew_dispatch(dtype, x, nullptr, y, n, OpTanh{}, s);
}
int vca_tanh_bwd(int dtype, const void* dy, const void* y, void* dx, long long n, cudaStream_t s) {
  VCA_CHECK_ARG(dy && y && dx && n >= 0);
  return ew_dispatch(dtype, dy, y, dx, n, OpTanhBwd{}, s);
}
int vca_axpby(int dtype, const void* a, const void* b, void* out, long long n, float alpha, float beta, cudaStream_t s) {
  VCA_CHECK_ARG(a && out && n >= 0);
  return ew_dispatch(dtype, a, b, out, n, OpAxpby{alpha, b ? beta : 0.f}, s);
}
int vca_mul(int dtype, const void* x, const void* m, void* y, long long n, cudaStream_t s) {
  VCA_CHECK_ARG(x && m && y && n > 0);
  return ew_dispatch(dtype, x, m, y, n, OpMul{}, s);
}
int vca_cast(int dt_in, int dt_out, const void* x, void* y, long long n, cudaStream_t s) {
  VCA_CHECK_ARG(x && y && n >= 0);
  OpAxpby id{1.f, 0.f};
  if (dt_in == VCA_F32 && dt_out == VCA_BF16) return ew_launch<float, bf16, OpAxpby>(x, nullptr, y, n, id, s);
  if (dt_in == VCA_BF16 && dt_out == VCA_F32) return ew_launch<bf16, float, OpAxpby>(x, nullptr, y, n, id, s);
  if (dt_in == VCA_F32) return ew_launch<float, float, OpAxpby>(x, nullptr, y, n, id, s);
  return ew_launch<bf16, bf16, OpAxpby>(x, nullptr, y, n, id, s);
}
// out[c] = sum_r x[r,c] (fp32).  scratch: device double[C].
int vca_colsum(int dtype, const void* x, long long R, int C, double* scratch, float* out, cudaStream_t s) {
  VCA_CHECK_ARG(x && out && scratch && R > 0 && C > 0);
  cudaMemsetAsync(scratch, 0, sizeof(double) * C, s);
  const bool ok = dtype == VCA_F32 ? vec_ok<float>(x, 0, 0, 0, 0, C) : vec_ok<bf16>(x, 0, 0, 0, 0, C);
  if (ok && dtype == VCA_F32) {
    RowColGrid g = row_col_grid(R, C / 4, MAX_ROW_BLOCKS);
    colsum_vec_kernel<float><<<g.grid, g.block, 0, s>>>((const float*)x, R, C, scratch);
  } else if (ok) {
    RowColGrid g = row_col_grid(R, C / 8, MAX_ROW_BLOCKS);
    colsum_vec_kernel<bf16><<<g.grid, g.block, 0, s>>>((const bf16*)x, R, C, scratch);
  } else {
    long long gx = (R + 7) / 8; if (gx > MAX_ROW_BLOCKS) gx = MAX_ROW_BLOCKS;
    dim3 grid((unsigned)gx, (unsigned)((C + 31) / 32)), block(32, 8);
    if (dtype == VCA_F32) colsum_scalar_kernel<float><<<grid, block, 0, s>>>((const float*)x, R, C, scratch);
    else colsum_scalar_kernel<bf16><<<grid, block, 0, s>>>((const bf16*)x, R, C, scratch);
  }
  VCA_LAUNCH_CHECK();
  d2f_kernel<<<(C + 127) / 128, 128, 0, s>>>(scratch, out, C);
  VCA_LAUNCH_CHECK();
  return VCA_OK;
}

}  // extern "C"
