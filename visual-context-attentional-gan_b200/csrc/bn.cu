// BatchNorm (+ residual) (+ activation) on channels-last tensors viewed as a [R, C] matrix, and the column sums used
// for bias gradients.  All HBM-bound: every kernel is one coalesced pass in which a thread owns a fixed group of V
// channels (so the per-channel constants live in registers and there is no index arithmetic per element) and keeps
// U rows of vector loads in flight.  What bounds them is memory-level parallelism, so
//   * the kernels are templated on activation / residual and (backward, bf16) use 4 channels per thread so that the
//     per-channel state fits in <= 85 registers and 3 CTAs of 256 threads stay resident per SM;
//   * grids are exactly SMs x resident CTAs (one wave, no tail), from the occupancy calculator;
//   * per-channel reductions are fp32 per thread over its rows in two levels (bf16 data) or fp64 (fp32 data: exact
//     enough to beat the fp32 CPU reference, see tests/diag_grad_errors.py), tree-reduced in fp64 across the CTA and
//     combined with one fp64 atomic per channel per CTA.
// Reference sites: visual_front.py:12-13, resnet.py:34-63, generator.py:105-126,179,209-225,325-329.
#include "vec.cuh"

int g_bn_vec = 4;   // bf16 channels per thread in the backward kernels (4 | 8); "bn_vec" in vca_set_option

namespace {

enum { ACT_NONE = 0, ACT_LRELU = 1, ACT_PRELU = 2, ACT_RELU = 3 };

template <class E> struct AccOf { typedef double type; };
template <> struct AccOf<bf16> { typedef float type; };

__device__ __forceinline__ float bf16_rn(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }
template <int ACT> __device__ __forceinline__ float act_fwd(float v, float s) {
  if (ACT == ACT_NONE) return v;
  if (ACT == ACT_RELU) return v > 0.f ? v : 0.f;
  return v > 0.f ? v : v * s;
}
template <int ACT> __device__ __forceinline__ float act_bwd(float g, float pre, float s) {
  if (ACT == ACT_NONE) return g;
  if (ACT == ACT_RELU) return pre > 0.f ? g : 0.f;
  return pre > 0.f ? g : g * s;
}

// Sum per-thread partials over threadIdx.y (fp64 tree in shared memory, sh = double[V][256]) and add the CTA's
// column totals to out[c] with one atomic per channel.
template <int V, class A, class O = double>
__device__ __forceinline__ void block_col_reduce(const A (&acc)[V], double* sh, O* out, int c0, bool active) {
  const int TX = blockDim.x, tid = threadIdx.y * TX + threadIdx.x;
  __syncthreads();
#pragma unroll
  for (int i = 0; i < V; ++i) sh[i * 256 + tid] = (double)acc[i];
  __syncthreads();
  for (int s = blockDim.y >> 1; s > 0; s >>= 1) {
    if ((int)threadIdx.y < s) {
#pragma unroll
      for (int i = 0; i < V; ++i) sh[i * 256 + tid] += sh[i * 256 + tid + s * TX];
    }
    __syncthreads();
  }
  if (threadIdx.y == 0 && active) {
#pragma unroll
    for (int i = 0; i < V; ++i) atomicAdd(&out[c0 + i], (O)sh[i * 256 + threadIdx.x]);
  }
}

// Row loop: U rows of NT vector loads are issued before any arithmetic; the remainder runs row by row.
#define BN_ROW_LOOP(U, NT, LOADS, BODY, FLUSH)                                                                     \
  {                                                                                                                \
    const long long stride__ = (long long)gridDim.x * blockDim.y;                                                 \
    long long r__ = (long long)blockIdx.x * blockDim.y + threadIdx.y;                                             \
    for (; r__ + (U - 1) * stride__ < R; r__ += U * stride__) {                                                   \
      typename VT::Raw raw__[U][NT];                                                                               \
      _Pragma("unroll") for (int u__ = 0; u__ < U; ++u__) { const long long o = (r__ + u__ * stride__) * C + cv * V; LOADS(raw__[u__]) } \
      _Pragma("unroll") for (int u__ = 0; u__ < U; ++u__) { const long long o = (r__ + u__ * stride__) * C + cv * V; BODY(raw__[u__]) }  \
      FLUSH                                                                                                        \
    }                                                                                                              \
    for (; r__ < R; r__ += stride__) {                                                                             \
      typename VT::Raw raw1__[NT];                                                                                 \
      const long long o = r__ * C + cv * V;                                                                        \
      LOADS(raw1__) BODY(raw1__) FLUSH                                                                             \
    }                                                                                                              \
  }

// ---- statistics ---------------------------------------------------------------------------------------------------
template <class VT>
__global__ void __launch_bounds__(256, 3) bn_stats_kernel(const typename VT::Elem* __restrict__ x, long long R, int C,
                                                          double* __restrict__ sums) {
  typedef typename AccOf<typename VT::Elem>::type Acc;
  constexpr int V = VT::N, U = 8;   // 128 B in flight per thread
  __shared__ double sh[256 * V];
  const int cv = blockIdx.y * blockDim.x + threadIdx.x;
  const bool active = cv < C / V;
  Acc s1[V], s2[V];
  float f1[V], f2[V];   // fp32 partials over one U-row group, folded into the running sums
#pragma unroll
  for (int i = 0; i < V; ++i) { s1[i] = s2[i] = 0; f1[i] = f2[i] = 0.f; }
  if (active) {
#define ST_LOADS(RAW) RAW[0] = VT::ldraw(x + o);
#define ST_BODY(RAW)                                                              \
  {                                                                               \
    float v[V];                                                                   \
    VT::unpack(RAW[0], v);                                                        \
    _Pragma("unroll") for (int i = 0; i < V; ++i) { f1[i] += v[i]; f2[i] = fmaf(v[i], v[i], f2[i]); } \
  }
#define ST_FLUSH _Pragma("unroll") for (int i = 0; i < V; ++i) { s1[i] += (Acc)f1[i]; s2[i] += (Acc)f2[i]; f1[i] = f2[i] = 0.f; }
    BN_ROW_LOOP(U, 1, ST_LOADS, ST_BODY, ST_FLUSH)
  }
  block_col_reduce<V>(s1, sh, sums, cv * V, active);
  block_col_reduce<V>(s2, sh, sums + C, cv * V, active);
}

__global__ void bn_finalize_kernel(double* __restrict__ sums, long long R, int C, float eps, float momentum,
                                   float* __restrict__ mean, float* __restrict__ invstd, float* __restrict__ running_mean,
                                   float* __restrict__ running_var, int rezero) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double m = sums[c] / (double)R;
  double var = sums[C + c] / (double)R - m * m;
  if (rezero) { sums[c] = 0; sums[C + c] = 0; }   // persistent scratch: leave it zeroed for the next call (no memset launch)
  if (var < 0) var = 0;
  mean[c] = (float)m;
  invstd[c] = (float)(1.0 / sqrt(var + (double)eps));
  if (running_mean) {
    double unb = R > 1 ? var * (double)R / (double)(R - 1) : var;
    running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)m;
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unb;
  }
}

// Statistics that arrive as `fold` groups of C accumulators each (a pixel-pair merged convolution writes its output as
// [.., pair position, C]: every channel has `fold` columns): sums = [fold][C] sums, then [fold][C] sums of squares.
__global__ void bn_finalize_fold_kernel(double* __restrict__ sums, long long R, int C, int fold, float eps, float momentum,
                                        float* __restrict__ mean, float* __restrict__ invstd, float* __restrict__ running_mean,
                                        float* __restrict__ running_var) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double s1 = 0, s2 = 0;
  for (int f = 0; f < fold; ++f) {
    s1 += sums[f * C + c]; s2 += sums[(fold + f) * C + c];
    sums[f * C + c] = 0; sums[(fold + f) * C + c] = 0;       // persistent scratch: left zeroed for the next call
  }
  double m = s1 / (double)R;
  double var = s2 / (double)R - m * m;
  if (var < 0) var = 0;
  mean[c] = (float)m;
  invstd[c] = (float)(1.0 / sqrt(var + (double)eps));
  if (running_mean) {
    double unb = R > 1 ? var * (double)R / (double)(R - 1) : var;
    running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)m;
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unb;
  }
}

__global__ void bn_fold_kernel(const float* __restrict__ rm, const float* __restrict__ rv, const float* __restrict__ gamma,
                               const float* __restrict__ beta, const float* __restrict__ bias, int C, float eps,
                               float* __restrict__ scale, float* __restrict__ shift) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float sc = gamma[c] / sqrtf(rv[c] + eps);
  scale[c] = sc;
  shift[c] = beta[c] + ((bias ? bias[c] : 0.f) - rm[c]) * sc;
}
__global__ void bn_eval_stats_kernel(const float* __restrict__ rm, const float* __restrict__ rv, int C, float eps,
                                     float* __restrict__ mean, float* __restrict__ invstd) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) { mean[c] = rm[c]; invstd[c] = 1.f / sqrtf(rv[c] + eps); }
}

struct BnParams {
  const float *mean, *invstd, *gamma, *beta, *prelu_w;
  float slope;
};

// ---- forward:  y = act( (x-mean)*invstd*gamma + beta  [+ res] ) -------------------------------------------------
template <class VT, int ACT, bool RES>
__global__ void __launch_bounds__(256, 3) bn_act_fwd_kernel(const typename VT::Elem* __restrict__ x,
                                                            const typename VT::Elem* __restrict__ res,
                                                            typename VT::Elem* __restrict__ y, long long R, int C, BnParams p) {
  constexpr int V = VT::N, U = 4, NS = ACT == ACT_PRELU ? V : 1;
  const int cv = blockIdx.y * blockDim.x + threadIdx.x;
  if (cv >= C / V) return;
  float mu[V], sc[V], be[V], sl[NS];
#pragma unroll
  for (int i = 0; i < V; ++i) {
    const int c = cv * V + i;
    mu[i] = p.mean[c]; sc[i] = p.invstd[c] * p.gamma[c]; be[i] = p.beta[c];
    if (ACT == ACT_PRELU) sl[i % NS] = p.prelu_w[c];
  }
  if (ACT != ACT_PRELU) sl[0] = p.slope;
#define FW_LOADS(RAW) RAW[0] = VT::ldraw(x + o); if (RES) RAW[1] = VT::ldraw(res + o);
#define FW_BODY(RAW)                                                              \
  {                                                                               \
    float v[V], rr[V];                                                            \
    VT::unpack(RAW[0], v);                                                        \
    if (RES) VT::unpack(RAW[1], rr);                                              \
    _Pragma("unroll") for (int i = 0; i < V; ++i) {                               \
      float t = (v[i] - mu[i]) * sc[i] + be[i];                                   \
      if (RES) t += rr[i];                                                        \
      v[i] = act_fwd<ACT>(t, sl[i % NS]);                                         \
    }                                                                             \
    VT::store(y + o, v);                                                          \
  }
  BN_ROW_LOOP(U, (RES ? 2 : 1), FW_LOADS, FW_BODY, )
}

// ---- backward, pass 1: per-channel sums  s[0][c] = sum dpre, s[1][c] = sum dpre*xhat, s[2][c] = sum dy*min(pre,0) --
template <class VT, int ACT, bool RES>
__global__ void __launch_bounds__(256, VT::N * sizeof(typename VT::Elem) == 8 ? 3 : 2)
bn_act_bwd_reduce_kernel(const typename VT::Elem* __restrict__ dy, const typename VT::Elem* __restrict__ x,
                         const typename VT::Elem* __restrict__ res, long long R, int C, BnParams p, double* __restrict__ sums) {
  typedef typename AccOf<typename VT::Elem>::type Acc;
  constexpr int V = VT::N, U = VT::N * sizeof(typename VT::Elem) == 8 ? (RES ? 6 : 8) : 4, NS = ACT == ACT_PRELU ? V : 1;
  __shared__ double sh[256 * V];
  const int cv = blockIdx.y * blockDim.x + threadIdx.x;
  const bool active = cv < C / V;
  Acc a[V], b[V], d[NS];
  float fa[V], fb[V], fd[NS];
#pragma unroll
  for (int i = 0; i < V; ++i) { a[i] = b[i] = 0; fa[i] = fb[i] = 0.f; }
#pragma unroll
  for (int i = 0; i < NS; ++i) { d[i] = 0; fd[i] = 0.f; }
  if (active) {
    float mu[V], is[V], ga[V], be[V], sl[NS];
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const int c = cv * V + i;
      mu[i] = p.mean[c]; is[i] = p.invstd[c]; ga[i] = p.gamma[c]; be[i] = p.beta[c];
      if (ACT == ACT_PRELU) sl[i % NS] = p.prelu_w[c];
    }
    if (ACT != ACT_PRELU) sl[0] = p.slope;
#define RD_LOADS(RAW) RAW[0] = VT::ldraw(x + o); RAW[1] = VT::ldraw(dy + o); if (RES) RAW[2] = VT::ldraw(res + o);
#define RD_BODY(RAW)                                                              \
  {                                                                               \
    float xv[V], g[V], rr[V];                                                     \
    VT::unpack(RAW[0], xv); VT::unpack(RAW[1], g);                                \
    if (RES) VT::unpack(RAW[2], rr);                                              \
    _Pragma("unroll") for (int i = 0; i < V; ++i) {                               \
      const float xh = (xv[i] - mu[i]) * is[i];                                   \
      float pre = xh * ga[i] + be[i];                                             \
      if (RES) pre += rr[i];                                                      \
      const float dpre = act_bwd<ACT>(g[i], pre, sl[i % NS]);                     \
      fa[i] += dpre;                                                              \
      fb[i] = fmaf(dpre, xh, fb[i]);                                              \
      if (ACT == ACT_PRELU && pre <= 0.f) fd[i % NS] = fmaf(g[i], pre, fd[i % NS]); \
    }                                                                             \
  }
#define RD_FLUSH                                                                                                   \
  _Pragma("unroll") for (int i = 0; i < V; ++i) { a[i] += (Acc)fa[i]; b[i] += (Acc)fb[i]; fa[i] = fb[i] = 0.f; }   \
  if (ACT == ACT_PRELU) { _Pragma("unroll") for (int i = 0; i < NS; ++i) { d[i] += (Acc)fd[i]; fd[i] = 0.f; } }
    BN_ROW_LOOP(U, (RES ? 3 : 2), RD_LOADS, RD_BODY, RD_FLUSH)
  }
  block_col_reduce<V>(a, sh, sums, cv * V, active);
  block_col_reduce<V>(b, sh, sums + C, cv * V, active);
  if (ACT == ACT_PRELU) block_col_reduce<NS>(d, sh, sums + 2 * C, cv * V, active);
}

// ---- backward, pass 2:  dx = gamma*invstd*(dpre - mean(dpre) - xhat*mean(dpre*xhat))  (train) | gamma*invstd*dpre
// (eval);  dres = dpre. ---------------------------------------------------------------------------------------------
template <class VT, int ACT, bool RES>
__global__ void __launch_bounds__(256, VT::N * sizeof(typename VT::Elem) == 8 ? 3 : 2)
bn_act_bwd_apply_kernel(const typename VT::Elem* __restrict__ dy, const typename VT::Elem* __restrict__ x,
                        const typename VT::Elem* __restrict__ res, typename VT::Elem* __restrict__ dx,
                        typename VT::Elem* __restrict__ dres, long long R, int C, BnParams p,
                        const double* __restrict__ sums, int train) {
  constexpr int V = VT::N, U = VT::N * sizeof(typename VT::Elem) == 8 ? (RES ? 6 : 8) : 4, NS = ACT == ACT_PRELU ? V : 1;
  const int cv = blockIdx.y * blockDim.x + threadIdx.x;
  if (cv >= C / V) return;
  float mu[V], is[V], ga[V], be[V], sl[NS], m1[V], m2[V];
  const double invR = 1.0 / (double)R;
#pragma unroll
  for (int i = 0; i < V; ++i) {
    const int c = cv * V + i;
    mu[i] = p.mean[c]; is[i] = p.invstd[c]; ga[i] = p.gamma[c]; be[i] = p.beta[c];
    if (ACT == ACT_PRELU) sl[i % NS] = p.prelu_w[c];
    m1[i] = train ? (float)(sums[c] * invR) : 0.f;
    m2[i] = train ? (float)(sums[C + c] * invR) : 0.f;
  }
  if (ACT != ACT_PRELU) sl[0] = p.slope;
#define AP_BODY(RAW)                                                              \
  {                                                                               \
    float xv[V], g[V], rr[V];                                                     \
    VT::unpack(RAW[0], xv); VT::unpack(RAW[1], g);                                \
    if (RES) VT::unpack(RAW[2], rr);                                              \
    _Pragma("unroll") for (int i = 0; i < V; ++i) {                               \
      const float xh = (xv[i] - mu[i]) * is[i];                                   \
      float pre = xh * ga[i] + be[i];                                             \
      if (RES) pre += rr[i];                                                      \
      const float dpre = act_bwd<ACT>(g[i], pre, sl[i % NS]);                     \
      g[i] = dpre;                                                                \
      xv[i] = (dpre - m1[i] - xh * m2[i]) * ga[i] * is[i];                        \
    }                                                                             \
    VT::store(dx + o, xv);                                                        \
    if (RES) VT::store(dres + o, g);                                              \
  }
  BN_ROW_LOOP(U, (RES ? 3 : 2), RD_LOADS, AP_BODY, )
}

// ---- visual front-end stem: BatchNorm3d -> PReLU -> MaxPool3d((1,3,3),(1,2,2),(0,1,1)) in one pass (visual_front.py:12-14)
// The stem's activation is the largest tensor of the step (B*T x 56 x 56 x 64 bf16 = 963 MB at B = 32, T = 75).  Unfused it
// is written by the conv, read + written by BN/PReLU, read by the pool (+ the mirror image backward).  Here the pool reads
// the RAW conv output, applies scale/shift/PReLU to each of the 9 taps in registers (rounded to bf16 first, so values and
// argmax are exactly those of the unfused path) and writes only the pooled tensors: y, the argmax code, and the raw x at
// the argmax (what the backward's reduction pass needs, at a quarter of the size).
template <class VT>
__global__ void __launch_bounds__(256) bn_prelu_maxpool_fwd_kernel(const bf16* __restrict__ x, bf16* __restrict__ y,
                                                                   unsigned char* __restrict__ idx, bf16* __restrict__ xmax, int NF,
                                                                   int H, int W, int C, int OH, int OW, BnParams p) {
  constexpr int V = 8;
  const int CV = C / V;
  const long long tid = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const int cv = (int)(tid % CV);                     // gridDim.x * blockDim.x is a multiple of CV: fixed channel group
  float mu[V], sc[V], be[V], sl[V];
#pragma unroll
  for (int i = 0; i < V; ++i) {
    const int c = cv * V + i;
    mu[i] = p.mean[c]; sc[i] = p.invstd[c] * p.gamma[c]; be[i] = p.beta[c]; sl[i] = p.prelu_w[c];
  }
  const long long total = (long long)NF * OH * OW * CV;
  for (long long i = tid; i < total; i += (long long)gridDim.x * blockDim.x) {
    unsigned r = (unsigned)(i / CV);
    const int ow = (int)(r % (unsigned)OW); r /= (unsigned)OW;
    const int oh = (int)(r % (unsigned)OH); const int n = (int)(r / (unsigned)OH);
    float best[V], xb[V]; int bi[V];
#pragma unroll
    for (int k = 0; k < V; ++k) { best[k] = -INFINITY; bi[k] = 0; xb[k] = 0.f; }
    uint4 raw[9]; bool ok[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) {                     // all taps in flight before any is used
      const int h = oh * 2 - 1 + t / 3, w = ow * 2 - 1 + t % 3;
      ok[t] = (unsigned)h < (unsigned)H && (unsigned)w < (unsigned)W;
      if (ok[t]) raw[t] = *reinterpret_cast<const uint4*>(x + (((long long)n * H + h) * W + w) * C + cv * V);
    }
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      if (!ok[t]) continue;
      float v[V];
      Vec<bf16>::unpack(raw[t], v);
#pragma unroll
      for (int k = 0; k < V; ++k) {
        float z = (v[k] - mu[k]) * sc[k] + be[k];        // the arithmetic of bn_act_fwd_kernel, so the values are bit-identical
        z = bf16_rn(z > 0.f ? z : z * sl[k]);
        if (z > best[k] || (z != z && best[k] == best[k])) { best[k] = z; bi[k] = t; xb[k] = v[k]; }
      }
    }
    const long long o = i * V;
    Vec<bf16>::store(y + o, best);
    Vec<bf16>::store(xmax + o, xb);
    unsigned lo = 0, hi = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) { lo |= (unsigned)bi[k] << (8 * k); hi |= (unsigned)bi[4 + k] << (8 * k); }
    *reinterpret_cast<uint2*>(idx + o) = make_uint2(lo, hi);
  }
}
// backward, dense pass: dx[pos] = gamma*invstd*(g' - mean(g') - xhat*mean(g'*xhat)),  g'[pos] = prelu'(z[pos]) * sum of the
// pooled gradients whose window has its argmax at pos (at most 2 x 2 windows contain a position); sums from the reduction
// over the POOLED tensors (bn_act_bwd_reduce_kernel on dy_pool / xmax), means over the R_full positions.
__global__ void __launch_bounds__(256) bn_prelu_maxpool_bwd_apply_kernel(const bf16* __restrict__ dy, const unsigned char* __restrict__ idx,
                                                                         const bf16* __restrict__ x, bf16* __restrict__ dx, int NF, int H,
                                                                         int W, int C, int OH, int OW, BnParams p,
                                                                         const double* __restrict__ sums, int train) {
  // One thread = 8 channels of a 2 x 2 block of positions  h in {2a-1, 2a}, w in {2b-1, 2b}.  Those four positions lie in
  // (at most) the four pooling windows (a-1 | a) x (b-1 | b), whose pooled gradient + argmax codes are fetched ONCE: 4 window
  // loads + 4 x loads for 4 outputs, all issued before the first use (the one-position-per-thread version gathered 4 windows
  // per output, one dependent batch at a time: 1.3 TB/s).
  constexpr int V = 8;
  const int CV = C / V;
  const long long tid = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const int cv = (int)(tid % CV);
  float mu[V], is[V], ga[V], be[V], sl[V], m1[V], m2[V];
  const double invR = 1.0 / ((double)NF * H * W);
#pragma unroll
  for (int i = 0; i < V; ++i) {
    const int c = cv * V + i;
    mu[i] = p.mean[c]; is[i] = p.invstd[c]; ga[i] = p.gamma[c]; be[i] = p.beta[c]; sl[i] = p.prelu_w[c];
    m1[i] = train ? (float)(sums[c] * invR) : 0.f;
    m2[i] = train ? (float)(sums[C + c] * invR) : 0.f;
  }
  const int AB = H / 2 + 1, BB = W / 2 + 1;
  const long long total = (long long)NF * AB * BB * CV;
  for (long long i = tid; i < total; i += (long long)gridDim.x * blockDim.x) {
    unsigned r = (unsigned)(i / CV);
    const int bb = (int)(r % (unsigned)BB); r /= (unsigned)BB;
    const int aa = (int)(r % (unsigned)AB); const int n = (int)(r / (unsigned)AB);
    const int hs[2] = {2 * aa - 1, 2 * aa}, ws[2] = {2 * bb - 1, 2 * bb};
    const bool hv[2] = {hs[0] >= 0, hs[1] < H}, wv[2] = {ws[0] >= 0, ws[1] < W};
    // windows q = 2 * i + j: (aa - 1 + i, bb - 1 + j)
    uint4 gr[4]; unsigned long long codes[4]; uint4 xr[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int oh = aa - 1 + (q >> 1), ow = bb - 1 + (q & 1);
      codes[q] = 0xffffffffffffffffull; gr[q] = make_uint4(0, 0, 0, 0);
      if (oh >= 0 && oh < OH && ow >= 0 && ow < OW) {
        const long long o = (((long long)n * OH + oh) * OW + ow) * C + cv * V;
        gr[q] = *reinterpret_cast<const uint4*>(dy + o);
        codes[q] = *reinterpret_cast<const unsigned long long*>(idx + o);
      }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      xr[q] = make_uint4(0, 0, 0, 0);
      if (hv[q >> 1] && wv[q & 1]) xr[q] = *reinterpret_cast<const uint4*>(x + (((long long)n * H + hs[q >> 1]) * W + ws[q & 1]) * C + cv * V);
    }
    float g[4][V];
#pragma unroll
    for (int q = 0; q < 4; ++q) Vec<bf16>::unpack(gr[q], g[q]);
#pragma unroll
    for (int pq = 0; pq < 4; ++pq) {                   // position (hs[pi], ws[pj])
      const int pi = pq >> 1, pj = pq & 1;
      if (!(hv[pi] && wv[pj])) continue;
      float acc[V], xv[V];
#pragma unroll
      for (int k = 0; k < V; ++k) acc[k] = 0.f;
      // an odd coordinate (index 0) lies in windows (-1 | 0) at in-window offsets (2 | 0); an even one only in window 0 at offset 1
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int qi = q >> 1, qj = q & 1;
        if ((pi == 1 && qi == 0) || (pj == 1 && qj == 0)) continue;
        const unsigned code = (unsigned)((pi == 1 ? 1 : (qi == 0 ? 2 : 0)) * 3 + (pj == 1 ? 1 : (qj == 0 ? 2 : 0)));
#pragma unroll
        for (int k = 0; k < V; ++k)
          if ((unsigned)((codes[q] >> (8 * k)) & 0xffull) == code) acc[k] += g[q][k];
      }
      Vec<bf16>::unpack(xr[pq], xv);
#pragma unroll
      for (int k = 0; k < V; ++k) {
        const float xh = (xv[k] - mu[k]) * is[k];
        const float pre = xh * ga[k] + be[k];
        const float dpre = pre > 0.f ? acc[k] : acc[k] * sl[k];
        xv[k] = (dpre - m1[k] - xh * m2[k]) * ga[k] * is[k];
      }
      Vec<bf16>::store(dx + (((long long)n * H + hs[pi]) * W + ws[pj]) * C + cv * V, xv);
    }
  }
}

__global__ void bn_param_grads_kernel(double* __restrict__ sums, int C, float* __restrict__ dgamma,
                                      float* __restrict__ dbeta, float* __restrict__ dprelu, int rezero, int accumulate) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  // accumulate: the destinations are live .grad views that other streams may be adding to concurrently -> atomics
  if (accumulate) {
    if (dgamma) atomicAdd(&dgamma[c], (float)sums[C + c]);
    if (dbeta) atomicAdd(&dbeta[c], (float)sums[c]);
    if (dprelu) atomicAdd(&dprelu[c], (float)sums[2 * C + c]);
  } else {
    if (dgamma) dgamma[c] = (float)sums[C + c];
    if (dbeta) dbeta[c] = (float)sums[c];
    if (dprelu) dprelu[c] = (float)sums[2 * C + c];
  }
  if (rezero) { sums[c] = 0; sums[C + c] = 0; sums[2 * C + c] = 0; }
}

// ---- column sums of a [R, C] matrix (bias gradients); the scalar kernel handles any C ---------------------------
template <class VT, class O = double>
__global__ void __launch_bounds__(256, 3) colsum_vec_kernel(const typename VT::Elem* __restrict__ x, long long R, int C,
                                                            O* __restrict__ out) {
  typedef typename AccOf<typename VT::Elem>::type Acc;
  constexpr int V = VT::N, U = 8;
  __shared__ double sh[256 * V];
  const int cv = blockIdx.y * blockDim.x + threadIdx.x;
  const bool active = cv < C / V;
  Acc a[V];
  float f[V];
#pragma unroll
  for (int i = 0; i < V; ++i) { a[i] = 0; f[i] = 0.f; }
  if (active) {
#define CS_BODY(RAW)                                                              \
  {                                                                               \
    float v[V];                                                                   \
    VT::unpack(RAW[0], v);                                                        \
    _Pragma("unroll") for (int i = 0; i < V; ++i) f[i] += v[i];                   \
  }
#define CS_FLUSH _Pragma("unroll") for (int i = 0; i < V; ++i) { a[i] += (Acc)f[i]; f[i] = 0.f; }
    BN_ROW_LOOP(U, 1, ST_LOADS, CS_BODY, CS_FLUSH)
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
__global__ void d2f_kernel(const double* __restrict__ in, float* __restrict__ out, int n, int accumulate) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { if (accumulate) atomicAdd(&out[i], (float)in[i]); else out[i] = (float)in[i]; }
}

// ---- host side ------------------------------------------------------------------------------------------------------
template <class VT>
bool vec_ok(const void* a, const void* b, const void* c, const void* d, const void* e, int C) {
  return C % VT::N == 0 && vca_aligned16(a) && (!b || vca_aligned16(b)) && (!c || vca_aligned16(c)) &&
         (!d || vca_aligned16(d)) && (!e || vca_aligned16(e));
}

// CTAs of 256 threads that stay resident per SM for kernel K (occupancy calculator, cached per instantiation)
template <auto K> int resident_ctas() {
  static const int n = [] {
    int v = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, K, 256, 0) != cudaSuccess || v < 1) { cudaGetLastError(); v = 1; }
    return v;
  }();
  return n;
}
// One-wave launch shape: all CTAs co-resident, rows strided over gridDim.x * blockDim.y.
template <auto K> RowColGrid one_wave_grid(long long R, int CV) {
  RowColGrid g = row_col_grid(R, CV, 1 << 30);
  long long cap = (long long)vca_num_sms() * resident_ctas<K>() / g.grid.y;
  if (cap < 1) cap = 1;
  if (g.grid.x > cap) g.grid.x = (unsigned)cap;
  return g;
}

template <class VT>
void launch_stats(const void* x, long long R, int C, double* sums, cudaStream_t s) {
  typedef typename VT::Elem T;
  RowColGrid g = one_wave_grid<bn_stats_kernel<VT>>(R, C / VT::N);
  bn_stats_kernel<VT><<<g.grid, g.block, 0, s>>>((const T*)x, R, C, sums);
}
template <class VT, int ACT, bool RES>
void launch_fwd(const void* x, const void* res, void* y, long long R, int C, const BnParams& p, cudaStream_t s) {
  typedef typename VT::Elem T;
  RowColGrid g = one_wave_grid<bn_act_fwd_kernel<VT, ACT, RES>>(R, C / VT::N);
  bn_act_fwd_kernel<VT, ACT, RES><<<g.grid, g.block, 0, s>>>((const T*)x, (const T*)res, (T*)y, R, C, p);
}
template <class VT, int ACT, bool RES>
void launch_bwd(const void* dy, const void* x, const void* res, void* dx, void* dres, long long R, int C, const BnParams& p,
                int train, double* sums, cudaStream_t s) {
  typedef typename VT::Elem T;
  RowColGrid g = one_wave_grid<bn_act_bwd_reduce_kernel<VT, ACT, RES>>(R, C / VT::N);
  bn_act_bwd_reduce_kernel<VT, ACT, RES><<<g.grid, g.block, 0, s>>>((const T*)dy, (const T*)x, (const T*)res, R, C, p, sums);
  RowColGrid g2 = one_wave_grid<bn_act_bwd_apply_kernel<VT, ACT, RES>>(R, C / VT::N);
  bn_act_bwd_apply_kernel<VT, ACT, RES><<<g2.grid, g2.block, 0, s>>>((const T*)dy, (const T*)x, (const T*)res, (T*)dx, (T*)dres,
                                                                      R, C, p, sums, train);
}

#define BN_ACT_SWITCH(FN, VT, RES, ...)                                            \
  switch (act) {                                                                   \
    case ACT_NONE: FN<VT, ACT_NONE, RES>(__VA_ARGS__); break;                      \
    case ACT_LRELU: FN<VT, ACT_LRELU, RES>(__VA_ARGS__); break;                    \
    case ACT_PRELU: FN<VT, ACT_PRELU, RES>(__VA_ARGS__); break;                    \
    default: FN<VT, ACT_RELU, RES>(__VA_ARGS__); break;                            \
  }
#define BN_DISPATCH(FN, VT, ...)                                                   \
  if (has_res) { BN_ACT_SWITCH(FN, VT, true, __VA_ARGS__) } else { BN_ACT_SWITCH(FN, VT, false, __VA_ARGS__) }

}  // namespace

extern "C" {

// sums: device scratch double[2*C]; sums_prezeroed = 0: zeroed here (memset); 1: the caller keeps a persistent scratch
// that is zero on entry and is left zeroed on exit.  Writes mean/invstd (biased var) and updates running stats
// (momentum, unbiased var) when running_mean != null.
int vca_bn_stats(int dtype, const void* x, long long R, int C, float eps, float momentum, double* sums, int sums_prezeroed,
                 float* mean, float* invstd, float* running_mean, float* running_var, cudaStream_t s) {
  VCA_CHECK_ARG(x && sums && mean && invstd && R > 0 && C > 0);
  const bool ok = dtype == VCA_F32 ? vec_ok<Vec<float>>(x, 0, 0, 0, 0, C) : vec_ok<Vec<bf16>>(x, 0, 0, 0, 0, C);
  if (!ok) { vca_set_error("vca_bn_stats: C must be a multiple of %d and x 16-byte aligned", dtype == VCA_F32 ? 4 : 8); return VCA_ERR_UNSUPPORTED; }
  if (!sums_prezeroed) cudaMemsetAsync(sums, 0, sizeof(double) * 2 * C, s);
  if (dtype == VCA_F32) launch_stats<Vec<float>>(x, R, C, sums, s);
  else launch_stats<Vec<bf16>>(x, R, C, sums, s);
  VCA_LAUNCH_CHECK();
  bn_finalize_kernel<<<(C + 127) / 128, 128, 0, s>>>(sums, R, C, eps, momentum, mean, invstd, running_mean, running_var, sums_prezeroed);
  VCA_LAUNCH_CHECK();
  return VCA_OK;
}
// Batch statistics that a producer kernel already accumulated (vca_conv_fwd_tc_stats): sums = double[2 * C * fold], laid
// out [fold][C] sums then [fold][C] sums of squares, R = rows per channel (all folds together).  Computes mean / invstd,
// updates the running buffers like vca_bn_stats and leaves `sums` zeroed.
int vca_bn_finalize_stats(double* sums, long long R, int C, int fold, float eps, float momentum, float* mean, float* invstd,
                          float* running_mean, float* running_var, cudaStream_t s) {
  VCA_CHECK_ARG(sums && mean && invstd && R > 0 && C > 0 && fold >= 1);
  bn_finalize_fold_kernel<<<(C + 127) / 128, 128, 0, s>>>(sums, R, C, fold, eps, momentum, mean, invstd, running_mean, running_var);
  VCA_LAUNCH_CHECK();
  return VCA_OK;
}
int vca_bn_fold(const float* running_mean, const float* running_var, const float* gamma, const float* beta, const float* bias, int C,
                float eps, float* scale, float* shift, cudaStream_t s) {
  VCA_CHECK_ARG(running_mean && running_var && gamma && beta && scale && shift && C > 0);
  bn_fold_kernel<<<(C + 127) / 128, 128, 0, s>>>(running_mean, running_var, gamma, beta, bias, C, eps, scale, shift);
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
  VCA_CHECK_ARG(x && y && mean && invstd && gamma && beta && R > 0 && C > 0 && act >= ACT_NONE && act <= ACT_RELU &&
                (act != ACT_PRELU || prelu_w));
  const bool ok = dtype == VCA_F32 ? vec_ok<Vec<float>>(x, res, y, 0, 0, C) : vec_ok<Vec<bf16>>(x, res, y, 0, 0, C);
  if (!ok) { vca_set_error("vca_bn_act_fwd: C must be a multiple of %d and tensors 16-byte aligned", dtype == VCA_F32 ? 4 : 8); return VCA_ERR_UNSUPPORTED; }
  BnParams p{mean, invstd, gamma, beta, prelu_w, slope};
  const bool has_res = res != nullptr;
  if (dtype == VCA_F32) { BN_DISPATCH(launch_fwd, Vec<float>, x, res, y, R, C, p, s) }
  else { BN_DISPATCH(launch_fwd, Vec<bf16>, x, res, y, R, C, p, s) }
  VCA_LAUNCH_CHECK();
  return VCA_OK;
}
// sums: device scratch double[3*C].  flags bit 0: sums is a persistent scratch, zero on entry and left zeroed (else it is
// zeroed here); bit 1: dgamma / dbeta / dprelu are ADDED to (gradient accumulation in place).  dres must be given iff
// res is; dgamma/dbeta/dprelu may be null.
int vca_bn_act_bwd(int dtype, const void* dy, const void* x, const void* res, void* dx, void* dres, long long R, int C,
                   const float* mean, const float* invstd, const float* gamma, const float* beta, int act, float slope,
                   const float* prelu_w, int train, double* sums, float* dgamma, float* dbeta, float* dprelu, int flags,
                   cudaStream_t s) {
  VCA_CHECK_ARG(dy && x && dx && mean && invstd && gamma && beta && sums && R > 0 && C > 0 && act >= ACT_NONE &&
                act <= ACT_RELU && (act != ACT_PRELU || prelu_w) && ((res != nullptr) == (dres != nullptr)));
  const bool ok = dtype == VCA_F32 ? vec_ok<Vec<float>>(dy, x, res, dx, dres, C) : vec_ok<Vec<bf16>>(dy, x, res, dx, dres, C);
  if (!ok) { vca_set_error("vca_bn_act_bwd: C must be a multiple of %d and tensors 16-byte aligned", dtype == VCA_F32 ? 4 : 8); return VCA_ERR_UNSUPPORTED; }
  if (!(flags & 1)) cudaMemsetAsync(sums, 0, sizeof(double) * 3 * C, s);
  BnParams p{mean, invstd, gamma, beta, prelu_w, slope};
  const bool has_res = res != nullptr;
  if (dtype == VCA_F32) { BN_DISPATCH(launch_bwd, Vec<float>, dy, x, res, dx, dres, R, C, p, train, sums, s) }
  else if (g_bn_vec == 8) { BN_DISPATCH(launch_bwd, Vec<bf16>, dy, x, res, dx, dres, R, C, p, train, sums, s) }
  else { BN_DISPATCH(launch_bwd, VecH4, dy, x, res, dx, dres, R, C, p, train, sums, s) }
  VCA_LAUNCH_CHECK();
  bn_param_grads_kernel<<<(C + 127) / 128, 128, 0, s>>>(sums, C, dgamma, dbeta, act == ACT_PRELU ? dprelu : nullptr, flags & 1, (flags >> 1) & 1);
  VCA_LAUNCH_CHECK();
  return VCA_OK;
}
// Fused stem tail (bf16, C % 8 == 0, C / 8 a power of two <= 256): y / idx / xmax are [NF, OH, OW, C] with OH = (H-1)/2+1.
int vca_bn_prelu_maxpool_fwd(const void* x, void* y, unsigned char* idx, void* xmax, int NF, int H, int W, int C, const float* mean,
                             const float* invstd, const float* gamma, const float* beta, const float* prelu_w, cudaStream_t s) {
  VCA_CHECK_ARG(x && y && idx && xmax && mean && invstd && gamma && beta && prelu_w && NF > 0 && H > 0 && W > 0 && C > 0);
  const int CV = C / 8;
  if (C % 8 || CV > 256 || (CV & (CV - 1)) || !vca_aligned16(x) || !vca_aligned16(y) || !vca_aligned16(xmax) ||
      (long long)NF * H * W >= (1LL << 31)) {
    vca_set_error("vca_bn_prelu_maxpool_fwd: unsupported shape"); return VCA_ERR_UNSUPPORTED;
  }
  const int OH = (H - 1) / 2 + 1, OW = (W - 1) / 2 + 1;
  BnParams p{mean, invstd, gamma, beta, prelu_w, 0.f};
  const long long total = (long long)NF * OH * OW * CV;
  long long gx = (total + 255) / 256; if (gx > 148LL * 16) gx = 148LL * 16;
  bn_prelu_maxpool_fwd_kernel<Vec<bf16>><<<(unsigned)gx, 256, 0, s>>>((const bf16*)x, (bf16*)y, idx, (bf16*)xmax, NF, H, W, C, OH, OW, p);
  VCA_LAUNCH_CHECK();
  return VCA_OK;
}
// Backward of the fused stem tail.  dy / idx / xmax: pooled tensors; x: the raw conv output; dx: its gradient.  sums: fp64
// scratch [3C] (flags bit 0: persistent and pre-zeroed, left zeroed; bit 1: dgamma / dbeta / dprelu are added to).
int vca_bn_prelu_maxpool_bwd(const void* dy, const unsigned char* idx, const void* xmax, const void* x, void* dx, int NF, int H, int W,
                             int C, const float* mean, const float* invstd, const float* gamma, const float* beta, const float* prelu_w,
                             int train, double* sums, float* dgamma, float* dbeta, float* dprelu, int flags, cudaStream_t s) {
  VCA_CHECK_ARG(dy && idx && xmax && x && dx && mean && invstd && gamma && beta && prelu_w && sums && NF > 0 && H > 0 && W > 0 && C > 0);
  const int CV = C / 8;
  if (C % 8 || CV > 256 || (CV & (CV - 1)) || !vca_aligned16(x) || !vca_aligned16(dy) || !vca_aligned16(xmax) || !vca_aligned16(dx) ||
      (long long)NF * H * W >= (1LL << 31)) {
    vca_set_error("vca_bn_prelu_maxpool_bwd: unsupported shape"); return VCA_ERR_UNSUPPORTED;
  }
  const int OH = (H - 1) / 2 + 1, OW = (W - 1) / 2 + 1;
  if (!(flags & 1)) cudaMemsetAsync(sums, 0, sizeof(double) * 3 * C, s);
  BnParams p{mean, invstd, gamma, beta, prelu_w, 0.f};
  const long long Rp = (long long)NF * OH * OW;
  {   // reduction over the pooled gradient and the raw x at its argmax: sum g', sum g'*xhat, PReLU slope gradient
    RowColGrid g = one_wave_grid<bn_act_bwd_reduce_kernel<VecH4, ACT_PRELU, false>>(Rp, C / 4);
    bn_act_bwd_reduce_kernel<VecH4, ACT_PRELU, false><<<g.grid, g.block, 0, s>>>((const bf16*)dy, (const bf16*)xmax, nullptr, Rp, C, p, sums);
  }
  VCA_LAUNCH_CHECK();
  const long long total = (long long)NF * (H / 2 + 1) * (W / 2 + 1) * CV;
  long long gx = (total + 255) / 256; if (gx > 148LL * 16) gx = 148LL * 16;
  bn_prelu_maxpool_bwd_apply_kernel<<<(unsigned)gx, 256, 0, s>>>((const bf16*)dy, idx, (const bf16*)x, (bf16*)dx, NF, H, W, C, OH, OW, p, sums,
                                                                train);
  VCA_LAUNCH_CHECK();
  bn_param_grads_kernel<<<(C + 127) / 128, 128, 0, s>>>(sums, C, dgamma, dbeta, dprelu, flags & 1, (flags >> 1) & 1);
  VCA_LAUNCH_CHECK();
  return VCA_OK;
}
// out[c] (+)= sum_r x[r,c] (fp32; accumulate != 0 adds to the existing contents).  scratch: device double[C].
int vca_colsum(int dtype, const void* x, long long R, int C, double* scratch, float* out, int accumulate, cudaStream_t s) {
  VCA_CHECK_ARG(x && out && scratch && R > 0 && C > 0);
  const bool ok = dtype == VCA_F32 ? vec_ok<Vec<float>>(x, 0, 0, 0, 0, C) : vec_ok<Vec<bf16>>(x, 0, 0, 0, 0, C);
  if (ok && dtype == VCA_BF16 && accumulate) {
    // bias gradient of a bf16 tensor added into a live fp32 .grad view: one launch -- per-thread fp32 partials, fp64 tree
    // per CTA, one fp32 atomic per channel and CTA straight into the destination (no scratch, no memset, no conversion pass)
    RowColGrid g = one_wave_grid<colsum_vec_kernel<Vec<bf16>, float>>(R, C / 8);
    colsum_vec_kernel<Vec<bf16>, float><<<g.grid, g.block, 0, s>>>((const bf16*)x, R, C, out);
    VCA_LAUNCH_CHECK();
    return VCA_OK;
  }
  cudaMemsetAsync(scratch, 0, sizeof(double) * C, s);
  if (ok && dtype == VCA_F32) {
    RowColGrid g = one_wave_grid<colsum_vec_kernel<Vec<float>>>(R, C / 4);
    colsum_vec_kernel<Vec<float>><<<g.grid, g.block, 0, s>>>((const float*)x, R, C, scratch);
  } else if (ok) {
    RowColGrid g = one_wave_grid<colsum_vec_kernel<Vec<bf16>>>(R, C / 8);
    colsum_vec_kernel<Vec<bf16>><<<g.grid, g.block, 0, s>>>((const bf16*)x, R, C, scratch);
  } else {
    long long gx = (R + 7) / 8; if (gx > 148 * 4) gx = 148 * 4;
    dim3 grid((unsigned)gx, (unsigned)((C + 31) / 32)), block(32, 8);
    if (dtype == VCA_F32) colsum_scalar_kernel<float><<<grid, block, 0, s>>>((const float*)x, R, C, scratch);
    else colsum_scalar_kernel<bf16><<<grid, block, 0, s>>>((const bf16*)x, R, C, scratch);
  }
  VCA_LAUNCH_CHECK();
  d2f_kernel<<<(C + 127) / 128, 128, 0, s>>>(scratch, out, C, accumulate);
  VCA_LAUNCH_CHECK();
  return VCA_OK;
}

}  // extern "C"
