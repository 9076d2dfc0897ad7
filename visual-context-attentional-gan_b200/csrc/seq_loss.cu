// GRU gate kernels, masked softmax, sync-discriminator losses, GAN/L1 loss reductions, fused Adam, RNG.
// All of these are small HBM/latency-bound kernels in fp32 (the tensors are at most B x T x 1536).
// Reference sites: visual_front.py:20,33-34 (nn.GRU), generator.py:154-171 (AVAttention softmax + key mask),
// generator.py:347-359 (sync losses), generator.py:363-366 (gan_loss), train.py:82-83 (Adam amsgrad).
#include "common.cuh"
#include "vec.cuh"

namespace {

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + __expf(-x)); }

// One GRU time step for `ndir` directions at once.
//   gi  : [ndir][T][B][3H]  input projections incl. b_ih (precomputed by one GEMM)
//   gh  : [ndir][B][3H]     h_prev * W_hh^T (GEMM just before this kernel);  bhh: [ndir][3H]
//   hprev: [ndir][B][H]; out: [T][B][ndir*H]; gates saved as [ndir][T][B][4H] = r, z, n, hn(=gh_n)
// direction d processes time index t_d = (d == 0 ? step : T-1-step).
__global__ void gru_gate_fwd_kernel(const float* __restrict__ gi, const float* __restrict__ gh,
                                    const float* __restrict__ bhh, const float* __restrict__ hprev, float* __restrict__ hnext, float* __restrict__ out,
                                    float* __restrict__ gates, int ndir, int T, int B, int H, int step) {
  long long total = (long long)ndir * B * H;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int j = (int)(i % H); long long r = i / H; int b = (int)(r % B); int d = (int)(r / B);
    int t = d == 0 ? step : T - 1 - step;
    const float* gi_p = gi + (((long long)d * T + t) * B + b) * 3 * H;
    const float* gh_p = gh + ((long long)d * B + b) * 3 * H;
    const float* bh = bhh + (long long)d * 3 * H;
    float rr = sigmoidf_(gi_p[j] + gh_p[j] + bh[j]);
    float zz = sigmoidf_(gi_p[H + j] + gh_p[H + j] + bh[H + j]);
    float hn = gh_p[2 * H + j] + bh[2 * H + j];
    float nn = tanhf(gi_p[2 * H + j] + rr * hn);
    float hp = hprev[((long long)d * B + b) * H + j];
    float h = (1.f - zz) * nn + zz * hp;
    hnext[((long long)d * B + b) * H + j] = h;
    out[((long long)t * B + b) * (ndir * H) + d * H + j] = h;
    float* gs = gates + (((long long)d * T + t) * B + b) * 4 * H;
    gs[j] = rr; gs[H + j] = zz; gs[2 * H + j] = nn; gs[3 * H + j] = hn;
  }
}
// Backward of one step.  dh_total = dout[t] + dh_carry.  Writes dgi[d][t][b][3H], dgh[d][t][b][3H] and the
// elementwise part of dh_prev (dh_total * z) into dh_carry; the GEMM dgh * W_hh is added by the caller.
// hprev_all: out tensor [T][B][ndir*H]; h_{t-1} for dir 0 is out[t-1], for dir 1 is out[t+1]; zero at the ends.
__global__ void gru_gate_bwd_kernel(const float* __restrict__ dout, float* __restrict__ dh_carry,
                                    const float* __restrict__ gates, const float* __restrict__ out,
                                    float* __restrict__ dgi, float* __restrict__ dgh, float* __restrict__ dgh_cur,
                                    int ndir, int T, int B, int H, int step) {
  long long total = (long long)ndir * B * H;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int j = (int)(i % H); long long r = i / H; int b = (int)(r % B); int d = (int)(r / B);
    // backward visits time in the reverse order of the forward
    int t = d == 0 ? T - 1 - step : step;
    int tp = d == 0 ? t - 1 : t + 1;
    float hp = (tp >= 0 && tp < T) ? out[((long long)tp * B + b) * (ndir * H) + d * H + j] : 0.f;
    const float* gs = gates + (((long long)d * T + t) * B + b) * 4 * H;
    float rr = gs[j], zz = gs[H + j], nn = gs[2 * H + j], hn = gs[3 * H + j];
    long long ci = ((long long)d * B + b) * H + j;
    float dh = dout[((long long)t * B + b) * (ndir * H) + d * H + j] + dh_carry[ci];
    float dn = dh * (1.f - zz);
    float dz = dh * (hp - nn);
    float dn_pre = dn * (1.f - nn * nn);
    float dr = dn_pre * hn;
    float dr_pre = dr * rr * (1.f - rr);
    float dz_pre = dz * zz * (1.f - zz);
    long long go = (((long long)d * T + t) * B + b) * 3 * H;
    dgi[go + j] = dr_pre; dgi[go + H + j] = dz_pre; dgi[go + 2 * H + j] = dn_pre;
    dgh[go + j] = dr_pre; dgh[go + H + j] = dz_pre; dgh[go + 2 * H + j] = dn_pre * rr;
    const long long cg = ((long long)d * B + b) * 3 * H;   // compact copy of this step's dgh for the recurrence GEMM
    dgh_cur[cg + j] = dr_pre; dgh_cur[cg + H + j] = dz_pre; dgh_cur[cg + 2 * H + j] = dn_pre * rr;
    dh_carry[ci] = dh * zz;
  }
}

// Skinny batched GEMM for the GRU recurrence (few rows, fp32, exact):  out[z,b,n] = beta*out[z,b,n] + sum_k in[z,b,k] * wt[z,n,k].
// One CTA owns 8 output columns n for every row b: its 8 weight rows (8 x K) and a 32-row slab of `in` are staged
// in shared memory (row stride K+4 floats: conflict-free 128-bit reads), threads = 8 (n) x 32 (b).
constexpr int SK_N = 8, SK_B = 32, SK_KC = 512;
__global__ void __launch_bounds__(256) skinny_gemm_kernel(const float* __restrict__ in, const float* __restrict__ wt,
                                                          float* __restrict__ out, int Bn, int N, int K, float beta) {
  extern __shared__ float sk_smem[];
  float* sW = sk_smem;                       // [SK_N][SK_KC]
  float* sI = sk_smem + SK_N * SK_KC;        // [SK_B][SK_KC + 4]
  const int z = blockIdx.y, n0 = blockIdx.x * SK_N;
  const int jj = threadIdx.x >> 5, bb = threadIdx.x & 31;
  const float* inz = in + (long long)z * Bn * K;
  const float* wz = wt + (long long)z * N * K;
  float* outz = out + (long long)z * Bn * N;
  for (int b0 = 0; b0 < Bn; b0 += SK_B) {
    float acc = 0.f;
    for (int k0 = 0; k0 < K; k0 += SK_KC) {
      const int kc = min(SK_KC, K - k0);   // multiple of 4 (checked on the host)
      __syncthreads();
      for (int i = threadIdx.x; i < SK_N * (kc >> 2); i += 256) {
        const int r = i / (kc >> 2), c4 = i - r * (kc >> 2);
        float4 v = make_float4(0, 0, 0, 0);
        if (n0 + r < N) v = *reinterpret_cast<const float4*>(wz + (long long)(n0 + r) * K + k0 + c4 * 4);
        *reinterpret_cast<float4*>(sW + r * SK_KC + c4 * 4) = v;
      }
      for (int i = threadIdx.x; i < SK_B * (kc >> 2); i += 256) {
        const int r = i / (kc >> 2), c4 = i - r * (kc >> 2);
        float4 v = make_float4(0, 0, 0, 0);
        if (b0 + r < Bn) v = *reinterpret_cast<const float4*>(inz + (long long)(b0 + r) * K + k0 + c4 * 4);
        *reinterpret_cast<float4*>(sI + r * (SK_KC + 4) + c4 * 4) = v;
      }
      __syncthreads();
      const float4* w4 = reinterpret_cast<const float4*>(sW + jj * SK_KC);
      const float4* i4 = reinterpret_cast<const float4*>(sI + bb * (SK_KC + 4));
#pragma unroll 4
      for (int k = 0; k < (kc >> 2); ++k) {
        const float4 a = w4[k], h = i4[k];
        acc = fmaf(a.x, h.x, acc); acc = fmaf(a.y, h.y, acc); acc = fmaf(a.z, h.z, acc); acc = fmaf(a.w, h.w, acc);
      }
    }
    if (b0 + bb < Bn && n0 + jj < N) {
      float* o = outz + (long long)(b0 + bb) * N + n0 + jj;
      *o = beta != 0.f ? fmaf(beta, *o, acc) : acc;
    }
  }
}

// Row softmax with a key-length mask: p[z,r,:] = softmax(x[z,r,:len[z]]), zeros beyond len (generator.py:161-164).
// One warp per row.
__global__ void masked_softmax_fwd_kernel(const float* __restrict__ x, float* __restrict__ p, const int* __restrict__ lens,
                                          int Z, int Rr, int S) {
  int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= Z * Rr) return;
  int z = row / Rr;
  int L = lens ? min(max(lens[z], 0), S) : S;
  const float* xr = x + (long long)row * S; float* pr = p + (long long)row * S;
  float m = -INFINITY;
  for (int i = lane; i < L; i += 32) m = fmaxf(m, xr[i]);
  m = warp_max(m);
  float sum = 0.f;
  for (int i = lane; i < L; i += 32) sum += __expf(xr[i] - m);
  sum = warp_sum(sum);
  float inv = 1.f / sum;  // L == 0 gives NaN rows exactly like the reference
  for (int i = lane; i < S; i += 32) pr[i] = i < L ? __expf(xr[i] - m) * inv : (L == 0 ? NAN : 0.f);
}
// dx = p * (dp - sum(dp*p))
__global__ void softmax_bwd_kernel(const float* __restrict__ dp, const float* __restrict__ p, float* __restrict__ dx, int rows,
                                   int S) {
  int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* a = dp + (long long)row * S; const float* b = p + (long long)row * S; float* o = dx + (long long)row * S;
  float dot = 0.f;
  for (int i = lane; i < S; i += 32) dot = fmaf(a[i], b[i], dot);
  dot = warp_sum(dot);
  for (int i = lane; i < S; i += 32) o[i] = b[i] * (a[i] - dot);
}

// L2-normalise rows: y = x / max(||x||, eps)  (F.normalize, eps 1e-12); norms saved.
__global__ void l2norm_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, float* __restrict__ norms, int rows,
                                  int D, float eps) {
  int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* a = x + (long long)row * D; float* o = y + (long long)row * D;
  float ss = 0.f;
  for (int i = lane; i < D; i += 32) ss = fmaf(a[i], a[i], ss);
  ss = warp_sum(ss);
  float n = sqrtf(ss);
  float dn = fmaxf(n, eps);
  if (lane == 0) norms[row] = n;
  for (int i = lane; i < D; i += 32) o[i] = a[i] / dn;
}
// dx = (dy - y*(y.dy)) / max(n,eps)   (for n > eps; for n <= eps the clamp is constant: dx = dy/eps)
__global__ void l2norm_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y, const float* __restrict__ norms,
                                  float* __restrict__ dx, int rows, int D, float eps) {
  int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* g = dy + (long long)row * D; const float* yy = y + (long long)row * D; float* o = dx + (long long)row * D;
  float n = norms[row];
  float dot = 0.f;
  if (n > eps) { for (int i = lane; i < D; i += 32) dot = fmaf(g[i], yy[i], dot); dot = warp_sum(dot); }
  float inv = 1.f / fmaxf(n, eps);
  for (int i = lane; i < D; i += 32) o[i] = (g[i] - (n > eps ? yy[i] * dot : 0.f)) * inv;
}

// InfoNCE on sim[B,S,S] (generator.py:354-359): loss[b] = -0.5*(mean_i logsoftmax_row(sim)[i,i] + mean_i logsoftmax_col(sim)[i,i])
// One CTA per batch element; also emits dsim (for dloss[b] == 1) so the backward is a scale.
__global__ void nce_diag_kernel(const float* __restrict__ sim, float* __restrict__ loss, float* __restrict__ dsim, int S) {
  extern __shared__ float sh[];  // rowlse[S], collse[S], red[40]
  float* rowlse = sh; float* collse = sh + S; float* red = sh + 2 * S;
  const float* a = sim + (long long)blockIdx.x * S * S;
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int i = w; i < S; i += nw) {
    float m = -INFINITY;
    for (int j = lane; j < S; j += 32) m = fmaxf(m, a[i * S + j]);
    m = warp_max(m);
    float s = 0.f;
    for (int j = lane; j < S; j += 32) s += __expf(a[i * S + j] - m);
    s = warp_sum(s);
    if (lane == 0) rowlse[i] = m + __logf(s);
    m = -INFINITY;
    for (int j = lane; j < S; j += 32) m = fmaxf(m, a[j * S + i]);
    m = warp_max(m);
    s = 0.f;
    for (int j = lane; j < S; j += 32) s += __expf(a[j * S + i] - m);
    s = warp_sum(s);
    if (lane == 0) collse[i] = m + __logf(s);
  }
  __syncthreads();
  float acc = 0.f;
  for (int i = threadIdx.x; i < S; i += blockDim.x) acc += 2.f * a[i * S + i] - rowlse[i] - collse[i];
  float tot = block_sum(acc, red);
  if (threadIdx.x == 0) loss[blockIdx.x] = -0.5f * tot / (float)S;
  if (dsim) {
    float* d = dsim + (long long)blockIdx.x * S * S;
    float c = -0.5f / (float)S;
    for (int e = threadIdx.x; e < S * S; e += blockDim.x) {
      int i = e / S, j = e % S;
      float v = a[e];
      float g = -__expf(v - rowlse[i]) - __expf(v - collse[j]);
      if (i == j) g += 2.f;
      d[e] = c * g;
    }
  }
}

// 5 - mean_t |cos(v_t, a_t)| per batch element (generator.py:347-349), eps 1e-8 on each norm (torch >= 1.12
// clamps the norms separately).  One CTA per b; one warp per t.  Saves cos, |v|, |a| for the backward.
__global__ void cos_abs_mean_fwd_kernel(const float* __restrict__ v, const float* __restrict__ a, float* __restrict__ loss,
                                        float* __restrict__ saved /*[B][S][3]*/, int S, int D, float eps) {
  __shared__ float red[40];
  int b = blockIdx.x, lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  float acc = 0.f;
  for (int t = w; t < S; t += nw) {
    const float* x = v + ((long long)b * S + t) * D; const float* y = a + ((long long)b * S + t) * D;
    float xy = 0.f, xx = 0.f, yy = 0.f;
    for (int i = lane; i < D; i += 32) { xy = fmaf(x[i], y[i], xy); xx = fmaf(x[i], x[i], xx); yy = fmaf(y[i], y[i], yy); }
    xy = warp_sum(xy); xx = warp_sum(xx); yy = warp_sum(yy);
    float nx = fmaxf(sqrtf(xx), eps), ny = fmaxf(sqrtf(yy), eps);
    float c = xy / (nx * ny);
    if (lane == 0) {
      float* sv = saved + ((long long)b * S + t) * 3; sv[0] = c; sv[1] = nx; sv[2] = ny;
      acc += fabsf(c);
    }
  }
  float tot = block_sum(acc, red);
  if (threadIdx.x == 0) loss[b] = 5.f - tot / (float)S;
}
// gradient w.r.t. a (and optionally v): d|c|/da = sign(c) * (v/(nv*na) - c*a/na^2), scaled by -dloss[b]/S
__global__ void cos_abs_mean_bwd_kernel(const float* __restrict__ dloss, const float* __restrict__ v,
                                        const float* __restrict__ a, const float* __restrict__ saved, float* __restrict__ da,
                                        float* __restrict__ dv, int Bn, int S, int D) {
  long long total = (long long)Bn * S * D;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long bt = i / D; int b = (int)(bt / S);
    const float* sv = saved + bt * 3;
    float c = sv[0], nx = sv[1], ny = sv[2];
    float sg = c > 0.f ? 1.f : (c < 0.f ? -1.f : 0.f);
    float k = -dloss[b] / (float)S * sg;
    if (da) da[i] = k * (v[i] / (nx * ny) - c * a[i] / (ny * ny));
    if (dv) dv[i] = k * (a[i] / (nx * ny) - c * v[i] / (nx * nx));
  }
}

// mean softplus(sign*x) (beta 1, threshold 20): out[0] = loss; dx (optional) = sign*sigmoid(sign*x)/n
__global__ void softplus_mean_kernel(const float* __restrict__ x, float* __restrict__ out, float* __restrict__ dx, int n,
                                     float sign) {
  __shared__ float red[40];
  float acc = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    float v = sign * x[i];
    acc += v > 20.f ? v : log1pf(__expf(v));
    if (dx) dx[i] = sign * (v > 20.f ? 1.f : sigmoidf_(v)) / (float)n;
  }
  float tot = block_sum(acc, red);
  if (threadIdx.x == 0) out[0] = tot / (float)n;
}

// L1 / sum-of-squares reductions: out[0] += scale * sum |a-b|   (mode 0)   or   scale * sum a^2   (mode 1)
template <class T>
__global__ void reduce_l1_sq_kernel(const T* __restrict__ a, const T* __restrict__ b, long long n, float scale, int mode,
                                    float* __restrict__ out) {
  __shared__ float red[40];
  float acc = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float v = to_f(a[i]);
    if (mode == 0) acc += fabsf(v - to_f(b[i])); else acc = fmaf(v, v, acc);
  }
  float tot = block_sum(acc, red);
  if (threadIdx.x == 0) atomicAdd(out, tot * scale);
}
// d/da of scale*sum|a-b| : g * scale * sign(a-b)
template <class T>
__global__ void l1_bwd_kernel(const T* __restrict__ a, const T* __restrict__ b, const float* __restrict__ g, long long n,
                              float scale, T* __restrict__ da) {
  float gs = g[0] * scale;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float d = to_f(a[i]) - to_f(b[i]);
    da[i] = from_f<T>(d > 0.f ? gs : (d < 0.f ? -gs : 0.f));
  }
}

// Fused Adam / AMSGrad over a flat fp32 buffer (torch.optim.Adam semantics, train.py:82-83):
//   g += wd*p; m = b1 m + (1-b1) g; v = b2 v + (1-b2) g^2; vmax = max(vmax, v);
//   p -= lr/bc1 * m / (sqrt(vmax)/sqrt(bc2) + eps)
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                            float* __restrict__ vmax, long long n, float lr, float b1, float b2, float eps, float wd,
                            float bc1, float bc2_sqrt, float gscale, const int* __restrict__ step_dev,
                            const float* __restrict__ lr_dev) {
  if (lr_dev) lr = *lr_dev;   // learning rate lives on the device: a MultiStepLR decay (train.py:85-86) reaches captured graphs
  if (step_dev) {   // step count lives on the device (CUDA-graph replay): bias corrections computed here
    const float t = (float)(*step_dev);
    bc1 = 1.f - powf(b1, t);
    bc2_sqrt = sqrtf(1.f - powf(b2, t));
  }
  float step = lr / bc1;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float pi = p[i];
    float gi = fmaf(wd, pi, g[i] * gscale);
    float mi = b1 * m[i] + (1.f - b1) * gi;
    float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi; v[i] = vi;
    float vh = vi;
    if (vmax) { vh = fmaxf(vmax[i], vi); vmax[i] = vh; }
    p[i] = pi - step * mi / (sqrtf(vh) / bc2_sqrt + eps);
  }
}

// The same update on four parameters per thread with 16-byte loads / stores (the flat parameter buffers are 16-byte aligned
// segments): 4x the bytes in flight per thread of this purely HBM-bound pass (36 B per parameter).  Element-wise arithmetic
// identical to adam_kernel.
__global__ void __launch_bounds__(256) adam_vec4_kernel(float4* __restrict__ p, const float4* __restrict__ g, float4* __restrict__ m,
                                                        float4* __restrict__ v, float4* __restrict__ vmax, long long n4, float lr, float b1,
                                                        float b2, float eps, float wd, float bc1, float bc2_sqrt, float gscale,
                                                        const int* __restrict__ step_dev, const float* __restrict__ lr_dev) {
  if (lr_dev) lr = *lr_dev;
  if (step_dev) {
    const float t = (float)(*step_dev);
    bc1 = 1.f - powf(b1, t);
    bc2_sqrt = sqrtf(1.f - powf(b2, t));
  }
  const float step = lr / bc1;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 p4 = p[i], g4 = g[i], m4 = m[i], v4 = v[i];
    float4 x4 = vmax ? vmax[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    float pa[4] = {p4.x, p4.y, p4.z, p4.w}, ga[4] = {g4.x, g4.y, g4.z, g4.w}, ma[4] = {m4.x, m4.y, m4.z, m4.w};
    float va[4] = {v4.x, v4.y, v4.z, v4.w}, xa[4] = {x4.x, x4.y, x4.z, x4.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float gi = fmaf(wd, pa[k], ga[k] * gscale);
      const float mi = b1 * ma[k] + (1.f - b1) * gi;
      const float vi = b2 * va[k] + (1.f - b2) * gi * gi;
      ma[k] = mi; va[k] = vi;
      float vh = vi;
      if (vmax) { vh = fmaxf(xa[k], vi); xa[k] = vh; }
      pa[k] = pa[k] - step * mi / (sqrtf(vh) / bc2_sqrt + eps);
    }
    m[i] = make_float4(ma[0], ma[1], ma[2], ma[3]);
    v[i] = make_float4(va[0], va[1], va[2], va[3]);
    if (vmax) vmax[i] = make_float4(xa[0], xa[1], xa[2], xa[3]);
    p[i] = make_float4(pa[0], pa[1], pa[2], pa[3]);
  }
}
static inline bool adam_vec_ok(const void* p, const void* g, const void* m, const void* v, const void* vmax, long long n) {
  return n % 4 == 0 && vca_aligned16(p) && vca_aligned16(g) && vca_aligned16(m) && vca_aligned16(v) && (!vmax || vca_aligned16(vmax));
}

// Philox-4x32-10 counter RNG
__device__ __forceinline__ uint4 philox(uint4 ctr, uint2 key) {
  const unsigned M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    unsigned hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
    unsigned hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += W0; key.y += W1;
  }
  return ctr;
}
__device__ __forceinline__ float u01(unsigned x) { return ((float)(x >> 8) + 0.5f) * (1.f / 16777216.f); }

// mode 0: standard normal (Box-Muller); mode 1: dropout keep-mask scaled by 1/(1-p) (param = p)
template <class T>
__global__ void rng_kernel(T* __restrict__ out, long long n, unsigned long long seed, unsigned long long offset, int mode,
                           float param, const unsigned long long* __restrict__ ctr_dev) {
  if (ctr_dev) offset += *ctr_dev;   // device-resident stream position (advanced by counter_add_kernel): graph-replay safe
  long long nq = (n + 3) / 4;
  for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < nq; q += (long long)gridDim.x * blockDim.x) {
    unsigned long long c = (unsigned long long)q + offset;
    uint4 r = philox(make_uint4((unsigned)c, (unsigned)(c >> 32), 0u, 0u), make_uint2((unsigned)seed, (unsigned)(seed >> 32)));
    float o[4];
    if (mode == 0) {
      float u0 = u01(r.x), u1 = u01(r.y), u2 = u01(r.z), u3 = u01(r.w);
      float ra = sqrtf(-2.f * __logf(u0)), rb = sqrtf(-2.f * __logf(u2));
      float s0, c0, s1, c1;
      __sincosf(6.2831853071795865f * u1, &s0, &c0);
      __sincosf(6.2831853071795865f * u3, &s1, &c1);
      o[0] = ra * c0; o[1] = ra * s0; o[2] = rb * c1; o[3] = rb * s1;
    } else if (mode == 2) {   // uniform angle in (-pi, pi): Griffin-Lim initial phase (audio_processing.py:59)
      o[0] = 6.2831853071795865f * u01(r.x) - 3.14159265358979f; o[1] = 6.2831853071795865f * u01(r.y) - 3.14159265358979f;
      o[2] = 6.2831853071795865f * u01(r.z) - 3.14159265358979f; o[3] = 6.2831853071795865f * u01(r.w) - 3.14159265358979f;
    } else {
      float keep = 1.f / (1.f - param);
      o[0] = u01(r.x) >= param ? keep : 0.f; o[1] = u01(r.y) >= param ? keep : 0.f;
      o[2] = u01(r.z) >= param ? keep : 0.f; o[3] = u01(r.w) >= param ? keep : 0.f;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) if (q * 4 + k < n) out[q * 4 + k] = from_f<T>(o[k]);
  }
}

__global__ void counter_add_kernel(unsigned long long* ctr, unsigned long long inc) { *ctr += inc; }
__global__ void counter_add_i32_kernel(int* ctr, int inc) { *ctr += inc; }

// y = x * mask  (elementwise, same dtype)
template <class T>
__global__ void mul_kernel(const T* __restrict__ x, const T* __restrict__ m, T* __restrict__ y, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    y[i] = from_f<T>(to_f(x[i]) * to_f(m[i]));
}

}  // namespace

#define DISPATCH_T(dtype, CALL_F32, CALL_BF16) \
  do { if ((dtype) == VCA_F32) { CALL_F32; } else { CALL_BF16; } } while (0)

extern "C" {

int vca_gru_gate_fwd(const float* gi, const float* gh, const float* bhh, const float* hprev, float* hnext, float* out,
                     float* gates, int ndir, int T, int B, int H, int step, cudaStream_t s) {
  VCA_CHECK_ARG(gi && gh && bhh && hprev && hnext && out && gates && ndir > 0 && T > 0 && B > 0 && H > 0 && step >= 0 && step < T);
  gru_gate_fwd_kernel<<<vca_grid_1d((long long)ndir * B * H, 128), 128, 0, s>>>(gi, gh, bhh, hprev, hnext, out, gates, ndir, T, B,
                                                                                H, step);
  VCA_LAUNCH_CHECK();
  return VCA_OK;
}
int vca_gru_gate_bwd(const float* dout, float* dh_carry, const float* gates, const float* out, float* dgi, float* dgh,
                     float* dgh_cur, int ndir, int T, int B, int H, int step, cudaStream_t s) {
  VCA_CHECK_ARG(dout && dh_carry && gates && out && dgi && dgh && dgh_cur && ndir > 0 && T > 0 && B > 0 && H > 0 && step >= 0 &&
                step < T);
  gru_gate_bwd_kernel<<<vca_grid_1d((long long)ndir * B * H, 128), 128, 0, s>>>(dout, dh_carry, gates, out, dgi, dgh, dgh_cur, ndir, T,
                                                                                B, H, step);
  VCA_LAUNCH_CHECK();
  return VCA_OK;
}
// out[z,b,n] = beta*out + sum_k in[z,b,k]*wt[z,n,k]; all fp32 contiguous; K % 4 == 0.  (GRU recurrence GEMMs.)
int vca_skinny_gemm(const float* in, const float* wt, float* out, int Z, int Bn, int N, int K, float beta, cudaStream_t s) {
  VCA_CHECK_ARG(in && wt && out && Z > 0 && Bn > 0 && N > 0 && K > 0 && K % 4 == 0);
  static bool attr_set = false;
  const size_t smem = (size_t)(SK_N * SK_KC + SK_B * (SK_KC + 4)) * sizeof(float);
  if (!attr_set) {
    if (cudaFuncSetAttribute(skinny_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
      vca_set_error("cudaFuncSetAttribute(skinny_gemm_kernel) failed"); return VCA_ERR_CUDA;
    }
    attr_set = true;
  }
  dim3 grid((N + SK_N - 1) / SK_N, Z);
  skinny_gemm_kernel<<<grid, 256, smem, s>>>(in, wt, out, Bn, N, K, beta);
  VCA_LAUNCH_CHECK();
  return VCA_OK;
}
int vca_masked_softmax_fwd(const float* x, float* p, const int* lens, int Z, int R, int S, cudaStream_t s) {
  VCA_CHECK_ARG(x && p && Z > 0 && R > 0 && S > 0);
  int rows = Z * R;
  masked_softmax_fwd_kernel<<<(rows + 3) / 4, 128, 0, s>>>(x, p, lens, Z, R, S);
  VCA_LAUNCH_CHECK();
  return VCA_OK;
}
int vca_softmax_bwd(const float* dp, const float* p, float* dx, int rows, int S, cudaStream_t s) {
  VCA_CHECK_ARG(dp && p && dx && rows > 0 && S > 0);
  softmax_bwd_kernel<<<(rows + 3) / 4, 128, 0, s>>>(dp, p, dx, rows, S);
  VCA_LAUNCH_CHECK();
  return VCA_OK;
}
int vca_l2norm_fwd(const float* x, float* y, float* norms, int rows, int D, float eps, cudaStream_t s) {
  VCA_CHECK_ARG(x && y && norms && rows > 0 && D > 0);
  l2norm_fwd_kernel<<<(rows + 3) / 4, 128, 0, s>>>(x, y, norms, rows, D, eps);
  VCA_LAUNCH_CHECK();
  return VCA_OK;
}
int vca_l2norm_bwd(const float* dy, const float* y, const float* norms, float* dx, int rows, int D, float eps,
                   cudaStream_t s) {
  VCA_CHECK_ARG(dy && y && norms && dx && rows > 0 && D > 0);
  l2norm_bwd_kernel<<<(rows + 3) / 4, 128, 0, s>>>(dy, y, norms, dx, rows, D, eps);
  VCA_LAUNCH_CHECK();
  return VCA_OK;
}
int vca_nce_diag(const float* sim, float* loss, float* dsim, int Bn, int S, cudaStream_t s) {
  VCA_CHECK_ARG(sim && loss && Bn > 0 && S > 0 && S <= 4096);
  nce_diag_kernel<<<Bn, 256, (2 * S + 40) * sizeof(float), s>>>(sim, loss, dsim, S);
  VCA_LAUNCH_CHECK();
  return VCA_OK;
}
int vca_cos_abs_mean_fwd(const float* v, const float* a, float* loss, float* saved, int Bn, int S, int D, cudaStream_t s) {
  VCA_CHECK_ARG(v && a && loss && saved && Bn > 0 && S > 0 && D > 0);
  cos_abs_mean_fwd_kernel<<<Bn, 256, 0, s>>>(v, a, loss, saved, S, D, 1e-8f);
  VCA_LAUNCH_CHECK();
  return VCA_OK;
}
int vca_cos_abs_mean_bwd(const float* dloss, const float* v, const float* a, const float* saved, float* da, float* dv, int Bn,
                         int S, int D, cudaStream_t s) {
  VCA_CHECK_ARG(dloss && v && a && saved && (da || dv) && Bn > 0 && S > 0 && D > 0);
  cos_abs_mean_bwd_kernel<<<vca_grid_1d((long long)Bn * S * D, 256), 256, 0, s>>>(dloss, v, a, saved, da, dv, Bn, S, D);
  VCA_LAUNCH_CHECK();
  return VCA_OK;
}
int vca_softplus_mean(const float* x, float* out, float* dx, int n, float sign, cudaStream_t s) {
  VCA_CHECK_ARG(x && out && n > 0);
  softplus_mean_kernel<<<1, 256, 0, s>>>(x, out, dx, n, sign);
  VCA_LAUNCH_CHECK();
  return VCA_OK;
}
// out[0] = scale*sum|a-b| (mode 0) or scale*sum a^2 (mode 1); out zeroed here.
int vca_reduce_l1_sq(int dtype, const void* a, const void* b, long long n, float scale, int mode, float* out, cudaStream_t s) {
  VCA_CHECK_ARG(a && out && n > 0 && (mode == 1 || b));
  cudaMemsetAsync(out, 0, sizeof(float), s);
  unsigned grid = vca_grid_1d(n, 256, 8);
  DISPATCH_T(dtype, (reduce_l1_sq_kernel<float><<<grid, 256, 0, s>>>((const float*)a, (const float*)b, n, scale, mode, out)),
             (reduce_l1_sq_kernel<bf16><<<grid, 256, 0, s>>>((const bf16*)a, (const bf16*)b, n, scale, mode, out)));
  VCA_LAUNCH_CHECK();
  return VCA_OK;
}
int vca_l1_bwd(int dtype, const void* a, const void* b, const float* g, long long n, float scale, void* da, cudaStream_t s) {
  VCA_CHECK_ARG(a && b && g && da && n > 0);
  unsigned grid = vca_grid_1d(n, 256, 4);
  DISPATCH_T(dtype, (l1_bwd_kernel<float><<<grid, 256, 0, s>>>((const float*)a, (const float*)b, g, n, scale, (float*)da)),
             (l1_bwd_kernel<bf16><<<grid, 256, 0, s>>>((const bf16*)a, (const bf16*)b, g, n, scale, (bf16*)da)));
  VCA_LAUNCH_CHECK();
  return VCA_OK;
}
// vmax == null -> plain Adam (train_LRS.py:97-98); step >= 1; gscale multiplies the gradient (1/world for DP means).
int vca_adam_step(float* p, const float* g, float* m, float* v, float* vmax, long long n, float lr, float beta1, float beta2,
                  float eps, float weight_decay, int step, float gscale, cudaStream_t s) {
  VCA_CHECK_ARG(p && g && m && v && n > 0 && step >= 1);
  float bc1 = 1.f - powf(beta1, (float)step);
  float bc2s = sqrtf(1.f - powf(beta2, (float)step));
  adam_kernel<<<vca_grid_1d(n, 256, 4), 256, 0, s>>>(p, g, m, v, vmax, n, lr, beta1, beta2, eps, weight_decay, bc1, bc2s,
                                                     gscale, nullptr, nullptr);
  VCA_LAUNCH_CHECK();
  return VCA_OK;
}
// Same, with the step counter AND the learning rate resident on the device: *step_dev is incremented, then used for
// the bias corrections; *lr_dev is read by the kernel.  This form can be captured in a CUDA graph and replayed (the
// host never sees the step number, and an lr schedule only has to write one float between replays).
// bump = 0: *step_dev is used as it is (the second of two calls that update disjoint slices of one parameter group in the
// same optimizer step, e.g. while the all-reduce of the other slice's gradients is still running).
int vca_adam_step_dev(float* p, const float* g, float* m, float* v, float* vmax, long long n, const float* lr_dev, float beta1,
                      float beta2, float eps, float weight_decay, int* step_dev, float gscale, int bump, cudaStream_t s) {
  VCA_CHECK_ARG(p && g && m && v && n > 0 && step_dev && lr_dev);
  if (bump) counter_add_i32_kernel<<<1, 1, 0, s>>>(step_dev, 1);
  if (adam_vec_ok(p, g, m, v, vmax, n))
    adam_vec4_kernel<<<vca_grid_1d(n / 4, 256, 2), 256, 0, s>>>((float4*)p, (const float4*)g, (float4*)m, (float4*)v, (float4*)vmax, n / 4, 0.f,
                                                                beta1, beta2, eps, weight_decay, 1.f, 1.f, gscale, step_dev, lr_dev);
  else
  adam_kernel<<<vca_grid_1d(n, 256, 4), 256, 0, s>>>(p, g, m, v, vmax, n, 0.f, beta1, beta2, eps, weight_decay, 1.f, 1.f, gscale,
                                                     step_dev, lr_dev);
  VCA_LAUNCH_CHECK();
  return VCA_OK;
}
int vca_rng(int dtype, void* out, long long n, unsigned long long seed, unsigned long long offset, int mode, float param,
            cudaStream_t s) {
  VCA_CHECK_ARG(out && n > 0 && (mode == 0 || (mode == 1 && param >= 0.f && param < 1.f)));
  unsigned grid = vca_grid_1d((n + 3) / 4, 256);
  DISPATCH_T(dtype, (rng_kernel<float><<<grid, 256, 0, s>>>((float*)out, n, seed, offset, mode, param, nullptr)),
             (rng_kernel<bf16><<<grid, 256, 0, s>>>((bf16*)out, n, seed, offset, mode, param, nullptr)));
  VCA_LAUNCH_CHECK();
  return VCA_OK;
}
// Same, with the Philox stream position resident on the device (*ctr_dev, advanced here by ceil(n/4)): replaying a
// captured graph draws fresh numbers every time.
int vca_rng_dev(int dtype, void* out, long long n, unsigned long long seed, unsigned long long* ctr_dev, int mode, float param,
                cudaStream_t s) {
  VCA_CHECK_ARG(out && ctr_dev && n > 0 && (mode == 0 || mode == 2 || (mode == 1 && param >= 0.f && param < 1.f)));
  unsigned grid = vca_grid_1d((n + 3) / 4, 256);
  DISPATCH_T(dtype, (rng_kernel<float><<<grid, 256, 0, s>>>((float*)out, n, seed, 0ull, mode, param, ctr_dev)),
             (rng_kernel<bf16><<<grid, 256, 0, s>>>((bf16*)out, n, seed, 0ull, mode, param, ctr_dev)));
  counter_add_kernel<<<1, 1, 0, s>>>(ctr_dev, (unsigned long long)((n + 3) / 4));
  VCA_LAUNCH_CHECK();
  return VCA_OK;
}
}  // extern "C"
