// Waveform tail and mel front of the test-time path:
//   * inverse_mel front   (src/data/vid_aud_grid.py:190-200): denormalize -> exp -> (B,T,80) x mel_basis (80,321) -> *1000
//   * LRS inverse_spec front (src/data/vid_aud_lrs2.py:257-263): denormalize -> exp -> *14  (exp_affine_kernel)
//   * mel_spectrogram tail (vid_aud_grid.py:291-307): mel_basis x |STFT| -> log(clamp(., 1e-5))
//   * de-emphasis + clip  (vid_aud_grid.py:205-209, 230-232): scipy lfilter([1], [1, -0.97]) per waveform on the host,
//     then np.clip(-1, 1).  Here: one CTA per clip, the first-order recurrence as a block-wide scan of affine maps in
//     fp64 (lfilter runs in float64), the clip fused into the store.
// All of it is HBM-bound and tiny next to Griffin-Lim; the point is that the waveform never leaves the device.
#include "common.cuh"

namespace {

// ---------------------------------------------------------------------------------------------------------------
// out[b][f][t] = post( sum_k pre(in[b][k][t]) * w[k][f] )
//   pre : 0 identity, 1 exp(in * pre_mul + pre_add)
//   post: 0 v * post_arg, 1 log(max(v, post_arg))
// block = 32 time steps x 8 output lanes; the pre-processed input tile [K][32] sits in shared memory; the basis value
// w[k][f] is the same for the 32 lanes of a warp (one broadcast load through L1).
// ---------------------------------------------------------------------------------------------------------------
constexpr int FB_TT = 32, FB_FL = 8;

__global__ void __launch_bounds__(FB_TT * FB_FL)
filterbank_kernel(const float* __restrict__ in, const float* __restrict__ w, float* __restrict__ out,
                  int K, int F, int T, int pre, int post, float pre_mul, float pre_add, float post_arg) {
  extern __shared__ float tile[];   // [K][FB_TT]
  const int b = blockIdx.y, t0 = blockIdx.x * FB_TT;
  const int tx = threadIdx.x & (FB_TT - 1), ty = threadIdx.x / FB_TT;
  const float* src = in + (size_t)b * K * T;
  for (int k = ty; k < K; k += FB_FL) {
    const int t = t0 + tx;
    float v = t < T ? src[(size_t)k * T + t] : 0.f;
    if (pre == 1) v = expf(fmaf(v, pre_mul, pre_add));
    tile[k * FB_TT + tx] = t < T ? v : 0.f;
  }
  __syncthreads();
  const int t = t0 + tx;
  float* dst = out + (size_t)b * F * T;
  for (int f = ty; f < F; f += FB_FL) {
    float acc = 0.f;
#pragma unroll 4
    for (int k = 0; k < K; ++k) acc = fmaf(tile[k * FB_TT + tx], __ldg(w + (size_t)k * F + f), acc);
    const float r = post == 1 ? logf(fmaxf(acc, post_arg)) : acc * post_arg;
    if (t < T) dst[(size_t)f * T + t] = r;
  }
}

__global__ void exp_affine_kernel(const float* __restrict__ x, float* __restrict__ y, long long n,
                                  float mul, float add, float scale) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    y[i] = expf(fmaf(x[i], mul, add)) * scale;
}

// ---------------------------------------------------------------------------------------------------------------
// y[n] = x[n] + a * y[n-1]  (y[-1] = 0), then clamp to [lo, hi].  One CTA per clip; 1024 threads x 4 samples per tile.
// Each thread folds its 4 samples into the affine map  c -> A c + V  (A = a^4); an inclusive scan of those maps over
// the block gives every thread the filter state entering its samples.
// ---------------------------------------------------------------------------------------------------------------
constexpr int DE_THREADS = 1024, DE_PER = 4;

__global__ void __launch_bounds__(DE_THREADS)
deemph_clip_kernel(const float* __restrict__ x, float* __restrict__ y, int L, double a, float lo, float hi) {
  __shared__ double sA[32], sV[32];
  __shared__ double s_carry;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  x += (size_t)blockIdx.x * L;
  y += (size_t)blockIdx.x * L;
  const double a2 = a * a, a3 = a2 * a, a4 = a2 * a2;
  double carry = 0.0;   // y[base - 1]
#pragma unroll 1
  for (int base = 0; base < L; base += DE_THREADS * DE_PER) {
    const int i0 = base + tid * DE_PER;
    float v[DE_PER];
#pragma unroll
    for (int j = 0; j < DE_PER; ++j) v[j] = i0 + j < L ? x[i0 + j] : 0.f;
    const double l0 = v[0], l1 = fma(a, l0, (double)v[1]), l2 = fma(a, l1, (double)v[2]), l3 = fma(a, l2, (double)v[3]);
    double A = a4, V = l3;   // this thread's map, then the inclusive composition over the warp
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const double Ap = __shfl_up_sync(0xffffffffu, A, o), Vp = __shfl_up_sync(0xffffffffu, V, o);
      if (lane >= o) { V = fma(Vp, A, V); A = Ap * A; }
    }
    if (lane == 31) { sA[wid] = A; sV[wid] = V; }
    __syncthreads();
    if (wid == 0) {
      double Aw = sA[lane], Vw = sV[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const double Ap = __shfl_up_sync(0xffffffffu, Aw, o), Vp = __shfl_up_sync(0xffffffffu, Vw, o);
        if (lane >= o) { Vw = fma(Vp, Aw, Vw); Aw = Ap * Aw; }
      }
      sA[lane] = Aw; sV[lane] = Vw;
    }
    __syncthreads();
    // state at the end of this thread's samples, counted from the tile's incoming carry
    double Ai = A, Vi = V;
    if (wid > 0) { Vi = fma(sV[wid - 1], A, V); Ai = sA[wid - 1] * A; }
    const double yend = fma(carry, Ai, Vi);
    double c = __shfl_up_sync(0xffffffffu, yend, 1);
    if (lane == 0) c = wid == 0 ? carry : fma(carry, sA[wid - 1], sV[wid - 1]);
    const double r[DE_PER] = {fma(a, c, l0), fma(a2, c, l1), fma(a3, c, l2), fma(a4, c, l3)};
#pragma unroll
    for (int j = 0; j < DE_PER; ++j)
      if (i0 + j < L) y[i0 + j] = fminf(fmaxf((float)r[j], lo), hi);
    if (tid == DE_THREADS - 1) s_carry = yend;
    __syncthreads();
    carry = s_carry;
  }
}

}  // namespace

extern "C" {

int vca_filterbank_apply(const float* in, const float* w, float* out, int B, int K, int F, int T, int pre, int post,
                         float pre_mul, float pre_add, float post_arg, cudaStream_t s) {
  VCA_CHECK_ARG(in && w && out && B > 0 && K > 0 && F > 0 && T > 0 && B <= 65535);
  VCA_CHECK_ARG((pre == 0 || pre == 1) && (post == 0 || post == 1));
  const size_t smem = (size_t)K * FB_TT * sizeof(float);
  VCA_CHECK_ARG(smem <= 48 * 1024);
  dim3 grid((T + FB_TT - 1) / FB_TT, B);
  filterbank_kernel<<<grid, FB_TT * FB_FL, smem, s>>>(in, w, out, K, F, T, pre, post, pre_mul, pre_add, post_arg);
  VCA_LAUNCH_CHECK();
  return VCA_OK;
}

int vca_exp_affine(const float* x, float* y, long long n, float mul, float add, float scale, cudaStream_t s) {
  VCA_CHECK_ARG(x && y && n > 0);
  exp_affine_kernel<<<vca_grid_1d(n, 256), 256, 0, s>>>(x, y, n, mul, add, scale);
  VCA_LAUNCH_CHECK();
  return VCA_OK;
}

int vca_deemphasis_clip(const float* x, float* y, int B, int L, double coef, float lo, float hi, cudaStream_t s) {
  VCA_CHECK_ARG(x && y && B > 0 && L > 0 && lo <= hi);
  deemph_clip_kernel<<<B, DE_THREADS, 0, s>>>(x, y, L, coef, lo, hi);
  VCA_LAUNCH_CHECK();
  return VCA_OK;
}

}  // extern "C"
