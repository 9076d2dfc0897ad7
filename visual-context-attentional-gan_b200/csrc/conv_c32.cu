// Cin = 1 -> Cout = 32 convolutions (stride 1, 5x5 or 3x3): the mel stems of the three discriminators
// (generator.py:272) at 80x300 / 40x150 / 20x75.  One input channel means 25 MACs per output element and no reuse
// a GEMM tile could exploit; the roofline is the single pass over the 32-channel tensor (y / dY), so the kernels are
// organised around exactly that pass:
//     lane = output channel, a warp walks one row segment of pixels,
// which makes the 32-channel access a coalesced 64/128-byte row per pixel, keeps the lane's 25 weights (forward /
// dgrad) or 25 accumulators (wgrad) in registers, and turns the single-channel side into a KH x KW sliding window
// (static register rotation: the segment length is a multiple of KW) fed by one broadcast shared-memory read per
// filter row per pixel.  dgrad sums over the 32 channels = over lanes: a 32x32 transpose-reduce (31 shuffles per 32
// pixels) leaves pixel j's total in lane j for a coalesced store.
#include "common.cuh"

namespace {

constexpr int WPB = 8;   // warps per CTA

template <int KH, int KW>
__device__ __forceinline__ bool stage_x_rows(float (*sx)[8 * KW + KW], const ConvGeom& g, const void* x, int is_bf16,
                                             long long n, int oh, int ow0, int lane) {
  constexpr int SXW = 8 * KW + KW - 1;
  for (int i = lane; i < KH * SXW; i += 32) {
    const int kh = i / SXW, c = i - kh * SXW;
    const int ih = oh - g.ph + kh, iw = ow0 - g.pw + c;
    float v = 0.f;
    if ((unsigned)ih < (unsigned)g.IH && (unsigned)iw < (unsigned)g.IW) {
      const long long o = (n * g.IH + ih) * (long long)g.IW + iw;
      v = is_bf16 ? __bfloat162float(((const bf16*)x)[o]) : ((const float*)x)[o];
    }
    sx[kh][c] = v;
  }
  return true;
}

// y[n,oh,ow,lane] = b[lane] + sum_{kh,kw} x[n,oh+kh-ph,ow+kw-pw] * w[lane,kh,kw];   wf = [tap][32]
template <class T, int KH, int KW>
__global__ void __launch_bounds__(32 * WPB) c32_fwd_kernel(ConvGeom g, const T* __restrict__ x, const T* __restrict__ wf,
                                                           const float* __restrict__ bias, T* __restrict__ y) {
  constexpr int L = 8 * KW;
  __shared__ float sx[WPB][KH][L + KW];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_w = (g.OW + L - 1) / L;
  const long long ntask = (long long)g.N * g.OH * tiles_w;
  float w[KH * KW];
#pragma unroll
  for (int t = 0; t < KH * KW; ++t) w[t] = to_f(wf[t * 32 + lane]);
  const float b = bias ? bias[lane] : 0.f;
  for (long long task = (long long)blockIdx.x * WPB + warp; task < ntask; task += (long long)gridDim.x * WPB) {
    const int tw = (int)(task % tiles_w); const long long r = task / tiles_w;
    const int oh = (int)(r % g.OH); const long long n = r / g.OH;
    const int ow0 = tw * L;
    stage_x_rows<KH, KW>(sx[warp], g, x, sizeof(T) == 2, n, oh, ow0, lane);
    __syncwarp();
    float ring[KH][KW];
#pragma unroll
    for (int kh = 0; kh < KH; ++kh)
#pragma unroll
      for (int c = 0; c < KW - 1; ++c) ring[kh][c] = sx[warp][kh][c];
    T* yrow = y + ((n * g.OH + oh) * (long long)g.OW + ow0) * 32 + lane;
#pragma unroll
    for (int i = 0; i < L; ++i) {
#pragma unroll
      for (int kh = 0; kh < KH; ++kh) ring[kh][(i + KW - 1) % KW] = sx[warp][kh][i + KW - 1];
      float acc = b;
#pragma unroll
      for (int kh = 0; kh < KH; ++kh)
#pragma unroll
        for (int kw = 0; kw < KW; ++kw) acc = fmaf(w[kh * KW + kw], ring[kh][(i + kw) % KW], acc);
      if (ow0 + i < g.OW) yrow[(long long)i * 32] = from_f<T>(acc);
    }
    __syncwarp();
  }
}

// dw[lane][kh][kw] += sum_{n,oh,ow} dy[n,oh,ow,lane] * x[n,oh+kh-ph,ow+kw-pw]
template <class T, int KH, int KW>
__global__ void __launch_bounds__(32 * WPB) c32_wgrad_kernel(ConvGeom g, const T* __restrict__ dy, const T* __restrict__ x,
                                                             float* __restrict__ dw) {
  constexpr int L = 8 * KW, TAPS = KH * KW;
  __shared__ float sx[WPB][KH][L + KW];
  __shared__ float red[WPB][TAPS][32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_w = (g.OW + L - 1) / L;
  const long long ntask = (long long)g.N * g.OH * tiles_w;
  float acc[TAPS];
#pragma unroll
  for (int t = 0; t < TAPS; ++t) acc[t] = 0.f;
  for (long long task = (long long)blockIdx.x * WPB + warp; task < ntask; task += (long long)gridDim.x * WPB) {
    const int tw = (int)(task % tiles_w); const long long r = task / tiles_w;
    const int oh = (int)(r % g.OH); const long long n = r / g.OH;
    const int ow0 = tw * L;
    // this lane's channel of the L pixels: all loads issued before any use
    const T* dyrow = dy + ((n * g.OH + oh) * (long long)g.OW + ow0) * 32 + lane;
    float dv[L];
#pragma unroll
    for (int i = 0; i < L; ++i) dv[i] = (ow0 + i < g.OW) ? to_f(dyrow[(long long)i * 32]) : 0.f;
    stage_x_rows<KH, KW>(sx[warp], g, x, sizeof(T) == 2, n, oh, ow0, lane);
    __syncwarp();
    float ring[KH][KW];
#pragma unroll
    for (int kh = 0; kh < KH; ++kh)
#pragma unroll
      for (int c = 0; c < KW - 1; ++c) ring[kh][c] = sx[warp][kh][c];
#pragma unroll
    for (int i = 0; i < L; ++i) {
#pragma unroll
      for (int kh = 0; kh < KH; ++kh) ring[kh][(i + KW - 1) % KW] = sx[warp][kh][i + KW - 1];
#pragma unroll
      for (int kh = 0; kh < KH; ++kh)
#pragma unroll
        for (int kw = 0; kw < KW; ++kw) acc[kh * KW + kw] = fmaf(dv[i], ring[kh][(i + kw) % KW], acc[kh * KW + kw]);
    }
    __syncwarp();
  }
#pragma unroll
  for (int t = 0; t < TAPS; ++t) red[warp][t][lane] = acc[t];
  __syncthreads();
  for (int i = threadIdx.x; i < TAPS * 32; i += blockDim.x) {
    const int t = i >> 5, co = i & 31;
    float s = 0.f;
#pragma unroll
    for (int q = 0; q < WPB; ++q) s += red[q][t][co];
    atomicAdd(&dw[co * TAPS + t], s);
  }
}

// dx[n,ih,iw] = sum_{kh,kw,co} dy[n,ih+ph-kh,iw+pw-kw,co] * w[co,kh,kw];   wd = [tap][32]
template <class T, int KH, int KW>
__global__ void __launch_bounds__(32 * WPB) c32_dgrad_kernel(ConvGeom g, const T* __restrict__ dy, const T* __restrict__ wd,
                                                             T* __restrict__ dx) {
  constexpr int L = 32, NV = L + KW - 1;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_w = (g.IW + L - 1) / L;
  const long long ntask = (long long)g.N * g.IH * tiles_w;
  float w[KH * KW];
#pragma unroll
  for (int t = 0; t < KH * KW; ++t) w[t] = to_f(wd[t * 32 + lane]);
  for (long long task = (long long)blockIdx.x * WPB + warp; task < ntask; task += (long long)gridDim.x * WPB) {
    const int tw = (int)(task % tiles_w); const long long r = task / tiles_w;
    const int ih = (int)(r % g.IH); const long long n = r / g.IH;
    const int iw0 = tw * L;
    float acc[L];
#pragma unroll
    for (int i = 0; i < L; ++i) acc[i] = 0.f;
#pragma unroll
    for (int kh = 0; kh < KH; ++kh) {
      const int u = ih + g.ph - kh;                       // warp-uniform
      if ((unsigned)u >= (unsigned)g.OH) continue;
      const T* row = dy + ((n * g.OH + u) * (long long)g.OW) * 32 + lane;
      const int v0 = iw0 + g.pw - (KW - 1);
      float vals[NV];
#pragma unroll
      for (int c = 0; c < NV; ++c) vals[c] = ((unsigned)(v0 + c) < (unsigned)g.OW) ? to_f(row[(long long)(v0 + c) * 32]) : 0.f;
#pragma unroll
      for (int i = 0; i < L; ++i)
#pragma unroll
        for (int kw = 0; kw < KW; ++kw) acc[i] = fmaf(vals[i + KW - 1 - kw], w[kh * KW + kw], acc[i]);
    }
    // transpose-reduce over lanes: afterwards acc[0] of lane j = sum over all lanes of their acc[j]
#pragma unroll
    for (int off = 16, nn = 32; off >= 1; off >>= 1, nn >>= 1) {
      const bool upper = (lane & off) != 0;
#pragma unroll
      for (int k = 0; k < nn / 2; ++k) {
        const float send = upper ? acc[k] : acc[k + nn / 2];
        const float keep = upper ? acc[k + nn / 2] : acc[k];
        acc[k] = keep + __shfl_xor_sync(0xffffffffu, send, off);
      }
    }
    if (iw0 + lane < g.IW) dx[(n * g.IH + ih) * (long long)g.IW + iw0 + lane] = from_f<T>(acc[0]);
  }
}

bool c32_geom(const ConvGeom& g) {
  return g.Cin == 1 && g.Cout == 32 && g.KD == 1 && g.ID == 1 && g.OD == 1 && g.sd == 1 && g.sh == 1 && g.sw == 1 && g.pd == 0 &&
         ((g.KH == 5 && g.KW == 5) || (g.KH == 3 && g.KW == 3));
}
unsigned c32_grid(long long ntask) {
  long long ctas = (ntask + WPB - 1) / WPB;
  const long long cap = 8LL * vca_num_sms();
  return (unsigned)(ctas < cap ? (ctas < 1 ? 1 : ctas) : cap);
}

template <class T>
int c32_fwd_t(const ConvGeom& g, const void* x, const void* wf, const float* bias, void* y, cudaStream_t s) {
  const long long rows = (long long)g.N * g.OH;
  if (g.KH == 5) c32_fwd_kernel<T, 5, 5><<<c32_grid(rows * ((g.OW + 39) / 40)), 32 * WPB, 0, s>>>(g, (const T*)x, (const T*)wf, bias, (T*)y);
  else c32_fwd_kernel<T, 3, 3><<<c32_grid(rows * ((g.OW + 23) / 24)), 32 * WPB, 0, s>>>(g, (const T*)x, (const T*)wf, bias, (T*)y);
  if (cudaGetLastError() != cudaSuccess) { vca_set_error("c32_fwd_kernel launch failed"); return VCA_ERR_CUDA; }
  return 1;
}
template <class T>
int c32_wgrad_t(const ConvGeom& g, const void* dy, const void* x, float* dw, cudaStream_t s) {
  const long long rows = (long long)g.N * g.OH;
  if (g.KH == 5) c32_wgrad_kernel<T, 5, 5><<<c32_grid(rows * ((g.OW + 39) / 40)), 32 * WPB, 0, s>>>(g, (const T*)dy, (const T*)x, dw);
  else c32_wgrad_kernel<T, 3, 3><<<c32_grid(rows * ((g.OW + 23) / 24)), 32 * WPB, 0, s>>>(g, (const T*)dy, (const T*)x, dw);
  if (cudaGetLastError() != cudaSuccess) { vca_set_error("c32_wgrad_kernel launch failed"); return VCA_ERR_CUDA; }
  return 1;
}
template <class T>
int c32_dgrad_t(const ConvGeom& g, const void* dy, const void* wd, void* dx, cudaStream_t s) {
  const unsigned grid = c32_grid((long long)g.N * g.IH * ((g.IW + 31) / 32));
  if (g.KH == 5) c32_dgrad_kernel<T, 5, 5><<<grid, 32 * WPB, 0, s>>>(g, (const T*)dy, (const T*)wd, (T*)dx);
  else c32_dgrad_kernel<T, 3, 3><<<grid, 32 * WPB, 0, s>>>(g, (const T*)dy, (const T*)wd, (T*)dx);
  if (cudaGetLastError() != cudaSuccess) { vca_set_error("c32_dgrad_kernel launch failed"); return VCA_ERR_CUDA; }
  return 1;
}

}  // namespace

// 1 = handled, 0 = not applicable, < 0 = error (same convention as conv_small.cu)
int conv_c32_fwd(int dtype, const ConvGeom& g, const void* x, const void* wf, const float* bias, void* y, cudaStream_t s) {
  if (!c32_geom(g)) return 0;
  return dtype == VCA_F32 ? c32_fwd_t<float>(g, x, wf, bias, y, s) : c32_fwd_t<bf16>(g, x, wf, bias, y, s);
}
int conv_c32_wgrad(int dtype, const ConvGeom& g, const void* dy, const void* x, float* dw, cudaStream_t s) {
  if (!c32_geom(g)) return 0;
  return dtype == VCA_F32 ? c32_wgrad_t<float>(g, dy, x, dw, s) : c32_wgrad_t<bf16>(g, dy, x, dw, s);
}
int conv_c32_dgrad(int dtype, const ConvGeom& g, const void* dy, const void* wd, void* dx, cudaStream_t s) {
  if (!c32_geom(g)) return 0;
  return dtype == VCA_F32 ? c32_dgrad_t<float>(g, dy, wd, dx, s) : c32_dgrad_t<bf16>(g, dy, wd, dx, s);
}
