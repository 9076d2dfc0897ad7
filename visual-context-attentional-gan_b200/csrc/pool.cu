// Pooling / resampling kernels on channels-last tensors [N(frames), H, W, C]; each thread owns 4/8 consecutive
// channels of one pixel (128-bit accesses, coalesced along C).
// Reference sites: visual_front.py:14 (MaxPool3d (1,3,3)/(1,2,2)/(0,1,1)), generator.py:74,83 (avg_pool2d 2),
// generator.py:112,121 (nearest x2), generator.py:140 / resnet.py:82 (spatial mean), visual_front.py:11 (stem im2col).
#include "vec.cuh"

namespace {

#define GRID_STRIDE(i, total) \
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < (total); i += (long long)gridDim.x * blockDim.x)


// Flat index -> (i0, i1, i2, i3) with extents (n0, n1, n2, *), innermost first.  In 32 bits whenever the index fits (it
// always does for the shapes of the path): a 64-bit div/mod costs ~100 instructions, four of them per element made
// these one-load-one-store kernels issue-bound.
__device__ __forceinline__ void split4(long long i, int n0, int n1, int n2, int& i0, int& i1, int& i2, int& i3) {
  if (i < (1LL << 32)) {
    unsigned r = (unsigned)i;
    i0 = (int)(r % (unsigned)n0); r /= (unsigned)n0;
    i1 = (int)(r % (unsigned)n1); r /= (unsigned)n1;
    i2 = (int)(r % (unsigned)n2); i3 = (int)(r / (unsigned)n2);
  } else {
    long long r = i;
    i0 = (int)(r % n0); r /= n0;
    i1 = (int)(r % n1); r /= n1;
    i2 = (int)(r % n2); i3 = (int)(r / n2);
  }
}

// 3x3 stride-2 pad-1 max pool per frame; idx = argmax position 0..8 (first max wins, like ATen).
template <class T, class VT>
__global__ void maxpool3x3s2_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, unsigned char* __restrict__ idx, int NF,
                                        int H, int W, int C, int OH, int OW) {
  constexpr int V = VT::N;
  const int CV = C / V;
  GRID_STRIDE(i, (long long)NF * OH * OW * CV) {
    int cv, ow, oh, n;
    split4(i, CV, OW, OH, cv, ow, oh, n);
    float best[V]; int bi[V];
#pragma unroll
    for (int k = 0; k < V; ++k) { best[k] = -INFINITY; bi[k] = 0; }
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      int h = oh * 2 - 1 + kh;
      if ((unsigned)h >= (unsigned)H) continue;
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        int w = ow * 2 - 1 + kw;
        if ((unsigned)w >= (unsigned)W) continue;
        float v[V];
        VT::load(x + (((long long)n * H + h) * W + w) * C + cv * V, v);
#pragma unroll
        for (int k = 0; k < V; ++k)
          if (v[k] > best[k] || (v[k] != v[k] && best[k] == best[k])) { best[k] = v[k]; bi[k] = kh * 3 + kw; }
      }
    }
    const long long o = i * V;
    VT::store(y + o, best);
    if (V == 8) {        // the 8 argmax codes as one 64-bit store
      unsigned lo = 0, hi = 0;
#pragma unroll
      for (int k = 0; k < 4; ++k) { lo |= (unsigned)bi[k] << (8 * k); hi |= (unsigned)bi[(4 + k) % V] << (8 * k); }
      *reinterpret_cast<uint2*>(idx + o) = make_uint2(lo, hi);
    } else if (V == 4) {
      unsigned lo = 0;
#pragma unroll
      for (int k = 0; k < 4; ++k) lo |= (unsigned)bi[k % V] << (8 * k);
      *reinterpret_cast<unsigned*>(idx + o) = lo;
    } else {
#pragma unroll
      for (int k = 0; k < V; ++k) idx[o + k] = (unsigned char)bi[k];
    }
  }
}
// gather form of the backward: every input pixel looks at the <=4 windows that contain it.
template <class T, class VT>
__global__ void maxpool3x3s2_bwd_kernel(const T* __restrict__ dy, const unsigned char* __restrict__ idx, T* __restrict__ dx,
                                        int NF, int H, int W, int C, int OH, int OW) {
  constexpr int V = VT::N;
  const int CV = C / V;
  GRID_STRIDE(i, (long long)NF * H * W * CV) {
    int cv, w, h, n;
    split4(i, CV, W, H, cv, w, h, n);
    float acc[V];
#pragma unroll
    for (int k = 0; k < V; ++k) acc[k] = 0.f;
    // windows oh cover rows 2oh-1..2oh+1: at most 2 x 2 windows contain this pixel.  All (<= 4) gradient / argmax
    // loads are issued before any is used (the kernel is latency bound otherwise: 4 dependent L2 round trips).
    const int ohs[2] = {h >> 1, (h + 1) >> 1}, ows[2] = {w >> 1, (w + 1) >> 1};
    const bool vh[2] = {ohs[0] < OH, ohs[1] < OH && ohs[1] != ohs[0]}, vw[2] = {ows[0] < OW, ows[1] < OW && ows[1] != ows[0]};
    float g[4][V];
    unsigned long long codes[4];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = 0; b < 2; ++b) {
        const int q = a * 2 + b;
        codes[q] = 0xffffffffffffffffull;                      // matches no argmax code
#pragma unroll
        for (int k = 0; k < V; ++k) g[q][k] = 0.f;
        if (vh[a] && vw[b]) {
          const long long o = (((long long)n * OH + ohs[a]) * OW + ows[b]) * C + cv * V;
          VT::load(dy + o, g[q]);
          if (V == 8) codes[q] = *reinterpret_cast<const unsigned long long*>(idx + o);
          else if (V == 4) codes[q] = *reinterpret_cast<const unsigned*>(idx + o);
          else codes[q] = idx[o];
        }
      }
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = 0; b < 2; ++b) {
        const int q = a * 2 + b;
        const unsigned code = (unsigned)((h - (ohs[a] * 2 - 1)) * 3 + (w - (ows[b] * 2 - 1)));
#pragma unroll
        for (int k = 0; k < V; ++k)
          if ((unsigned)((codes[q] >> (8 * k)) & 0xffull) == code) acc[k] += g[q][k];
      }
    VT::store(dx + i * V, acc);
  }
}
// y[oh,ow] = scale * sum of the 2x2 block (scale 0.25: avg_pool2d(x,2); scale 1: backward of nearest x2).
template <class T, class VT>
__global__ void pool2x2_sum_kernel(const T* __restrict__ x, T* __restrict__ y, int NF, int H, int W, int C, int OH, int OW,
                                   float scale) {
  constexpr int V = VT::N;
  const int CV = C / V;
  GRID_STRIDE(i, (long long)NF * OH * OW * CV) {
    int cv, ow, oh, n;
    split4(i, CV, OW, OH, cv, ow, oh, n);
    const T* p = x + (((long long)n * H + oh * 2) * W + ow * 2) * C + cv * V;
    float a[V], b[V], c[V], d[V];
    VT::load(p, a); VT::load(p + C, b); VT::load(p + (long long)W * C, c); VT::load(p + (long long)W * C + C, d);
#pragma unroll
    for (int k = 0; k < V; ++k) a[k] = (a[k] + b[k] + c[k] + d[k]) * scale;
    VT::store(y + i * V, a);
  }
}
// y[h,w] = scale * x[h/2, w/2] when (h/2 < IH && w/2 < IW) else 0.  scale 1: nearest x2; 0.25: backward of avg pool.
template <class T, class VT>
__global__ void expand2x2_kernel(const T* __restrict__ x, T* __restrict__ y, int NF, int IH, int IW, int C, int H, int W,
                                 float scale) {
  constexpr int V = VT::N;
  const int CV = C / V;
  GRID_STRIDE(i, (long long)NF * H * W * CV) {
    int cv, w, h, n;
    split4(i, CV, W, H, cv, w, h, n);
    const int ih = h >> 1, iw = w >> 1;
    float v[V];
#pragma unroll
    for (int k = 0; k < V; ++k) v[k] = 0.f;
    if (ih < IH && iw < IW) {
      VT::load(x + (((long long)n * IH + ih) * IW + iw) * C + cv * V, v);
#pragma unroll
      for (int k = 0; k < V; ++k) v[k] *= scale;
    }
    VT::store(y + i * V, v);
  }
}
// y[n,c] = scale * sum_p x[n,p,c]
template <class T, class VT>
__global__ void spatial_sum_kernel(const T* __restrict__ x, T* __restrict__ y, int NF, int P, int C, float scale) {
  constexpr int V = VT::N;
  const int CV = C / V;
  GRID_STRIDE(i, (long long)NF * CV) {
    int cv = (int)(i % CV); int n = (int)(i / CV);
    const T* p = x + (long long)n * P * C + cv * V;
    float a[V];
#pragma unroll
    for (int k = 0; k < V; ++k) a[k] = 0.f;
    for (int q = 0; q < P; ++q) {
      float v[V];
      VT::load(p + (long long)q * C, v);
#pragma unroll
      for (int k = 0; k < V; ++k) a[k] += v[k];
    }
#pragma unroll
    for (int k = 0; k < V; ++k) a[k] *= scale;
    VT::store(y + i * V, a);
  }
}
// y[n,p,c] = scale * x[n,c]
template <class T, class VT>
__global__ void spatial_bcast_kernel(const T* __restrict__ x, T* __restrict__ y, int NF, int P, int C, float scale) {
  constexpr int V = VT::N;
  const int CV = C / V;
  GRID_STRIDE(i, (long long)NF * P * CV) {
    int cv = (int)(i % CV); int n = (int)(i / ((long long)P * CV));
    float v[V];
    VT::load(x + (long long)n * C + cv * V, v);
#pragma unroll
    for (int k = 0; k < V; ++k) v[k] *= scale;
    VT::store(y + i * V, v);
  }
}

// Space-to-depth with a zero border for stride-2 3x3 / pad-1 convolutions:
//   y[n, i, j, (pa*2+pb)*C + c] = x[n, 2i+pa-1, 2j+pb-1, c]   (0 outside), y is [NF, H2, W2, 4C].
// A 3x3 stride-2 conv over x is then a 2x2 stride-1 conv over y (see ops.conv_s2), which runs on the tcgen05 path.
template <class T, class VT>
__global__ void s2d_kernel(const T* __restrict__ x, T* __restrict__ y, int NF, int H, int W, int C, int H2, int W2) {
  constexpr int V = VT::N;
  const int CV = C / V;
  GRID_STRIDE(i, (long long)NF * H2 * W2 * 4 * CV) {
    int cv = (int)(i % CV); long long r = i / CV;
    int par = (int)(r % 4); r /= 4;
    int j = (int)(r % W2); r /= W2; int ii = (int)(r % H2); int n = (int)(r / H2);
    const int h = 2 * ii + (par >> 1) - 1, w = 2 * j + (par & 1) - 1;
    float v[V];
#pragma unroll
    for (int k = 0; k < V; ++k) v[k] = 0.f;
    if ((unsigned)h < (unsigned)H && (unsigned)w < (unsigned)W) VT::load(x + (((long long)n * H + h) * W + w) * C + cv * V, v);
    VT::store(y + i * V, v);
  }
}
// inverse gather (the transpose of s2d): x[n,h,w,c] = y[n,(h+1)/2,(w+1)/2,(((h+1)&1)*2+((w+1)&1))*C + c]
template <class T, class VT>
__global__ void d2s_kernel(const T* __restrict__ y, T* __restrict__ x, int NF, int H, int W, int C, int H2, int W2) {
  constexpr int V = VT::N;
  const int CV = C / V;
  GRID_STRIDE(i, (long long)NF * H * W * CV) {
    int cv, w, h, n;
    split4(i, CV, W, H, cv, w, h, n);
    const int ii = (h + 1) >> 1, j = (w + 1) >> 1, par = (((h + 1) & 1) << 1) | ((w + 1) & 1);
    float v[V];
#pragma unroll
    for (int k = 0; k < V; ++k) v[k] = 0.f;
    if (ii < H2 && j < W2) VT::load(y + ((((long long)n * H2 + ii) * W2 + j) * 4 + par) * C + cv * V, v);
    VT::store(x + i * V, v);
  }
}

// im2col of the visual front-end stem over its two spatial kernel dims (visual_front.py:11: k=(5,7,7), s=(1,2,2),
// p=(2,3,3), Cin=1): y[f, oh, ow, kh*7+kw] = x[f, 2oh-3+kh, 2ow-3+kw] for the 49 taps, channels 49..63 zero.
// The remaining temporal 5-tap convolution (64 -> 64 channels) then runs as a (5,1) conv on the tcgen05 path.
template <class TI, class TO>
__global__ void stem_im2col_kernel(const TI* __restrict__ x, TO* __restrict__ y, long long NF, int H, int W, int OH, int OW) {
  GRID_STRIDE(i, NF * OH * OW * 8) {   // 8 channel-groups of 8
    int cg = (int)(i & 7); long long r = i >> 3;
    int ow = (int)(r % OW); r /= OW; int oh = (int)(r % OH); long long f = r / OH;
    float v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int t = cg * 8 + k;
      const int kh = t / 7, kw = t - kh * 7;
      const int h = 2 * oh - 3 + kh, w = 2 * ow - 3 + kw;
      v[k] = (t < 49 && (unsigned)h < (unsigned)H && (unsigned)w < (unsigned)W) ? to_f(x[(f * H + h) * W + w]) : 0.f;
    }
    TO* o = y + i * 8;   // 8 channels = one 128-bit (bf16) / two 128-bit (fp32) stores
    if (sizeof(TO) == 2) Vec<bf16>::store(reinterpret_cast<bf16*>(o), v);
    else { Vec<float>::store(reinterpret_cast<float*>(o), v); Vec<float>::store(reinterpret_cast<float*>(o) + 4, v + 4); }
  }
}

// Tiled form: a CTA stages the (2*RB+5) x (W+6) input rows of one band of RB output rows in shared memory (zero padded,
// coalesced loads), then every thread emits (pixel, 8-channel group) items from shared memory -- 8 LDS + one 128-bit
// store instead of 8 scattered global loads per store; the kernel is then bound by the 963 MB it has to write.
template <class TI, class TO>
__global__ void __launch_bounds__(256) stem_im2col_tiled_kernel(const TI* __restrict__ x, TO* __restrict__ y, int H, int W, int OH,
                                                                int OW, int bands, int IM2COL_RB) {
  extern __shared__ float sx[];
  const long long f = blockIdx.x / bands;
  const int oh0 = (blockIdx.x % bands) * IM2COL_RB;
  const int PW = W + 6, rows = 2 * IM2COL_RB + 5, h0 = 2 * oh0 - 3;
  for (int i = threadIdx.x; i < rows * PW; i += blockDim.x) {
    const int r = i / PW, c = i - r * PW;
    const int h = h0 + r, w = c - 3;
    sx[i] = ((unsigned)h < (unsigned)H && (unsigned)w < (unsigned)W) ? to_f(x[(f * H + h) * W + w]) : 0.f;
  }
  __syncthreads();
  // thread = (channel group cg, pixel lane): its 8 taps' offsets inside the staged band are constants, the pixel walks
  // forward 32 at a time with an incremental (row, column) -- the item loop used to spend ~40 integer instructions per
  // 16-byte store on index arithmetic (tap -> (kh, kw), item -> (row, column)).  8 consecutive threads still cover one
  // pixel's 128 bytes, so a warp stores four whole lines.
  const int cg = threadIdx.x & 7, plane = threadIdx.x >> 3;
  int off[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int t = cg * 8 + k;
    const int kh = (t * 37) >> 8, kw = t - kh * 7;          // t / 7 for t < 64
    off[k] = t < 49 ? kh * PW + kw : -1;
  }
  int r = plane / OW, ow = plane - r * OW;
  const int rstep = 32 / OW, cstep = 32 - rstep * OW;
  for (; r < IM2COL_RB && oh0 + r < OH; ) {
    const float* b = sx + (2 * r) * PW + 2 * ow;
    float v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = off[k] >= 0 ? b[off[k]] : 0.f;
    TO* o = y + (((f * OH + oh0 + r) * OW + ow) * 64 + cg * 8);
    if (sizeof(TO) == 2) Vec<bf16>::store(reinterpret_cast<bf16*>(o), v);
    else { Vec<float>::store(reinterpret_cast<float*>(o), v); Vec<float>::store(reinterpret_cast<float*>(o) + 4, v + 4); }
    r += rstep; ow += cstep;
    if (ow >= OW) { ow -= OW; ++r; }
  }
}

template <class T>
bool use_vec(int C, const void* a, const void* b) { return C % Vec<T>::N == 0 && vca_aligned16(a) && vca_aligned16(b); }

}  // namespace

#define POOL_DISPATCH_T(KERNEL, T, total_per_c, a_ptr, b_ptr, ...)                                             \
  do {                                                                                                          \
    if (use_vec<T>(C, a_ptr, b_ptr))                                                                            \
      KERNEL<T, Vec<T>><<<vca_grid_1d((total_per_c) * (C / Vec<T>::N), 256), 256, 0, s>>>(__VA_ARGS__);         \
    else                                                                                                        \
      KERNEL<T, Vec1<T>><<<vca_grid_1d((total_per_c) * C, 256), 256, 0, s>>>(__VA_ARGS__);                      \
    VCA_LAUNCH_CHECK();                                                                                         \
  } while (0)

#define TP(T, p) ((T*)(p))
#define CTP(T, p) ((const T*)(p))

int g_im2col_rb = 14;  // output rows per CTA of the tiled stem im2col ("im2col_rb")

extern "C" {

int vca_maxpool3x3s2_fwd(int dtype, const void* x, void* y, unsigned char* idx, int NF, int H, int W, int C, cudaStream_t s) {
  VCA_CHECK_ARG(x && y && idx && NF > 0 && H > 0 && W > 0 && C > 0);
  int OH = (H + 2 - 3) / 2 + 1, OW = (W + 2 - 3) / 2 + 1;
  long long tot = (long long)NF * OH * OW;
  if (dtype == VCA_F32) POOL_DISPATCH_T(maxpool3x3s2_fwd_kernel, float, tot, x, y, CTP(float, x), TP(float, y), idx, NF, H, W, C, OH, OW);
  else POOL_DISPATCH_T(maxpool3x3s2_fwd_kernel, bf16, tot, x, y, CTP(bf16, x), TP(bf16, y), idx, NF, H, W, C, OH, OW);
  return VCA_OK;
}
int vca_maxpool3x3s2_bwd(int dtype, const void* dy, const unsigned char* idx, void* dx, int NF, int H, int W, int C,
                         cudaStream_t s) {
  VCA_CHECK_ARG(dy && dx && idx && NF > 0 && H > 0 && W > 0 && C > 0);
  int OH = (H + 2 - 3) / 2 + 1, OW = (W + 2 - 3) / 2 + 1;
  long long tot = (long long)NF * H * W;
  if (dtype == VCA_F32) POOL_DISPATCH_T(maxpool3x3s2_bwd_kernel, float, tot, dy, dx, CTP(float, dy), idx, TP(float, dx), NF, H, W, C, OH, OW);
  else POOL_DISPATCH_T(maxpool3x3s2_bwd_kernel, bf16, tot, dy, dx, CTP(bf16, dy), idx, TP(bf16, dx), NF, H, W, C, OH, OW);
  return VCA_OK;
}
// y[NF, H/2, W/2, C] = scale * (2x2 block sums of x[NF,H,W,C])
int vca_pool2x2_sum(int dtype, const void* x, void* y, int NF, int H, int W, int C, float scale, cudaStream_t s) {
  VCA_CHECK_ARG(x && y && NF > 0 && H >= 2 && W >= 2 && C > 0);
  int OH = H / 2, OW = W / 2;
  long long tot = (long long)NF * OH * OW;
  if (dtype == VCA_F32) POOL_DISPATCH_T(pool2x2_sum_kernel, float, tot, x, y, CTP(float, x), TP(float, y), NF, H, W, C, OH, OW, scale);
  else POOL_DISPATCH_T(pool2x2_sum_kernel, bf16, tot, x, y, CTP(bf16, x), TP(bf16, y), NF, H, W, C, OH, OW, scale);
  return VCA_OK;
}
// y[NF,H,W,C] = scale * x[NF,IH,IW,C] replicated 2x2 (zero where h/2>=IH or w/2>=IW)
int vca_expand2x2(int dtype, const void* x, void* y, int NF, int IH, int IW, int C, int H, int W, float scale,
                  cudaStream_t s) {
  VCA_CHECK_ARG(x && y && NF > 0 && IH > 0 && IW > 0 && C > 0 && H >= 2 * IH && W >= 2 * IW && H <= 2 * IH + 1 &&
                W <= 2 * IW + 1);
  long long tot = (long long)NF * H * W;
  if (dtype == VCA_F32) POOL_DISPATCH_T(expand2x2_kernel, float, tot, x, y, CTP(float, x), TP(float, y), NF, IH, IW, C, H, W, scale);
  else POOL_DISPATCH_T(expand2x2_kernel, bf16, tot, x, y, CTP(bf16, x), TP(bf16, y), NF, IH, IW, C, H, W, scale);
  return VCA_OK;
}
int vca_spatial_sum(int dtype, const void* x, void* y, int NF, int P, int C, float scale, cudaStream_t s) {
  VCA_CHECK_ARG(x && y && NF > 0 && P > 0 && C > 0);
  long long tot = NF;
  if (dtype == VCA_F32) POOL_DISPATCH_T(spatial_sum_kernel, float, tot, x, y, CTP(float, x), TP(float, y), NF, P, C, scale);
  else POOL_DISPATCH_T(spatial_sum_kernel, bf16, tot, x, y, CTP(bf16, x), TP(bf16, y), NF, P, C, scale);
  return VCA_OK;
}
int vca_spatial_bcast(int dtype, const void* x, void* y, int NF, int P, int C, float scale, cudaStream_t s) {
  VCA_CHECK_ARG(x && y && NF > 0 && P > 0 && C > 0);
  long long tot = (long long)NF * P;
  if (dtype == VCA_F32) POOL_DISPATCH_T(spatial_bcast_kernel, float, tot, x, y, CTP(float, x), TP(float, y), NF, P, C, scale);
  else POOL_DISPATCH_T(spatial_bcast_kernel, bf16, tot, x, y, CTP(bf16, x), TP(bf16, y), NF, P, C, scale);
  return VCA_OK;
}
// x [NF,H,W,C] -> y [NF,H2,W2,4C] with H2 = OH+1, W2 = OW+1 for OH = (H-1)/2+1 (3x3/s2/p1 output size)
int vca_s2d(int dtype, const void* x, void* y, int NF, int H, int W, int C, int H2, int W2, cudaStream_t s) {
  VCA_CHECK_ARG(x && y && NF > 0 && H > 0 && W > 0 && C > 0 && 2 * H2 >= H + 1 && 2 * W2 >= W + 1);
  long long tot = (long long)NF * H2 * W2 * 4;
  if (dtype == VCA_F32) POOL_DISPATCH_T(s2d_kernel, float, tot, x, y, CTP(float, x), TP(float, y), NF, H, W, C, H2, W2);
  else POOL_DISPATCH_T(s2d_kernel, bf16, tot, x, y, CTP(bf16, x), TP(bf16, y), NF, H, W, C, H2, W2);
  return VCA_OK;
}
int vca_d2s(int dtype, const void* y, void* x, int NF, int H, int W, int C, int H2, int W2, cudaStream_t s) {
  VCA_CHECK_ARG(x && y && NF > 0 && H > 0 && W > 0 && C > 0 && 2 * H2 >= H + 1 && 2 * W2 >= W + 1);
  long long tot = (long long)NF * H * W;
  if (dtype == VCA_F32) POOL_DISPATCH_T(d2s_kernel, float, tot, y, x, CTP(float, y), TP(float, x), NF, H, W, C, H2, W2);
  else POOL_DISPATCH_T(d2s_kernel, bf16, tot, y, x, CTP(bf16, y), TP(bf16, x), NF, H, W, C, H2, W2);
  return VCA_OK;
}
// x: [NF,H,W] (Cin = 1) fp32 or bf16; y: [NF,OH,OW,64] in dt_out, OH = (H-1)/2+1.
int vca_stem_im2col(int dt_in, int dt_out, const void* x, void* y, long long NF, int H, int W, cudaStream_t s) {
  VCA_CHECK_ARG(x && y && NF > 0 && H > 0 && W > 0);
  int OH = (H + 6 - 7) / 2 + 1, OW = (W + 6 - 7) / 2 + 1;
  const int IM2COL_RB = g_im2col_rb > 0 ? g_im2col_rb : 4;
  const int bands = (OH + IM2COL_RB - 1) / IM2COL_RB;
  const size_t smem = (size_t)(2 * IM2COL_RB + 5) * (W + 6) * sizeof(float);
  if (g_im2col_rb > 0 && smem <= 48 * 1024 && NF * bands < (1ll << 31)) {
    const unsigned g = (unsigned)(NF * bands);
    if (dt_in == VCA_F32 && dt_out == VCA_F32) stem_im2col_tiled_kernel<float, float><<<g, 256, smem, s>>>((const float*)x, (float*)y, H, W, OH, OW, bands, IM2COL_RB);
    else if (dt_in == VCA_F32) stem_im2col_tiled_kernel<float, bf16><<<g, 256, smem, s>>>((const float*)x, (bf16*)y, H, W, OH, OW, bands, IM2COL_RB);
    else if (dt_out == VCA_F32) stem_im2col_tiled_kernel<bf16, float><<<g, 256, smem, s>>>((const bf16*)x, (float*)y, H, W, OH, OW, bands, IM2COL_RB);
    else stem_im2col_tiled_kernel<bf16, bf16><<<g, 256, smem, s>>>((const bf16*)x, (bf16*)y, H, W, OH, OW, bands, IM2COL_RB);
    VCA_LAUNCH_CHECK();
    return VCA_OK;
  }
  unsigned grid = vca_grid_1d(NF * OH * OW * 8, 256);
  if (dt_in == VCA_F32 && dt_out == VCA_F32) stem_im2col_kernel<float, float><<<grid, 256, 0, s>>>((const float*)x, (float*)y, NF, H, W, OH, OW);
  else if (dt_in == VCA_F32) stem_im2col_kernel<float, bf16><<<grid, 256, 0, s>>>((const float*)x, (bf16*)y, NF, H, W, OH, OW);
  else if (dt_out == VCA_F32) stem_im2col_kernel<bf16, float><<<grid, 256, 0, s>>>((const bf16*)x, (float*)y, NF, H, W, OH, OW);
  else stem_im2col_kernel<bf16, bf16><<<grid, 256, 0, s>>>((const bf16*)x, (bf16*)y, NF, H, W, OH, OW);
  VCA_LAUNCH_CHECK();
  return VCA_OK;
}

}  // extern "C"
