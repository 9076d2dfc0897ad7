// Pooling / resampling kernels on channels-last tensors [N(frames), H, W, C]; threads run along C (coalesced).
// Reference sites: visual_front.py:14 (MaxPool3d (1,3,3)/(1,2,2)/(0,1,1)), generator.py:74,83 (avg_pool2d 2),
// generator.py:112,121 (nearest x2), generator.py:140 / resnet.py:82 (spatial mean).
#include "common.cuh"

namespace {

// 3x3 stride-2 pad-1 max pool per frame; idx = argmax position 0..8 (first max wins, like ATen).
template <class T>
__global__ void maxpool3x3s2_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, unsigned char* __restrict__ idx, int NF,
                                        int H, int W, int C, int OH, int OW) {
  long long total = (long long)NF * OH * OW * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % C); long long r = i / C;
    int ow = (int)(r % OW); r /= OW; int oh = (int)(r % OH); int n = (int)(r / OH);
    float best = -INFINITY; int bi = 0;
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      int h = oh * 2 - 1 + kh;
      if ((unsigned)h >= (unsigned)H) continue;
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        int w = ow * 2 - 1 + kw;
        if ((unsigned)w >= (unsigned)W) continue;
        float v = to_f(x[(((long long)n * H + h) * W + w) * C + c]);
        if (v > best || (v != v && !(best != best))) { best = v; bi = kh * 3 + kw; }
      }
    }
    y[i] = from_f<T>(best);
    idx[i] = (unsigned char)bi;
  }
}
// gather form of the backward: every input pixel looks at the <=4 windows that contain it.
template <class T>
__global__ void maxpool3x3s2_bwd_kernel(const T* __restrict__ dy, const unsigned char* __restrict__ idx, T* __restrict__ dx,
                                        int NF, int H, int W, int C, int OH, int OW) {
  long long total = (long long)NF * H * W * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % C); long long r = i / C;
    int w = (int)(r % W); r /= W; int h = (int)(r % H); int n = (int)(r / H);
    float acc = 0.f;
    int oh_lo = h >> 1, oh_hi = (h + 1) >> 1;  // windows oh cover rows 2oh-1..2oh+1
    int ow_lo = w >> 1, ow_hi = (w + 1) >> 1;
    for (int oh = oh_lo; oh <= oh_hi; ++oh) {
      if (oh >= OH) continue;
      int kh = h - (oh * 2 - 1);
      for (int ow = ow_lo; ow <= ow_hi; ++ow) {
        if (ow >= OW) continue;
        int kw = w - (ow * 2 - 1);
        long long o = (((long long)n * OH + oh) * OW + ow) * C + c;
        if (idx[o] == kh * 3 + kw) acc += to_f(dy[o]);
      }
    }
    dx[i] = from_f<T>(acc);
  }
}

// avg_pool2d(x, 2) (floor): y[oh,ow] = scale * sum of the 2x2 block.  With scale=0.25 it is the forward; with
// scale=1 it is the backward of nearest-x2 upsampling.
template <class T>
__global__ void pool2x2_sum_kernel(const T* __restrict__ x, T* __restrict__ y, int NF, int H, int W, int C, int OH, int OW,
                                   float scale) {
  long long total = (long long)NF * OH * OW * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % C); long long r = i / C;
    int ow = (int)(r % OW); r /= OW; int oh = (int)(r % OH); int n = (int)(r / OH);
    const T* p = x + (((long long)n * H + oh * 2) * W + ow * 2) * C + c;
    float v = to_f(p[0]) + to_f(p[C]) + to_f(p[(long long)W * C]) + to_f(p[(long long)W * C + C]);
    y[i] = from_f<T>(v * scale);
  }
}
// y[h,w] = scale * x[h/2, w/2] when (h/2 < IH && w/2 < IW) else 0; output H x W (>= 2*IH, 2*IW).
// scale=1: nearest x2 upsample; scale=0.25: backward of avg_pool2d(2) (odd trailing row/col get zero).
template <class T>
__global__ void expand2x2_kernel(const T* __restrict__ x, T* __restrict__ y, int NF, int IH, int IW, int C, int H, int W,
                                 float scale) {
  long long total = (long long)NF * H * W * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % C); long long r = i / C;
    int w = (int)(r % W); r /= W; int h = (int)(r % H); int n = (int)(r / H);
    int ih = h >> 1, iw = w >> 1;
    float v = 0.f;
    if (ih < IH && iw < IW) v = scale * to_f(x[(((long long)n * IH + ih) * IW + iw) * C + c]);
    y[i] = from_f<T>(v);
  }
}
// spatial mean: y[n,c] = scale * sum_p x[n,p,c]
template <class T>
__global__ void spatial_sum_kernel(const T* __restrict__ x, T* __restrict__ y, int NF, int P, int C, float scale) {
  long long total = (long long)NF * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % C); int n = (int)(i / C);
    const T* p = x + (long long)n * P * C + c;
    float a = 0.f;
    for (int q = 0; q < P; ++q) a += to_f(p[(long long)q * C]);
    y[i] = from_f<T>(a * scale);
  }
}
// broadcast: y[n,p,c] = scale * x[n,c]
template <class T>
__global__ void spatial_bcast_kernel(const T* __restrict__ x, T* __restrict__ y, int NF, int P, int C, float scale) {
  long long total = (long long)NF * P * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % C); int n = (int)(i / ((long long)P * C));
    y[i] = from_f<T>(scale * to_f(x[(long long)n * C + c]));
  }
}

}  // namespace

#define DISPATCH_T(dtype, CALL_F32, CALL_BF16) \
  do { if ((dtype) == VCA_F32) { CALL_F32; } else { CALL_BF16; } } while (0)

extern "C" {

int vca_maxpool3x3s2_fwd(int dtype, const void* x, void* y, unsigned char* idx, int NF, int H, int W, int C, cudaStream_t s) {
  VCA_CHECK_ARG(x && y && idx && NF > 0 && H > 0 && W > 0 && C > 0);
  int OH = (H + 2 - 3) / 2 + 1, OW = (W + 2 - 3) / 2 + 1;
  unsigned grid = vca_grid_1d((long long)NF * OH * OW * C, 256, 2);
  DISPATCH_T(dtype, (maxpool3x3s2_fwd_kernel<float><<<grid, 256, 0, s>>>((const float*)x, (float*)y, idx, NF, H, W, C, OH, OW)),
             (maxpool3x3s2_fwd_kernel<bf16><<<grid, 256, 0, s>>>((const bf16*)x, (bf16*)y, idx, NF, H, W, C, OH, OW)));
  VCA_LAUNCH_CHECK();
  return VCA_OK;
}
int vca_maxpool3x3s2_bwd(int dtype, const void* dy, const unsigned char* idx, void* dx, int NF, int H, int W, int C,
                         cudaStream_t s) {
  VCA_CHECK_ARG(dy && dx && idx && NF > 0 && H > 0 && W > 0 && C > 0);
  int OH = (H + 2 - 3) / 2 + 1, OW = (W + 2 - 3) / 2 + 1;
  unsigned grid = vca_grid_1d((long long)NF * H * W * C, 256, 2);
  DISPATCH_T(dtype, (maxpool3x3s2_bwd_kernel<float><<<grid, 256, 0, s>>>((const float*)dy, idx, (float*)dx, NF, H, W, C, OH, OW)),
             (maxpool3x3s2_bwd_kernel<bf16><<<grid, 256, 0, s>>>((const bf16*)dy, idx, (bf16*)dx, NF, H, W, C, OH, OW)));
  VCA_LAUNCH_CHECK();
  return VCA_OK;
}
// y[NF, H/2, W/2, C] = scale * (2x2 block sums of x[NF,H,W,C])
int vca_pool2x2_sum(int dtype, const void* x, void* y, int NF, int H, int W, int C, float scale, cudaStream_t s) {
  VCA_CHECK_ARG(x && y && NF > 0 && H >= 2 && W >= 2 && C > 0);
  int OH = H / 2, OW = W / 2;
  unsigned grid = vca_grid_1d((long long)NF * OH * OW * C, 256, 2);
  DISPATCH_T(dtype, (pool2x2_sum_kernel<float><<<grid, 256, 0, s>>>((const float*)x, (float*)y, NF, H, W, C, OH, OW, scale)),
             (pool2x2_sum_kernel<bf16><<<grid, 256, 0, s>>>((const bf16*)x, (bf16*)y, NF, H, W, C, OH, OW, scale)));
  VCA_LAUNCH_CHECK();
  return VCA_OK;
}
// y[NF,H,W,C] = scale * x[NF,IH,IW,C] replicated 2x2 (zero where h/2>=IH or w/2>=IW)
int vca_expand2x2(int dtype, const void* x, void* y, int NF, int IH, int IW, int C, int H, int W, float scale,
                  cudaStream_t s) {
  VCA_CHECK_ARG(x && y && NF > 0 && IH > 0 && IW > 0 && C > 0 && H >= 2 * IH && W >= 2 * IW && H <= 2 * IH + 1 &&
                W <= 2 * IW + 1);
  unsigned grid = vca_grid_1d((long long)NF * H * W * C, 256, 2);
  DISPATCH_T(dtype, (expand2x2_kernel<float><<<grid, 256, 0, s>>>((const float*)x, (float*)y, NF, IH, IW, C, H, W, scale)),
             (expand2x2_kernel<bf16><<<grid, 256, 0, s>>>((const bf16*)x, (bf16*)y, NF, IH, IW, C, H, W, scale)));
  VCA_LAUNCH_CHECK();
  return VCA_OK;
}
int vca_spatial_sum(int dtype, const void* x, void* y, int NF, int P, int C, float scale, cudaStream_t s) {
  VCA_CHECK_ARG(x && y && NF > 0 && P > 0 && C > 0);
  unsigned grid = vca_grid_1d((long long)NF * C, 128);
  DISPATCH_T(dtype, (spatial_sum_kernel<float><<<grid, 128, 0, s>>>((const float*)x, (float*)y, NF, P, C, scale)),
             (spatial_sum_kernel<bf16><<<grid, 128, 0, s>>>((const bf16*)x, (bf16*)y, NF, P, C, scale)));
  VCA_LAUNCH_CHECK();
  return VCA_OK;
}
int vca_spatial_bcast(int dtype, const void* x, void* y, int NF, int P, int C, float scale, cudaStream_t s) {
  VCA_CHECK_ARG(x && y && NF > 0 && P > 0 && C > 0);
  unsigned grid = vca_grid_1d((long long)NF * P * C, 256, 2);
  DISPATCH_T(dtype, (spatial_bcast_kernel<float><<<grid, 256, 0, s>>>((const float*)x, (float*)y, NF, P, C, scale)),
             (spatial_bcast_kernel<bf16><<<grid, 256, 0, s>>>((const bf16*)x, (bf16*)y, NF, P, C, scale)));
  VCA_LAUNCH_CHECK();
  return VCA_OK;
}

}  // extern "C"
