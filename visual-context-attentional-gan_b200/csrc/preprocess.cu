// Clip preprocessing of the data loader on the device:
//   MultiDataset.build_tensor, src/data/vid_aud_grid.py:94-121 (fixed crop 136x136) and src/data/vid_aud_lrs2.py:87-120
//   (per-frame 80x80 crops around the mouth landmark):
//   ToPILImage -> Crop -> Resize([112,112]) -> [hflip] -> Grayscale -> ToTensor -> Normalize(0.4136, 0.17) per frame,
//   zero frames behind the clip's last frame, one 56x56 random-erasing box per clip.
// The reference does this with PIL on 6 loader processes.  Everything here is the same INTEGER arithmetic, so the
// result is bit-exact:
//   * PIL's bilinear resize of an 8-bit image (Pillow src/libImaging/Resample.c, third-party, un-vendored): a separable
//     "convolution" with support scaled by the shrink factor; coefficients are computed in double on the host exactly
//     as precompute_coeffs/normalize_coeffs_8bpc do (vcagan_b200/preprocess.py) and arrive here as 22-bit fixed point;
//     horizontal pass first, rounded to uint8, then the vertical pass: acc = 2^21 + sum(pixel * k) >> 22, clamped.
//   * Image.convert("L"): (19595 R + 38470 G + 7471 B + 0x8000) >> 16   (Pillow Convert.c, ITU-R 601-2 luma).
//   * ToTensor / Normalize: ((g / 255) - mean) / std in fp32 with IEEE division, as torch does on the CPU.
//   * Crop boxes reaching outside the frame read zeros (PIL pads a crop with black).
// One CTA per frame; the uint8 intermediate of the horizontal pass (crop_h x 112 x 3 bytes, 45 KB for GRID) lives in
// shared memory, so a frame is read once from HBM (only its crop window) and written once as fp32.
#include "common.cuh"

namespace {

constexpr int PP_THREADS = 256, PP_PREC = 22, PP_META = 10;   // meta: l,u,r,b, flip, ex0,ey0,ex1,ey1, valid

__device__ __forceinline__ int clip8(int acc) {
  const int v = acc >> PP_PREC;
  return v < 0 ? 0 : (v > 255 ? 255 : v);
}

// Generic fall-back (any tap counts, thread per output): used for non-square crops or more than 7 taps.
__global__ void __launch_bounds__(PP_THREADS)
clip_preprocess_generic_kernel(const unsigned char* __restrict__ frames, int H, int W, const int* __restrict__ meta,
                       const int* __restrict__ kx, const int* __restrict__ bx, const int* __restrict__ ky,
                       const int* __restrict__ by, int ksx, int ksy, int cw, int ch, int OW, int OH,
                       float mean, float stdv, float* __restrict__ out) {
  extern __shared__ unsigned char tmp[];   // [ch][OW][3]
  const int n = blockIdx.x;
  const int* m = meta + (size_t)n * PP_META;
  float* dst = out + (size_t)n * OH * OW;
  if (!m[9]) {                              // behind the clip's last frame: temporalVolume stays zero
    for (int i = threadIdx.x; i < OH * OW; i += PP_THREADS) dst[i] = 0.f;
    return;
  }
  const int l = m[0], u = m[1], flip = m[4], ex0 = m[5], ey0 = m[6], ex1 = m[7], ey1 = m[8];
  const unsigned char* src = frames + (size_t)n * H * W * 3;

  for (int i = threadIdx.x; i < ch * OW; i += PP_THREADS) {          // horizontal pass
    const int y = i / OW, xx = i - y * OW;
    const int fy = u + y;
    int a0 = 1 << (PP_PREC - 1), a1 = a0, a2 = a0;
    if (fy >= 0 && fy < H) {
      const int x0 = l + bx[2 * xx], cnt = bx[2 * xx + 1];
      const int* k = kx + xx * ksx;
      const unsigned char* row = src + (size_t)fy * W * 3;
      for (int j = 0; j < cnt; ++j) {
        const int fx = x0 + j;
        if (fx >= 0 && fx < W) {
          const int c = k[j];
          a0 += row[fx * 3 + 0] * c; a1 += row[fx * 3 + 1] * c; a2 += row[fx * 3 + 2] * c;
        }
      }
    }
    tmp[i * 3 + 0] = (unsigned char)clip8(a0);
    tmp[i * 3 + 1] = (unsigned char)clip8(a1);
    tmp[i * 3 + 2] = (unsigned char)clip8(a2);
  }
  __syncthreads();

  for (int i = threadIdx.x; i < OH * OW; i += PP_THREADS) {          // vertical pass, luma, normalise, flip, erase
    const int yy = i / OW, xx = i - yy * OW;
    const int y0 = by[2 * yy], cnt = by[2 * yy + 1];
    const int* k = ky + yy * ksy;
    int a0 = 1 << (PP_PREC - 1), a1 = a0, a2 = a0;
    for (int j = 0; j < cnt; ++j) {
      const unsigned char* p = tmp + ((y0 + j) * OW + xx) * 3;
      const int c = k[j];
      a0 += p[0] * c; a1 += p[1] * c; a2 += p[2] * c;
    }
    const int g = (clip8(a0) * 19595 + clip8(a1) * 38470 + clip8(a2) * 7471 + 0x8000) >> 16;
    float v = __fdiv_rn(__fsub_rn(__fdiv_rn((float)g, 255.f), mean), stdv);
    const int ox = flip ? OW - 1 - xx : xx;
    if (ox >= ex0 && ox < ex1 && yy >= ey0 && yy < ey1) v = 0.f;
    dst[yy * OW + ox] = v;
  }
}

// The kernel the path uses (tap counts 3, 5 or 7 per axis = enlarging, up to 2x and up to 3x shrinking).  Same
// arithmetic as the generic one, organised so that the inner loops carry no index arithmetic and no branches:
//   * a lane owns an output COLUMN xx for the whole frame: its horizontal taps (coefficients, clamped byte offsets,
//     zeroed where the window leaves the frame) are loaded once into registers; a warp walks the rows, so the row
//     base and the vertical taps are warp-uniform;
//   * taps beyond a window's real count carry coefficient 0 and a clamped address, so every window is KS taps long;
//   * ToTensor + Normalize is a 256-entry table built per CTA with the same two IEEE divisions (g is a byte).
template <int KS>
__global__ void __launch_bounds__(PP_THREADS)
clip_preprocess_kernel(const unsigned char* __restrict__ frames, int H, int W, const int* __restrict__ meta,
                       const int* __restrict__ kx, const int* __restrict__ bx, const int* __restrict__ ky,
                       const int* __restrict__ by, int cw, int ch, int OW, int OH, float mean, float stdv,
                       float* __restrict__ out) {
  extern __shared__ unsigned char tmp[];   // [ch][OW][3]
  __shared__ float lut[256];
  const int n = blockIdx.x;
  const int* m = meta + (size_t)n * PP_META;
  float* dst = out + (size_t)n * OH * OW;
  if (!m[9]) {                              // behind the clip's last frame: temporalVolume stays zero
    for (int i = threadIdx.x; i < OH * OW; i += PP_THREADS) dst[i] = 0.f;
    return;
  }
  lut[threadIdx.x] = __fdiv_rn(__fsub_rn(__fdiv_rn((float)threadIdx.x, 255.f), mean), stdv);   // PP_THREADS == 256
  const int l = m[0], u = m[1], flip = m[4], ex0 = m[5], ey0 = m[6], ex1 = m[7], ey1 = m[8];
  const unsigned char* src = frames + (size_t)n * H * W * 3;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int NW = PP_THREADS / 32, HALF = 1 << (PP_PREC - 1);

  for (int c0 = 0; c0 < OW; c0 += 32) {                                // horizontal pass
    const int xx = c0 + lane;
    const bool act = xx < OW;
    int k[KS], off[KS];
#pragma unroll
    for (int j = 0; j < KS; ++j) { k[j] = 0; off[j] = 0; }
    if (act) {
      const int x0 = l + bx[2 * xx], cnt = bx[2 * xx + 1];
#pragma unroll
      for (int j = 0; j < KS; ++j) {
        const int fx = x0 + j;
        const bool in = j < cnt && fx >= 0 && fx < W;
        k[j] = in ? kx[xx * KS + j] : 0;
        off[j] = in ? fx * 3 : 0;
      }
    }
    for (int y = warp; y < ch; y += NW) {
      const int fy = u + y;                                            // warp-uniform
      int a0 = HALF, a1 = HALF, a2 = HALF;
      if (fy >= 0 && fy < H) {
        const unsigned char* row = src + fy * W * 3;
#pragma unroll
        for (int j = 0; j < KS; ++j) {
          const unsigned char* p = row + off[j];
          a0 += p[0] * k[j]; a1 += p[1] * k[j]; a2 += p[2] * k[j];
        }
      }
      if (act) {
        unsigned char* t = tmp + (y * OW + xx) * 3;
        t[0] = (unsigned char)clip8(a0); t[1] = (unsigned char)clip8(a1); t[2] = (unsigned char)clip8(a2);
      }
    }
  }
  __syncthreads();

  for (int c0 = 0; c0 < OW; c0 += 32) {                                // vertical pass, luma, normalise, flip, erase
    const int xx = c0 + lane;
    if (xx >= OW) continue;
    const int ox = flip ? OW - 1 - xx : xx;
    const bool ecol = ox >= ex0 && ox < ex1;
    const unsigned char* col = tmp + xx * 3;
    for (int yy = warp; yy < OH; yy += NW) {                           // warp-uniform row: uniform taps
      const int y0 = by[2 * yy], cnt = by[2 * yy + 1];
      int a0 = HALF, a1 = HALF, a2 = HALF;
#pragma unroll
      for (int j = 0; j < KS; ++j) {
        const int c = j < cnt ? ky[yy * KS + j] : 0;
        const unsigned char* p = col + min(y0 + j, ch - 1) * (OW * 3);
        a0 += p[0] * c; a1 += p[1] * c; a2 += p[2] * c;
      }
      const int g = (clip8(a0) * 19595 + clip8(a1) * 38470 + clip8(a2) * 7471 + 0x8000) >> 16;
      dst[yy * OW + ox] = (ecol && yy >= ey0 && yy < ey1) ? 0.f : lut[g];
    }
  }
}

}  // namespace

extern "C" {

int vca_clip_preprocess(const unsigned char* frames, int n_frames, int H, int W, const int* meta, const int* kx,
                        const int* bx, const int* ky, const int* by, int ksx, int ksy, int crop_w, int crop_h,
                        int OW, int OH, float mean, float stdv, float* out, cudaStream_t s) {
  VCA_CHECK_ARG(frames && meta && kx && bx && ky && by && out);
  VCA_CHECK_ARG(n_frames > 0 && H > 0 && W > 0 && crop_w > 0 && crop_h > 0 && OW > 0 && OH > 0 && ksx > 0 && ksy > 0);
  VCA_CHECK_ARG(stdv != 0.f);
  const size_t smem = (size_t)crop_h * OW * 3;
  VCA_CHECK_ARG(smem <= 200 * 1024);
  if (smem > 48 * 1024) {   // per-device attribute: set on every such call (crops taller than 146 rows only)
    cudaError_t e = cudaFuncSetAttribute(clip_preprocess_generic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      vca_set_error("%s:%d: cudaFuncSetAttribute failed: %s", __FILE__, __LINE__, cudaGetErrorString(e));
      return VCA_ERR_CUDA;
    }
  }
  const bool fast = ksx == ksy && (ksx == 3 || ksx == 5 || ksx == 7);
#define VCA_PP_LAUNCH(KS)                                                                                            \
  do {                                                                                                               \
    if (smem > 48 * 1024) cudaFuncSetAttribute(clip_preprocess_kernel<KS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    clip_preprocess_kernel<KS><<<n_frames, PP_THREADS, smem, s>>>(frames, H, W, meta, kx, bx, ky, by, crop_w, crop_h, OW, OH, \
                                                                  mean, stdv, out);                                  \
  } while (0)
  if (fast && ksx == 3) VCA_PP_LAUNCH(3);
  else if (fast && ksx == 5) VCA_PP_LAUNCH(5);
  else if (fast) VCA_PP_LAUNCH(7);
  else
    clip_preprocess_generic_kernel<<<n_frames, PP_THREADS, smem, s>>>(frames, H, W, meta, kx, bx, ky, by, ksx, ksy, crop_w,
                                                                      crop_h, OW, OH, mean, stdv, out);
#undef VCA_PP_LAUNCH
  VCA_LAUNCH_CHECK();
  return VCA_OK;
}

}  // extern "C"
