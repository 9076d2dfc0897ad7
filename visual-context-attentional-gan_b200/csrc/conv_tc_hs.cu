// Halo-resident activations + streamed weights: tcgen05 convolution (forward / dgrad, stride 1) for the layers whose
// weights do not fit in shared memory but whose streaming form (conv_tc_fwd_kernel) is bound by L2 -> SM traffic --
// 64-channel 5x5 (205 KB of weights), 128-channel 5x5 / 3x3, 256-channel 5x5.  conv_tc_fwd_kernel moves 32 KB per 4
// MMAs (the activation window is re-fetched for every filter tap and the weight slab for every 128-pixel tile); ncu
// shows 614 MB of L2 reads for a 39 GFLOP layer and the tensor pipe waiting on TMA.  Here, per persistent CTA:
//   * ONE TMA box per (work item, 64-channel chunk) fetches the halo of MT = 2 vertically stacked output tiles in
//     "pitched" pixel order (pitch P = tw+KW-1), so the A operand of tile mt and filter tap (a,b) is the same buffer
//     read from row offset (mt*th + a)*P + b (a start-address shift of the SWIZZLE_128B descriptor, as in
//     conv_tc_ws_kernel);
//   * the weights are streamed tap by tap through a ring of [BN x 64] tiles, and every tile feeds BOTH output tiles
//     (two TMEM accumulators), which halves the weight traffic per MMA;
//   * accumulators are double-buffered in TMEM (2 x MT x BN <= 512 columns): the epilogue of item i overlaps the
//     MMAs of item i+1; activations and weights have their own producer warps and mbarrier rings.
// L2 traffic per MMA drops ~4x (128-ch 5x5: 3200 KB -> 850 KB per 256 output rows).
#include "tc_common.cuh"

using namespace tc;

int g_hs_mode = 1;   // "hs_mode" in vca_set_option: 0 off, 1 auto, 2 force whenever the geometry fits

namespace {

constexpr int KC = 64;
constexpr int MT = 2;          // output tiles per work item (share every weight tile)
constexpr int NTHREADS = 224;  // warp 0: weight TMA, 1: MMA + TMEM alloc, 2-5: epilogue, 6: activation TMA

struct HsParams {
  int NF, OH, OW, Cout;
  int th, tw, P;
  int tiles_w, groups_h, num_items;
  int KH, KW, ph, pw, flip;
  int kchunks, ksteps_last;
  int BN;
  int sa, sw;
  uint32_t a_stage_bytes, a_tx_bytes, w_tile_bytes, tmem_cols;
  const float* bias;
  double* stats;           // [2 * Cout] BatchNorm sum / sum-of-squares accumulators (fp64, added to) or null
  EpiExtra ex;             // inference epilogue (scale / residual / activation); has_ex = 0: plain bias epilogue
  int has_ex;
  bf16* y;
};

__global__ void __launch_bounds__(NTHREADS, 1) conv_tc_hs_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                 const __grid_constant__ CUtensorMap tmB, const HsParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int taps = p.KH * p.KW;
  uint8_t* sA = smem;
  uint8_t* sW = smem + (size_t)p.sa * p.a_stage_bytes;
  uint64_t* a_full = (uint64_t*)(sW + (size_t)p.sw * p.w_tile_bytes);
  uint64_t* a_empty = a_full + p.sa;
  uint64_t* w_full = a_empty + p.sa;
  uint64_t* w_empty = w_full + p.sw;
  uint64_t* t_full = w_empty + p.sw;   // [2]
  uint64_t* t_empty = t_full + 2;      // [2]
  uint32_t* tmem_slot = (uint32_t*)(t_empty + 2);
  uint32_t* s_arel = tmem_slot + 4;    // [taps <= 256] per-tap row shift of the A window, in 16-byte units
  float* s_sum = (float*)(s_arel + 256);   // [BN] + [BN]: BatchNorm statistics of the current item
  float* s_sq = s_sum + p.BN;
  if (p.stats) for (int i = threadIdx.x; i < 2 * p.BN; i += blockDim.x) s_sum[i] = 0.f;
  if (p.has_ex) epi_stage(s_sum, p.BN, blockIdx.y * p.BN, p.Cout, p.ex, p.bias);     // [3][BN] epilogue vectors in the same space

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int co0 = blockIdx.y * p.BN;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < p.sa; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < p.sw; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&t_full[i], 1); mbar_init(&t_empty[i], 4); }
    mbar_init_fence();
  }
  if (warp == 1) tmem_alloc(tmem_slot, p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 6) {
    // ---- activation halos: one box per (item, K chunk)
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
        int t = item;
        const int tw_i = t % p.tiles_w; t /= p.tiles_w;
        const int hg = t % p.groups_h; const int n = t / p.groups_h;
        for (int kc = 0; kc < p.kchunks; ++kc) {
          mbar_wait(&a_empty[stage], phase ^ 1);
          mbar_expect_tx(&a_full[stage], p.a_tx_bytes);
          tma_load_4d(sA + (size_t)stage * p.a_stage_bytes, &tmA, &a_full[stage], kc * KC, tw_i * p.tw - p.pw,
                      hg * MT * p.th - p.ph, n);
          if (++stage == p.sa) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 0) {
    // ---- weights: the (K chunk, tap) tiles of my BN output channels, streamed once per item
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
        for (int kc = 0; kc < p.kchunks; ++kc) {
          for (int t = 0; t < taps; ++t) {
            const int a = t / p.KW, b = t - a * p.KW;
            const int wtap = p.flip ? (p.KH - 1 - a) * p.KW + (p.KW - 1 - b) : t;
            mbar_wait(&w_empty[stage], phase ^ 1);
            mbar_expect_tx(&w_full[stage], p.w_tile_bytes);
            tma_load_3d(sW + (size_t)stage * p.w_tile_bytes, &tmB, &w_full[stage], kc * KC, co0, wtap);
            if (++stage == p.sw) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    for (int t = lane; t < taps; t += 32) s_arel[t] = (uint32_t)((t / p.KW) * p.P + (t % p.KW)) * 8u;   // rows*128 B >> 4
    __syncwarp();
    if (lane == 0) {
      const uint32_t idesc = make_idesc(128, p.BN, 0, 0);
      const uint64_t HI = (uint64_t)(64u | (1u << 14) | (2u << 29)) << 32;   // SBO = 1024 B, version 1, SWIZZLE_128B
      const uint32_t mt_step = (uint32_t)(p.th * p.P) * 8u;                  // rows of one output tile, in 16-byte units
      int as = 0, ws = 0; uint32_t aphase = 0, wphase = 0; int it = 0;
      for (int item = blockIdx.x; item < p.num_items; item += gridDim.x, ++it) {
        const int acc = it & 1;
        mbar_wait(&t_empty[acc], ((uint32_t)(it >> 1) & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t d0 = tmem_base + (uint32_t)((acc * MT + 0) * p.BN);
        const uint32_t d1 = tmem_base + (uint32_t)((acc * MT + 1) * p.BN);
        uint32_t accum = 0;
        for (int kc = 0; kc < p.kchunks; ++kc) {
          mbar_wait(&a_full[as], aphase);
          tc_fence_after();
          const uint32_t a_lo = smem_u32(sA + (size_t)as * p.a_stage_bytes) >> 4;
          const int ksteps = (kc == p.kchunks - 1) ? p.ksteps_last : 4;
          for (int t = 0; t < taps; ++t) {
            mbar_wait(&w_full[ws], wphase);
            tc_fence_after();
            const uint32_t b_lo = smem_u32(sW + (size_t)ws * p.w_tile_bytes) >> 4;
            const uint32_t al0 = a_lo + s_arel[t], al1 = al0 + mt_step;
            for (int k = 0; k < ksteps; ++k) {
              umma_bf16(d0, HI | (al0 + 2 * k), HI | (b_lo + 2 * k), idesc, accum | (uint32_t)k);
              umma_bf16(d1, HI | (al1 + 2 * k), HI | (b_lo + 2 * k), idesc, accum | (uint32_t)k);
            }
            accum = 1;
            umma_commit(&w_empty[ws]);
            if (++ws == p.sw) { ws = 0; wphase ^= 1; }
          }
          umma_commit(&a_empty[as]);
          if (++as == p.sa) { as = 0; aphase ^= 1; }
        }
        umma_commit(&t_full[acc]);
      }
    }
    __syncwarp();
  } else {
    const int q = warp & 3;
    const int m = q * 32 + lane;
    const int r = m / p.P, wq = m - r * p.P;
    int it = 0;
    for (int item = blockIdx.x; item < p.num_items; item += gridDim.x, ++it) {
      int t = item;
      const int tw_i = t % p.tiles_w; t /= p.tiles_w;
      const int hg = t % p.groups_h; const int n = t / p.groups_h;
      const int acc = it & 1;
      if (p.has_ex && p.ex.res) {
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          const int oh = (hg * MT + mt) * p.th + r, ow = tw_i * p.tw + wq;
          if (r < p.th && wq < p.tw && oh < p.OH && ow < p.OW)
            epi_prefetch_row(p.ex.res + (((long long)n * p.OH + oh) * p.OW + ow) * p.Cout + co0, min(p.BN, p.Cout - co0) * 2);
        }
      }
      mbar_wait(&t_full[acc], (uint32_t)(it >> 1) & 1u);
      tc_fence_after();
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        const int oh = (hg * MT + mt) * p.th + r, ow = tw_i * p.tw + wq;
        const bool row_ok = r < p.th && wq < p.tw && oh < p.OH && ow < p.OW;
        bf16* yrow = p.y + (((long long)n * p.OH + oh) * p.OW + ow) * p.Cout + co0;
        if (p.has_ex) {
          const bf16* rrow = p.ex.res ? p.ex.res + (yrow - p.y) : nullptr;
          for (int c = 0; c < p.BN; c += 64)
            epi_group64(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((acc * MT + mt) * p.BN + c), s_sum + c, p.BN, p.ex.res_scale,
                        co0 + c, p.BN - c, p.Cout, yrow + c, rrow ? rrow + c : nullptr, row_ok);
          continue;
        }
        for (int c = 0; c < p.BN; c += 16) {
          float v[16];
          tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((acc * MT + mt) * p.BN + c), v);
          if (p.bias && co0 + c < p.Cout) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] += (co0 + c + i < p.Cout) ? __ldg(p.bias + co0 + c + i) : 0.f;
          }
          if (p.stats) epi_stats16(v, row_ok, co0 + c, p.Cout, s_sum + c, s_sq + c, lane);
          if (row_ok && co0 + c < p.Cout) {
            if (co0 + c + 16 <= p.Cout) {
              uint32_t w[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) { __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]); w[i] = *reinterpret_cast<uint32_t*>(&h); }
              *reinterpret_cast<uint4*>(yrow + c) = make_uint4(w[0], w[1], w[2], w[3]);
              *reinterpret_cast<uint4*>(yrow + c + 8) = make_uint4(w[4], w[5], w[6], w[7]);
            } else {
              for (int i = 0; i < 16 && co0 + c + i < p.Cout; ++i) yrow[c + i] = __float2bfloat16_rn(v[i]);
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&t_empty[acc]);
    }
    // BatchNorm statistics: the shared accumulators collect ALL tiles of this persistent CTA (fp32 over a few thousand
    // rows), published once -- a flush per tile would put tens of thousands of fp64 atomics on each channel's address
    if (p.stats) epi_stats_flush(s_sum, s_sq, p.BN, co0, p.Cout, p.stats, (int)threadIdx.x - 64);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, p.tmem_cols);
}

// output tile th x tw (pitch P = tw+KW-1, th*P <= 128), rows taken in groups of MT*th: best useful-row fraction
bool choose_hs_tile(int H, int W, int KH, int KW, int& th, int& tw) {
  double best = -1; th = 0; tw = 0;
  for (int w = 1; w <= W && w + KW - 1 <= 128; ++w) {
    const int P = w + KW - 1;
    for (int h = 1; h * P <= 128 && h <= H; ++h) {
      const int gh = MT * h;
      const double util = (double)(h * w) / 128.0 * ((double)W / (((W + w - 1) / w) * w)) * ((double)H / (((H + gh - 1) / gh) * gh));
      const double amp = (double)((gh + KH - 1) * P) / (double)(gh * w);   // halo bytes per useful pixel
      const double score = util / (1.0 + 0.03 * amp);
      if (score > best) { best = score; th = h; tw = w; }
    }
  }
  return th > 0 && best > 0.45;
}

}  // namespace

// Tries the halo-resident / streamed-weights kernel.  Returns 1 if it was launched, 0 if the geometry does not fit or
// the streaming kernel is the better choice, negative on error.  Arguments as conv_tc.cu::fwd_like.
int conv_hs_try(int NF, int IH, int IW, int Kdim, int OH, int OW, int Nout, int KH, int KW, int ph, int pw, int flip,
                const void* x, const void* wpk, const float* bias, void* y, double* stats, const tc::EpiExtra* ex, cudaStream_t s) {
  if (g_hs_mode == 0) return 0;
  const int taps = KH * KW, kchunks = (Kdim + KC - 1) / KC;
  if (taps < 4 || taps > 256 || KW > 32) return 0;
  int bn = ((Nout + 15) / 16) * 16; if (bn > 128) bn = 128;
  const int n_tiles = (Nout + bn - 1) / bn;
  if (g_hs_mode == 1) {
    // auto: only where it measured faster than conv_tc_fwd_kernel (128-channel 3x3: 664 -> 736 TF/s).  Measured: the L2
    // traffic does drop 4x, yet the 5x5 layers are no faster (581 vs 582 TF/s at 128 ch, 384 vs 413 at 64 ch) and a
    // deeper weight ring changes nothing -- so at N <= 128 the bound is on the SM side (an SS-mode MMA with N = 128
    // reads 4 KB of A and 4 KB of B per 64 cycles), not in L2.  Next step for these layers: cta_group::2 tiles.
    if (Nout > 128 || Kdim < 128 || taps > 9) return 0;
  }
  HsParams p;
  p.NF = NF; p.OH = OH; p.OW = OW; p.Cout = Nout;
  if (!choose_hs_tile(OH, OW, KH, KW, p.th, p.tw)) return 0;
  p.P = p.tw + KW - 1;
  p.tiles_w = (OW + p.tw - 1) / p.tw; p.groups_h = (OH + MT * p.th - 1) / (MT * p.th);
  const long long ni = (long long)NF * p.tiles_w * p.groups_h;
  if (ni > 0x7fffffff) return 0;
  p.num_items = (int)ni;
  p.KH = KH; p.KW = KW; p.ph = ph; p.pw = pw; p.flip = flip;
  p.kchunks = kchunks;
  p.ksteps_last = (Kdim - (kchunks - 1) * KC + 15) / 16;
  p.BN = bn;
  const int box_rows = MT * p.th + KH - 1;
  p.a_tx_bytes = (uint32_t)(p.P * box_rows) * 128u;
  const uint32_t a_need = (uint32_t)((MT - 1) * p.th * p.P + (KH - 1) * p.P + KW + 127) * 128u;   // rows any tap window may touch
  p.a_stage_bytes = ((a_need > p.a_tx_bytes ? a_need : p.a_tx_bytes) + 1023u) & ~1023u;
  p.w_tile_bytes = (uint32_t)bn * 128u;
  const size_t budget = 219 * 1024 - 4096;
  // Two halo stages are enough (the next box is issued half an item / a whole item ahead); everything else goes to
  // the weight ring: a tile is consumed every 8 MMAs (~512 cycles) and a TMA round trip under load is 3-4k cycles.
  const int sa = 2;
  if ((size_t)sa * p.a_stage_bytes + 4 * (size_t)p.w_tile_bytes > budget) return 0;
  int sw = (int)((budget - (size_t)sa * p.a_stage_bytes) / p.w_tile_bytes);
  if (sw > 16) sw = 16;
  p.sa = sa; p.sw = sw;
  p.tmem_cols = pow2_cols(2 * MT * bn);
  if (p.tmem_cols > 512) return 0;
  p.bias = bias; p.y = (bf16*)y; p.stats = stats;
  p.has_ex = ex != nullptr;
  if (ex) p.ex = *ex; else p.ex = EpiExtra{nullptr, nullptr, 0.f, 0, 0.f, nullptr};
  const size_t smem = (size_t)sa * p.a_stage_bytes + (size_t)sw * p.w_tile_bytes + 1024 + 5120;   // + alignment + barriers/tables/statistics or epilogue vectors

  CUtensorMap tmA, tmB;
  long long dA[4] = {Kdim, IW, IH, NF}; int bA[4] = {KC, p.P, box_rows, 1};
  long long dB[3] = {Kdim, Nout, (long long)taps}; int bB[3] = {KC, bn, 1};
  if (bA[1] > 256 || bA[2] > 256) return 0;
  int rc = make_map(&tmA, x, 4, dA, bA); if (rc) return rc;
  rc = make_map(&tmB, wpk, 3, dB, bB); if (rc) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(conv_tc_hs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) {
      vca_set_error("cudaFuncSetAttribute(conv_tc_hs_kernel) failed"); return VCA_ERR_CUDA;
    }
    attr_set = true;
  }
  int gx = vca_num_sms() / n_tiles; if (gx < 1) gx = 1; if (gx > p.num_items) gx = p.num_items;
  dim3 grid((unsigned)gx, (unsigned)n_tiles, 1);
  conv_tc_hs_kernel<<<grid, NTHREADS, smem, s>>>(tmA, tmB, p);
  VCA_LAUNCH_CHECK();
  return 1;
}
