// Weights-stationary, halo-resident tcgen05 convolution (forward / dgrad, stride 1) for layers with <= 128 input
// channels -- the 32/64-channel 5x5 layers at 80x300 / 40x150 (generator g2/g3, discriminator stems), the ResNet
// layer-1 3x3s and the (5,1) temporal stem conv.  In conv_tc_fwd_kernel those layers are bound by L2->SM bandwidth:
// the 128-pixel activation window is re-fetched once per filter tap (25x for a 5x5) and the (small) weights once
// per CTA.  Here
//   * each persistent CTA loads ALL taps of its BN output channels into shared memory once (<= 120 KB) and keeps
//     them for every tile it processes,
//   * per output tile (th x tw pixels) ONE TMA box fetches the (th+KH-1) x (tw+KW-1) halo of 64 channels; it is kept
//     in "pitched" pixel order (pitch P = tw+KW-1), so the A operand of filter tap (a,b) is the SAME buffer read
//     from row offset a*P+b -- a plain start-address shift of the SWIZZLE_128B matrix descriptor.  The M rows whose
//     column index is >= tw are garbage and are dropped by the epilogue (th*P <= 128),
//   * the accumulator is double-buffered in TMEM so the epilogue of tile i overlaps the MMAs of tile i+1.
// L2 traffic per output pixel drops from taps*128 B to ~(halo/tile)*128 B (2-3x instead of 25x).
#include "tc_common.cuh"

using namespace tc;

extern int g_wgws_mode;     // conv_tc_wgrad_ws.cu
extern int g_bn_vec;        // bn.cu
extern int g_gru_cluster, g_gru_bs;   // gru_cluster.cu
extern int g_hs_mode;       // conv_tc_hs.cu
extern int g_wgws_waves;    // conv_tc_wgrad_ws.cu
extern int g_gl_fpw;        // stft.cu

namespace {

constexpr int KC = 64;
int g_ws_mode = 1;       // 0 off, 1 auto, 2 force whenever the geometry fits
int g_ws_base_off = 0;    // descriptor base-offset mode for shifted A windows (0: none, 1: (addr >> 7) & 7)
int g_ws_dbg = 0;         // probe switches (tools/ws_bound_probe.py): 1 no global stores, 2 no TMEM reads, 4 no MMAs, 8 / 16 MMA N forced to 128 / 256
int g_ws_min_taps = 2;    // smallest filter the kernel takes (1: also pointwise convs)

struct WsParams {
  int NF, OH, OW, Cout;
  int th, tw, P;
  int tiles_w, tiles_h, num_tiles;
  int KH, KW, ph, pw, flip;
  int kchunks, ksteps_last;
  int BN;
  int sa;
  uint32_t a_stage_bytes, a_tx_bytes, w_bytes, tmem_cols;
  int base_off_mode, dbg;
  const float* bias;
  double* stats;           // [2 * Cout] BatchNorm sum / sum-of-squares accumulators (fp64, added to) or null
  EpiExtra ex;             // inference epilogue (scale / residual / activation); has_ex = 0: plain bias epilogue
  int has_ex;
  bf16* y;
};

// MODE 0: plain (+bias) epilogue; 1: + BatchNorm statistics; 2: inference epilogue (scale / shift / residual / activation).
// A template so that each variant gets its own register allocation (the statistics keep 128 running sums per thread, the
// inference epilogue prefetches the residual row): none of it is paid for by the plain forward / dgrad launches.
template <int MODE>
__global__ void __launch_bounds__(192, 1) conv_tc_ws_kernel(const __grid_constant__ CUtensorMap tmA,
                                                            const __grid_constant__ CUtensorMap tmB, const WsParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int taps = p.KH * p.KW;
  const uint32_t w_tile = (uint32_t)p.BN * 128u;
  uint8_t* sW = smem;
  uint8_t* sA = smem + p.w_bytes;
  uint64_t* a_full = (uint64_t*)(sA + (size_t)p.sa * p.a_stage_bytes);
  uint64_t* a_empty = a_full + p.sa;
  uint64_t* w_bar = a_empty + p.sa;
  uint64_t* t_full = w_bar + 1;     // [2]
  uint64_t* t_empty = t_full + 2;   // [2]
  uint32_t* tmem_slot = (uint32_t*)(t_empty + 2);
  uint32_t* s_arel = tmem_slot + 4;   // [taps <= 256] per-tap row shift of the A window, in 16-byte units
  float* s_sum = (float*)(s_arel + 256);   // [BN] + [BN]: BatchNorm statistics of the current tile
  float* s_sq = s_sum + p.BN;
  if (MODE == 1) for (int i = threadIdx.x; i < 2 * p.BN; i += blockDim.x) s_sum[i] = 0.f;
  if (MODE == 2) epi_stage(s_sum, p.BN, blockIdx.y * p.BN, p.Cout, p.ex, p.bias);     // [3][BN] epilogue vectors in the same space
  float* s_bias = s_sum + 2 * p.BN;          // the bias of the plain / statistics epilogues, staged once (no global load per chunk)
  if (MODE != 2) for (int i = threadIdx.x; i < p.BN; i += blockDim.x) s_bias[i] = (p.bias && blockIdx.y * p.BN + i < p.Cout) ? p.bias[blockIdx.y * p.BN + i] : 0.f;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int co0 = blockIdx.y * p.BN;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < p.sa; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
    mbar_init(w_bar, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&t_full[i], 1); mbar_init(&t_empty[i], 4); }
    mbar_init_fence();
  }
  if (warp == 1) tmem_alloc(tmem_slot, p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // ---- weights: every (tap, K chunk) tile of my BN output channels, once
      mbar_expect_tx(w_bar, (uint32_t)(taps * p.kchunks) * w_tile);
      for (int t = 0; t < taps; ++t) {
        const int a = t / p.KW, b = t % p.KW;
        const int wtap = p.flip ? (p.KH - 1 - a) * p.KW + (p.KW - 1 - b) : t;
        for (int kc = 0; kc < p.kchunks; ++kc)
          tma_load_3d(sW + (size_t)(t * p.kchunks + kc) * w_tile, &tmB, w_bar, kc * KC, co0, wtap);
      }
      // ---- activation halos, one box per (tile, K chunk)
      int stage = 0; uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        int t = tile;
        const int tw_i = t % p.tiles_w; t /= p.tiles_w;
        const int th_i = t % p.tiles_h; const int n = t / p.tiles_h;
        for (int kc = 0; kc < p.kchunks; ++kc) {
          mbar_wait(&a_empty[stage], phase ^ 1);
          mbar_expect_tx(&a_full[stage], p.a_tx_bytes);
          tma_load_4d(sA + (size_t)stage * p.a_stage_bytes, &tmA, &a_full[stage], kc * KC, tw_i * p.tw - p.pw, th_i * p.th - p.ph, n);
          if (++stage == p.sa) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // One thread issues every tcgen05.mma of this CTA.  With N = 32..64 an MMA occupies the tensor pipe for only
    // 16..32 cycles, so the issue loop itself must be that lean: descriptors are (constant high word, low word =
    // smem address >> 4), the per-tap row shift of the A window comes from a table built once, and there is one
    // mbarrier wait per (tile, K chunk) instead of one per tap.
    for (int t = lane; t < taps; t += 32) s_arel[t] = (uint32_t)((t / p.KW) * p.P + (t % p.KW)) * 8u;   // rows*128 B >> 4
    __syncwarp();
    if (lane == 0) {
      const uint32_t idesc = make_idesc(128, (p.dbg & 8) ? 128 : (p.dbg & 16) ? 256 : p.BN, 0, 0);
      const uint64_t HI = (uint64_t)(64u | (1u << 14) | (2u << 29)) << 32;   // SBO = 1024 B, version 1, SWIZZLE_128B
      const uint32_t w_lo = smem_u32(sW) >> 4, w_step = w_tile >> 4;
      mbar_wait(w_bar, 0);
      int stage = 0; uint32_t phase = 0; int it = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
        const int acc = it & 1;
        mbar_wait(&t_empty[acc], ((uint32_t)(it >> 1) & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * p.BN);
        uint32_t accum = 0;
        for (int kc = 0; kc < p.kchunks; ++kc) {
          mbar_wait(&a_full[stage], phase);
          tc_fence_after();
          const uint32_t a_lo = smem_u32(sA + (size_t)stage * p.a_stage_bytes) >> 4;
          const int ksteps = (kc == p.kchunks - 1) ? p.ksteps_last : 4;
          uint32_t b_lo = w_lo + (uint32_t)kc * w_step;
          const uint32_t b_inc = (uint32_t)p.kchunks * w_step;
          if (p.dbg & 4) {
          } else if (ksteps == 4) {
            for (int t = 0; t < taps; ++t, b_lo += b_inc) {
              const uint32_t al = a_lo + s_arel[t];
              umma_bf16(d_tmem, HI | al, HI | b_lo, idesc, accum); accum = 1;
              umma_bf16(d_tmem, HI | (al + 2), HI | (b_lo + 2), idesc, 1);
              umma_bf16(d_tmem, HI | (al + 4), HI | (b_lo + 4), idesc, 1);
              umma_bf16(d_tmem, HI | (al + 6), HI | (b_lo + 6), idesc, 1);
            }
          } else if (ksteps == 2) {
            for (int t = 0; t < taps; ++t, b_lo += b_inc) {
              const uint32_t al = a_lo + s_arel[t];
              umma_bf16(d_tmem, HI | al, HI | b_lo, idesc, accum); accum = 1;
              umma_bf16(d_tmem, HI | (al + 2), HI | (b_lo + 2), idesc, 1);
            }
          } else {
            for (int t = 0; t < taps; ++t, b_lo += b_inc) {
              const uint32_t al = a_lo + s_arel[t];
              for (int k = 0; k < ksteps; ++k) { umma_bf16(d_tmem, HI | (al + 2 * k), HI | (b_lo + 2 * k), idesc, accum); accum = 1; }
            }
          }
          umma_commit(&a_empty[stage]);
          if (++stage == p.sa) { stage = 0; phase ^= 1; }
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
    // BatchNorm statistics (BN <= 64 whenever they are requested): every thread keeps running sums of ITS tile row's
    // 64 columns over all tiles of this persistent CTA -- two FMAs per value; the cross-row reduction happens once, below
    constexpr int NR = MODE == 1 ? 64 : 1;
    float rs[NR], rq[NR];
#pragma unroll
    for (int i = 0; i < NR; ++i) { rs[i] = 0.f; rq[i] = 0.f; }
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
      int t = tile;
      const int tw_i = t % p.tiles_w; t /= p.tiles_w;
      const int th_i = t % p.tiles_h; const int n = t / p.tiles_h;
      const int oh = th_i * p.th + r, ow = tw_i * p.tw + wq;
      const bool row_ok = r < p.th && wq < p.tw && oh < p.OH && ow < p.OW;
      bf16* yrow = p.y + (((long long)n * p.OH + oh) * p.OW + ow) * p.Cout + co0;
      const int acc = it & 1;
      if (MODE == 2 && p.ex.res && row_ok) epi_prefetch_row(p.ex.res + (yrow - p.y), min(p.BN, p.Cout - co0) * 2);
      mbar_wait(&t_full[acc], (uint32_t)(it >> 1) & 1u);
      tc_fence_after();
      if (MODE == 2) {
        const bf16* rrow = p.ex.res ? p.ex.res + (yrow - p.y) : nullptr;
        for (int c = 0; c < p.BN; c += 64)
          epi_group64(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * p.BN + c), s_sum + c, p.BN, p.ex.res_scale, co0 + c,
                      p.BN - c, p.Cout, yrow + c, rrow ? rrow + c : nullptr, row_ok);
      } else
#pragma unroll
      for (int cc = 0; cc < 16; ++cc) {
        const int c = cc * 16;
        if (c >= p.BN || (p.dbg & 2)) break;
        float v[16];
        tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * p.BN + c), v);
        if (p.bias) {
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] += s_bias[c + i];
        }
        if (MODE == 1 && cc < 4 && row_ok) {
#pragma unroll
          for (int i = 0; i < 16; ++i) { const float t = bf16_round(v[i]); rs[((cc & 3) * 16 + i) % NR] += t; rq[((cc & 3) * 16 + i) % NR] = fmaf(t, t, rq[((cc & 3) * 16 + i) % NR]); }
        }
        if (row_ok && co0 + c < p.Cout && !((p.dbg & 1) && v[0] != 123.456f)) {
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
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&t_empty[acc]);
    }
    // BatchNorm statistics: the shared accumulators collect ALL tiles of this persistent CTA (fp32 over a few thousand
    // rows), published once -- a flush per tile would put tens of thousands of fp64 atomics on each channel's address
    if (MODE == 1) {
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) {
        if (cc * 16 < p.BN) {
          float a[16], b[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) { a[i] = rs[(cc * 16 + i) % NR]; b[i] = rq[(cc * 16 + i) % NR]; }
          const float sa = colsum16(a, lane), sb = colsum16(b, lane);
          const int col = epi_col(lane);
          if (!(lane & 1) && co0 + cc * 16 + col < p.Cout) { atomicAdd(s_sum + cc * 16 + col, sa); atomicAdd(s_sq + cc * 16 + col, sb); }
        }
      }
      epi_stats_flush(s_sum, s_sq, p.BN, co0, p.Cout, p.stats, (int)threadIdx.x - 64);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, p.tmem_cols);
}

// pick the output tile th x tw (pitch P = tw+KW-1, th*P <= 128) with the best useful-row fraction
void choose_ws_tile(int H, int W, int KH, int KW, int& th, int& tw) {
  double best = -1; th = 1; tw = 1;
  for (int w = 1; w <= W && w + KW - 1 <= 128; ++w) {
    const int P = w + KW - 1;
    int h = 128 / P; if (h > H) h = H; if (h < 1) continue;
    const double util = (double)(h * w) / 128.0 * ((double)W / (((W + w - 1) / w) * w)) * ((double)H / (((H + h - 1) / h) * h));
    const double amp = (double)((h + KH - 1) * P) / (double)(h * w);   // halo bytes per useful pixel
    const double score = util / (1.0 + 0.05 * amp);
    if (score > best) { best = score; th = h; tw = w; }
  }
}

}  // namespace

// Tries the weights-stationary kernel.  Returns 1 if it was launched, 0 if the geometry does not fit (caller uses the
// streaming kernel), negative on error.  Arguments as conv_tc.cu::fwd_like.
int conv_ws_try(int NF, int IH, int IW, int Kdim, int OH, int OW, int Nout, int KH, int KW, int ph, int pw, int flip,
                    const void* x, const void* wpk, const float* bias, void* y, double* stats, const tc::EpiExtra* ex, cudaStream_t s) {
  // x == nullptr: dry run -- 1 when this kernel would take the geometry (and, with stats != nullptr, emit the statistics)
  if (g_ws_mode == 0) return 0;
  const int taps = KH * KW, kchunks = (Kdim + KC - 1) / KC;
  if (taps < g_ws_min_taps || taps > 256 || kchunks > 2 || KW > 64) return 0;
  int bn = ((Nout + 15) / 16) * 16; if (bn > 256) bn = 256;
  const size_t W_BUDGET = 120 * 1024;
  while ((size_t)taps * kchunks * bn * 128 > W_BUDGET && bn > 32) bn = (bn / 2 + 15) / 16 * 16;
  if ((size_t)taps * kchunks * bn * 128 > W_BUDGET) return 0;
  const int n_tiles = (Nout + bn - 1) / bn;
  if (g_ws_mode == 1) {
    // auto: only where the streaming kernel is bandwidth bound and re-reading the halo per n-tile stays cheap
    if (Kdim > 64 || n_tiles > 1) return 0;
  }
  WsParams p;
  p.NF = NF; p.OH = OH; p.OW = OW; p.Cout = Nout;
  choose_ws_tile(OH, OW, KH, KW, p.th, p.tw);
  p.P = p.tw + KW - 1;
  p.tiles_w = (OW + p.tw - 1) / p.tw; p.tiles_h = (OH + p.th - 1) / p.th;
  const long long nt = (long long)NF * p.tiles_w * p.tiles_h;
  if (nt > 0x7fffffff) return 0;
  p.num_tiles = (int)nt;
  p.KH = KH; p.KW = KW; p.ph = ph; p.pw = pw; p.flip = flip;
  p.kchunks = kchunks;
  const int last = Kdim - (kchunks - 1) * KC;
  p.ksteps_last = (last + 15) / 16;
  p.BN = bn;
  p.w_bytes = (uint32_t)(taps * kchunks * bn * 128);
  p.a_tx_bytes = (uint32_t)(p.P * (p.th + KH - 1)) * 128u;
  const uint32_t a_need = (uint32_t)(128 + (KH - 1) * p.P + KW) * 128u;      // rows any tap window may touch
  p.a_stage_bytes = ((a_need > p.a_tx_bytes ? a_need : p.a_tx_bytes) + 1023u) & ~1023u;
  int sa = (int)((220 * 1024 - (size_t)p.w_bytes - 4096) / p.a_stage_bytes);
  if (sa > 4) sa = 4;
  if (sa < 2) return 0;
  p.sa = sa;
  p.tmem_cols = g_ws_dbg & 24 ? 512 : pow2_cols(2 * bn);
  if (stats && bn > 64) return 0;          // the epilogue keeps the statistics of at most 64 columns in registers
  if (!x) return 1;
  p.base_off_mode = g_ws_base_off; p.dbg = g_ws_dbg;
  p.bias = bias; p.y = (bf16*)y; p.stats = stats;
  p.has_ex = ex != nullptr;
  if (ex) p.ex = *ex; else p.ex = EpiExtra{nullptr, nullptr, 0.f, 0, 0.f, nullptr};
  const size_t smem = (size_t)p.w_bytes + (size_t)sa * p.a_stage_bytes + 1024 + 5120;   // + alignment + barriers/tables/statistics or epilogue vectors

  CUtensorMap tmA, tmB;
  long long dA[4] = {Kdim, IW, IH, NF}; int bA[4] = {KC, p.P, p.th + KH - 1, 1};
  long long dB[3] = {Kdim, Nout, (long long)taps}; int bB[3] = {KC, bn, 1};
  if (bA[1] > 256 || bA[2] > 256) return 0;
  int rc = make_map(&tmA, x, 4, dA, bA); if (rc) return rc;
  rc = make_map(&tmB, wpk, 3, dB, bB); if (rc) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(conv_tc_ws_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess ||
        cudaFuncSetAttribute(conv_tc_ws_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess ||
        cudaFuncSetAttribute(conv_tc_ws_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) {
      vca_set_error("cudaFuncSetAttribute(conv_tc_ws_kernel) failed"); return VCA_ERR_CUDA;
    }
    attr_set = true;
  }
  int gx = vca_num_sms() / n_tiles; if (gx < 1) gx = 1; if (gx > p.num_tiles) gx = p.num_tiles;
  dim3 grid((unsigned)gx, (unsigned)n_tiles, 1);
  if (p.stats) conv_tc_ws_kernel<1><<<grid, 192, smem, s>>>(tmA, tmB, p);
  else if (p.has_ex) conv_tc_ws_kernel<2><<<grid, 192, smem, s>>>(tmA, tmB, p);
  else conv_tc_ws_kernel<0><<<grid, 192, smem, s>>>(tmA, tmB, p);
  VCA_LAUNCH_CHECK();
  return 1;
}

extern "C" {
// Runtime switches (testing / tuning): "ws_mode" 0 off, 1 auto (default), 2 force; "ws_base_off" 0/1.
int vca_set_option(const char* key, int value) {
  VCA_CHECK_ARG(key);
  const char* k = key;
  auto eq = [&](const char* s) { const char* a = k; while (*a && *s && *a == *s) { ++a; ++s; } return *a == 0 && *s == 0; };
  if (eq("ws_mode")) { g_ws_mode = value; return VCA_OK; }
  if (eq("ws_base_off")) { g_ws_base_off = value; return VCA_OK; }
  if (eq("ws_dbg")) { g_ws_dbg = value; return VCA_OK; }
  if (eq("ws_min_taps")) { g_ws_min_taps = value < 1 ? 1 : value; return VCA_OK; }
  if (eq("wgws_mode")) { g_wgws_mode = value; return VCA_OK; }
  if (eq("bn_vec")) { g_bn_vec = value; return VCA_OK; }
  if (eq("gru_cluster")) { g_gru_cluster = value; return VCA_OK; }
  if (eq("gru_bs")) { g_gru_bs = value; return VCA_OK; }
  if (eq("hs_mode")) { g_hs_mode = value; return VCA_OK; }
  if (eq("wgws_waves")) { g_wgws_waves = value < 1 ? 1 : value; return VCA_OK; }
  if (eq("gl_fpw")) { g_gl_fpw = value < 1 ? 1 : value; return VCA_OK; }
  vca_set_error("vca_set_option: unknown key %s", key);
  return VCA_ERR_ARG;
}
}
