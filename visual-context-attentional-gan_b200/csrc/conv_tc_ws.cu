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
//
// Round 2, from the probe of tools/ws_bound_probe.sh (profiles/ws_bound_probe_r02.txt): with <= 128 columns an SS-mode
// tcgen05.mma 128 x N x 16 takes ~100 cycles whatever N is (its A slice is re-read from shared memory), 162 cycles at
// N = 256, and the row-per-thread epilogue (32 partial lines per store instruction) took as long as the MMAs.  Hence
//   * FILTER-ROW STACKING (SK = KW > 1): the KW taps of one filter row share ONE A window.  Their weight tiles lie
//     back to back in shared memory, so a single MMA with N = KW * BN columns computes, for window shift (kh, 0),
//         D_j[p] = sum_ci X[p + (kh, 0)] . W[kh, j]      j = 0 .. KW-1
//     and the convolution is out[q] = sum_j D_j[q + j]: pixel q + j is lane + j of the same warp when the pitch P
//     divides 32, so the epilogue combines the KW column groups with warp shuffles.  9 (25) A-slice reads per K step
//     become 3 (5).
//   * TMA-STORE EPILOGUE: the finished bf16 tile is staged in shared memory (swizzled, dense th x tw rows) and written
//     by one cp.async.bulk.tensor store per tile (full 128-byte lines, clipped at the image edge by the TMA unit),
//     double-buffered so the store of tile i drains underneath the epilogue of tile i+1.
#include "tc_common.cuh"

using namespace tc;

extern int g_wgws_mode;     // conv_tc_wgrad_ws.cu
extern int g_bn_vec;        // bn.cu
extern int g_gru_cluster, g_gru_bs;   // gru_cluster.cu
extern int g_hs_mode;       // conv_tc_hs.cu
extern int g_wgws_waves;    // conv_tc_wgrad_ws.cu
extern int g_wgws_mstack, g_wg_dbg;
extern int g_fwd_smem_kb, g_wg_smem_kb;   // conv_tc.cu
extern int g_im2col_rb;                    // pool.cu
extern int g_gl_fpw;        // stft.cu

namespace {

constexpr int KC = 64;
constexpr int WS_EPI_WARPS = 8;                       // two per TMEM lane quarter: even / odd 16-column chunks
constexpr int WS_THREADS = 64 + 32 * WS_EPI_WARPS;    // warp 0 TMA, warp 1 MMA, then the epilogue warps
int g_ws_mode = 1;       // 0 off, 1 auto, 2 force whenever the geometry fits
int g_ws_base_off = 0;    // descriptor base-offset mode for shifted A windows (0: none, 1: (addr >> 7) & 7)
int g_ws_dbg = 0;         // probe switches (tools/ws_bound_probe.py): 1 no global stores, 2 no TMEM reads, 4 no MMAs, 8 / 16 MMA N forced to 128 / 256
int g_ws_min_taps = 1;    // smallest filter the kernel takes (1: also pointwise convs)
int g_ws_wbudget_kb = 148;  // shared memory the stationary weights may take
int g_ws_kmax = 64;       // auto mode: largest input-channel count without filter-row stacking (128 with)
int g_ws_stack = 1;       // filter-row stacking (KW taps per MMA) where the geometry allows
int g_ws_tma_out = 1;     // epilogue through shared memory + TMA store

struct WsParams {
  int NF, OH, OW, Cout;
  int th, tw, P;
  int tiles_w, tiles_h, num_tiles;
  int KH, KW, ph, pw, flip;
  int kchunks, ksteps_last;
  int BN;
  int SK;                  // taps stacked along N per MMA (KW with filter-row stacking, else 1)
  int tma_out;             // 1: epilogue through the shared-memory staging tile + TMA store
  int sa;
  uint32_t a_stage_bytes, a_tx_bytes, w_bytes, tmem_cols, out_bytes;
  int base_off_mode, dbg;
  const float* bias;
  const float* slope;      // inference: per-channel negative-side slope applied after the bias (LeakyReLU / PReLU / ReLU) or null
  double* stats;           // [2 * Cout] BatchNorm sum / sum-of-squares accumulators (fp64, added to) or null
  EpiExtra ex;             // inference epilogue (scale / residual / activation); has_ex = 0: plain bias epilogue
  int has_ex;
  bf16* y;
};

__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, const void* src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               ::"l"((uint64_t)map), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// MODE 0: plain (+bias) epilogue; 1: + BatchNorm statistics; 2: inference epilogue (scale / shift / residual / activation).
// SK: taps stacked along N (1 = one MMA per tap).  Templates so that each variant gets its own register allocation (the
// statistics keep 128 running sums per thread, the stacked epilogue holds SK x 16 accumulator columns at a time).
template <int MODE, int SK>
__global__ void __launch_bounds__(WS_THREADS, 1) conv_tc_ws_kernel(const __grid_constant__ CUtensorMap tmA,
                                                            const __grid_constant__ CUtensorMap tmB,
                                                            const __grid_constant__ CUtensorMap tmC, const WsParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int taps = p.KH * p.KW;
  const uint32_t w_tile = (uint32_t)p.BN * 128u;
  uint8_t* sW = smem;                                   // [K chunk][tap][BN rows][128 B]
  uint8_t* sA = smem + p.w_bytes;
  uint8_t* sO = sA + (size_t)p.sa * p.a_stage_bytes;    // [2][out_bytes] staging tiles of the TMA-store epilogue (1024-B aligned)
  uint64_t* a_full = (uint64_t*)(sO + 2 * (size_t)p.out_bytes);
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
  float* s_slope = s_bias + p.BN;
  if (MODE == 0 && p.slope) for (int i = threadIdx.x; i < p.BN; i += blockDim.x) s_slope[i] = blockIdx.y * p.BN + i < p.Cout ? p.slope[blockIdx.y * p.BN + i] : 1.f;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int co0 = blockIdx.y * p.BN;
  const int NACC = SK * p.BN;               // accumulator columns per buffer

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < p.sa; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
    mbar_init(w_bar, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&t_full[i], 1); mbar_init(&t_empty[i], WS_EPI_WARPS); }
    mbar_init_fence();
  }
  if (warp == 1) tmem_alloc(tmem_slot, p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // ---- weights: every (K chunk, tap) tile of my BN output channels, once
      mbar_expect_tx(w_bar, (uint32_t)(taps * p.kchunks) * w_tile);
      for (int kc = 0; kc < p.kchunks; ++kc)
        for (int t = 0; t < taps; ++t) {
          const int a = t / p.KW, b = t % p.KW;
          const int wtap = p.flip ? (p.KH - 1 - a) * p.KW + (p.KW - 1 - b) : t;
          tma_load_3d(sW + (size_t)(kc * taps + t) * w_tile, &tmB, w_bar, kc * KC, co0, wtap);
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
    // One thread issues every tcgen05.mma of this CTA; the issue loop is table look-ups and adds only: descriptors are
    // (constant high word, low word = smem address >> 4), the row shift of each A window comes from a table built once,
    // and there is one mbarrier wait per (tile, K chunk).  MMA groups: SK == 1: one per tap (window shift a*P+b);
    // SK > 1: one per filter row (window shift a*P, the KW weight tiles of the row as ONE B operand of SK*BN rows).
    const int groups = SK == 1 ? taps : p.KH;
    for (int t = lane; t < groups; t += 32)
      s_arel[t] = SK == 1 ? (uint32_t)((t / p.KW) * p.P + (t % p.KW)) * 8u : (uint32_t)(t * p.P) * 8u;   // rows*128 B >> 4
    __syncwarp();
    if (lane == 0) {
      const uint32_t idesc = make_idesc(128, (p.dbg & 8) ? 128 : (p.dbg & 16) ? 256 : NACC, 0, 0);
      const uint64_t HI = (uint64_t)(64u | (1u << 14) | (2u << 29)) << 32;   // SBO = 1024 B, version 1, SWIZZLE_128B
      const uint32_t w_lo = smem_u32(sW) >> 4, w_step = (w_tile >> 4) * (uint32_t)SK;
      mbar_wait(w_bar, 0);
      int stage = 0; uint32_t phase = 0; int it = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
        const int acc = it & 1;
        mbar_wait(&t_empty[acc], ((uint32_t)(it >> 1) & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * NACC);
        uint32_t accum = 0;
        for (int kc = 0; kc < p.kchunks; ++kc) {
          mbar_wait(&a_full[stage], phase);
          tc_fence_after();
          const uint32_t a_lo = smem_u32(sA + (size_t)stage * p.a_stage_bytes) >> 4;
          const int ksteps = (kc == p.kchunks - 1) ? p.ksteps_last : 4;
          uint32_t b_lo = w_lo + (uint32_t)(kc * taps) * (w_tile >> 4);
          if (p.dbg & 4) {
          } else if (ksteps == 4) {
            for (int t = 0; t < groups; ++t, b_lo += w_step) {
              const uint32_t al = a_lo + s_arel[t];
              umma_bf16(d_tmem, HI | al, HI | b_lo, idesc, accum); accum = 1;
              umma_bf16(d_tmem, HI | (al + 2), HI | (b_lo + 2), idesc, 1);
              umma_bf16(d_tmem, HI | (al + 4), HI | (b_lo + 4), idesc, 1);
              umma_bf16(d_tmem, HI | (al + 6), HI | (b_lo + 6), idesc, 1);
            }
          } else if (ksteps == 2) {
            for (int t = 0; t < groups; ++t, b_lo += w_step) {
              const uint32_t al = a_lo + s_arel[t];
              umma_bf16(d_tmem, HI | al, HI | b_lo, idesc, accum); accum = 1;
              umma_bf16(d_tmem, HI | (al + 2), HI | (b_lo + 2), idesc, 1);
            }
          } else {
            for (int t = 0; t < groups; ++t, b_lo += w_step) {
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
    // TMEM lane quarter q = warp % 4 (a hardware rule); the two warps of a quarter take the even / odd 16-column chunks
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const int m = q * 32 + lane;
    const int r = m / p.P, wq = m - r * p.P;
    const bool in_tile = r < p.th && wq < p.tw;
    const bool issuer = threadIdx.x == 64;                 // first epilogue thread: issues the TMA stores
    // staging tile of the TMA-store epilogue: dense rows md = r*tw + wq of BN bf16 (rb bytes), 16-byte chunks XOR-swizzled
    // the way the store's tensor map expects (SWIZZLE_128B / 64B for 128 / 64-byte rows)
    const int md = r * p.tw + wq;
    const uint32_t rb = (uint32_t)p.BN * 2u;
    const uint32_t swz = rb == 128 ? (uint32_t)(md & 7) : (uint32_t)((md >> 1) & 3);
    int it = 0;
    // bias / activation slope of this thread's first two 16-column chunks in REGISTERS (every tile reuses them; as broadcast
    // LDS they doubled the instruction count of the lean epilogue: the folded-BatchNorm inference convs ran 2x slower)
    constexpr int NRB = MODE == 0 ? 16 : 1;
    float rbias[2][NRB], rsl[2][NRB];
    if (MODE == 0) {
#pragma unroll
      for (int ci = 0; ci < 2; ++ci)
#pragma unroll
        for (int i = 0; i < NRB; ++i) {
          const int c = (2 * ci + half) * 16 + i;
          rbias[ci][i] = c < p.BN ? s_bias[c] : 0.f;
          rsl[ci][i] = (p.slope && c < p.BN) ? s_slope[c] : 1.f;
        }
    }
    // BatchNorm statistics (BN <= 64 whenever they are requested): every thread keeps running sums of ITS tile row's
    // 64 columns over all tiles of this persistent CTA -- two FMAs per value; the cross-row reduction happens once, below
    constexpr int NR = MODE == 1 ? 32 : 1;
    float rs[NR], rq[NR];
#pragma unroll
    for (int i = 0; i < NR; ++i) { rs[i] = 0.f; rq[i] = 0.f; }
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
      int t = tile;
      const int tw_i = t % p.tiles_w; t /= p.tiles_w;
      const int th_i = t % p.tiles_h; const int n = t / p.tiles_h;
      const int oh = th_i * p.th + r, ow = tw_i * p.tw + wq;
      const bool row_ok = in_tile && oh < p.OH && ow < p.OW;
      bf16* yrow = p.y + (((long long)n * p.OH + oh) * p.OW + ow) * p.Cout + co0;
      const int acc = it & 1;
      uint8_t* so = sO + (size_t)acc * p.out_bytes + (size_t)md * rb;
      if (MODE == 2 && p.ex.res && row_ok) epi_prefetch_row(p.ex.res + (yrow - p.y), min(p.BN, p.Cout - co0) * 2);
      mbar_wait(&t_full[acc], (uint32_t)(it >> 1) & 1u);
      tc_fence_after();
      const uint32_t tacc = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * NACC);
      if (MODE == 2) {
        const bf16* rrow = p.ex.res ? p.ex.res + (yrow - p.y) : nullptr;
        for (int c = half * 64; c < p.BN; c += 128)
          epi_group64(tacc + (uint32_t)c, s_sum + c, p.BN, p.ex.res_scale, co0 + c,
                      p.BN - c, p.Cout, yrow + c, rrow ? rrow + c : nullptr, row_ok);
      } else
#pragma unroll
      for (int ci = 0; ci < 8; ++ci) {
        const int cc = 2 * ci + half;
        const int c = cc * 16;
        if (c >= p.BN || (p.dbg & 2)) break;
        float v[16];
        {
          // the SK column groups of this chunk, at most three TMEM loads in flight per wait (register budget)
          constexpr int G0 = SK < 3 ? SK : 3;
          uint32_t u[G0][16];
#pragma unroll
          for (int j = 0; j < G0; ++j) tmem_ld16_nowait(tacc + (uint32_t)(j * p.BN + c), u[j]);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(u[0][i]);
          // filter-row stacking: out[q] = sum_j D_j[q + j]; pixel q + j is lane + j (P divides 32, q + j stays inside
          // the pitched row for every valid output column)
#pragma unroll
          for (int j = 1; j < G0; ++j)
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] += __shfl_down_sync(0xffffffffu, __uint_as_float(u[j][i]), j);
          if (SK > 3) {
#pragma unroll
            for (int j = 3; j < SK; ++j) tmem_ld16_nowait(tacc + (uint32_t)(j * p.BN + c), u[j - 3]);
            tmem_ld_wait();
#pragma unroll
            for (int j = 3; j < SK; ++j)
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] += __shfl_down_sync(0xffffffffu, __uint_as_float(u[j - 3][i]), j);
          }
        }
        if (MODE == 0 && ci < 2) {
          if (p.bias) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] += rbias[ci][i % NRB];
          }
          if (p.slope) {                     // inference: the activation that follows a BatchNorm folded into weights + bias
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.f) + rsl[ci][i % NRB] * fminf(v[i], 0.f);
          }
        } else {
          if (p.bias) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] += s_bias[c + i];
          }
          if (MODE == 0 && p.slope) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.f) + s_slope[c + i] * fminf(v[i], 0.f);
          }
        }
        if (MODE == 1 && ci < 2 && row_ok) {
#pragma unroll
          for (int i = 0; i < 16; ++i) { const float t = bf16_round(v[i]); rs[(ci * 16 + i) % NR] += t; rq[(ci * 16 + i) % NR] = fmaf(t, t, rq[(ci * 16 + i) % NR]); }
        }
        if (p.tma_out) {
          if (in_tile) {
            uint32_t w[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) { __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]); w[i] = *reinterpret_cast<uint32_t*>(&h); }
            const uint32_t ch = (uint32_t)(c >> 3);
            *reinterpret_cast<uint4*>(so + ((ch ^ swz) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
            *reinterpret_cast<uint4*>(so + (((ch + 1) ^ swz) << 4)) = make_uint4(w[4], w[5], w[6], w[7]);
          }
        } else if (row_ok && co0 + c < p.Cout && !((p.dbg & 1) && v[0] != 123.456f)) {
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
      if (MODE != 2 && p.tma_out) {
        // my rows of the staging tile are written: make them visible to the async proxy; the issuer first makes sure
        // that every EARLIER store has finished reading shared memory (so the other buffer is free for the next tile)
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        if (issuer) tma_store_wait_read0();
        asm volatile("bar.sync 1, %0;" ::"r"(32 * WS_EPI_WARPS) : "memory");
        if (issuer && !(p.dbg & 1)) {
          tma_store_4d(&tmC, sO + (size_t)acc * p.out_bytes, co0, tw_i * p.tw, th_i * p.th, n);
          tma_store_commit();
        }
      }
    }
    if (MODE != 2 && p.tma_out && issuer) tma_store_wait_all();
    // BatchNorm statistics: the shared accumulators collect ALL tiles of this persistent CTA (fp32 over a few thousand
    // rows), published once -- a flush per tile would put tens of thousands of fp64 atomics on each channel's address
    if (MODE == 1) {
      // DETERMINISTIC within the CTA: every warp writes its partial sums into its own slot of the (now idle) first A stage,
      // and the four TMEM lane quarters are added in a fixed order -- shared-memory atomics here made two identical
      // forward passes differ in the last fp32 bit of the batch statistics, which the bf16 rounding of the normalised
      // activations and the train-mode BatchNorms downstream amplified to 4e-3 at the visual front-end's output.
      float* s_part = reinterpret_cast<float*>(sA);        // [4 quarters][sum | sum of squares][BN]
#pragma unroll
      for (int ci = 0; ci < 2; ++ci) {
        const int cc = 2 * ci + half;
        if (cc * 16 < p.BN) {
          float a[16], b[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) { a[i] = rs[(ci * 16 + i) % NR]; b[i] = rq[(ci * 16 + i) % NR]; }
          const float sa = colsum16(a, lane), sb = colsum16(b, lane);
          const int col = epi_col(lane);
          if (!(lane & 1)) { s_part[(q * 2) * p.BN + cc * 16 + col] = sa; s_part[(q * 2 + 1) * p.BN + cc * 16 + col] = sb; }
        }
      }
      asm volatile("bar.sync 1, %0;" ::"r"(32 * WS_EPI_WARPS) : "memory");
      for (int c = (int)threadIdx.x - 64; c < p.BN; c += 32 * WS_EPI_WARPS) {
        if (co0 + c < p.Cout) {
          double ts = 0.0, tq = 0.0;
#pragma unroll
          for (int qq = 0; qq < 4; ++qq) { ts += (double)s_part[(qq * 2) * p.BN + c]; tq += (double)s_part[(qq * 2 + 1) * p.BN + c]; }
          atomicAdd(p.stats + co0 + c, ts); atomicAdd(p.stats + p.Cout + co0 + c, tq);
        }
      }
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
// filter-row stacking needs a pitch that divides the warp (8 / 16 / 32): best useful-row fraction among those; 0 if none fits
double choose_ws_tile_stacked(int H, int W, int KW, int& th, int& tw, int& P) {
  double best = 0; th = tw = P = 0;
  for (int pp = 8; pp <= 32; pp *= 2) {
    int w = pp - (KW - 1); if (w < 1) continue; if (w > W) w = W;
    int h = 128 / pp; if (h > H) h = H;
    const double util = (double)(h * w) / 128.0 * ((double)W / (((W + w - 1) / w) * w)) * ((double)H / (((H + h - 1) / h) * h));
    if (util > best) { best = util; th = h; tw = w; P = pp; }
  }
  return best;
}

template <int MODE>
cudaError_t launch_ws(int sk, dim3 grid, size_t smem, cudaStream_t s, const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC,
                      const WsParams& p) {
#define VCA_WS_CASE(K)                                                                                                               \
  case K: {                                                                                                                          \
    static bool attr = false;                                                                                                        \
    if (!attr) {                                                                                                                     \
      cudaError_t e = cudaFuncSetAttribute(conv_tc_ws_kernel<MODE, K>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);     \
      if (e != cudaSuccess) return e;                                                                                                \
      attr = true;                                                                                                                   \
    }                                                                                                                                \
    conv_tc_ws_kernel<MODE, K><<<grid, WS_THREADS, smem, s>>>(tmA, tmB, tmC, p);                                                            \
    return cudaSuccess;                                                                                                              \
  }
  switch (sk) {
    VCA_WS_CASE(1) VCA_WS_CASE(2) VCA_WS_CASE(3) VCA_WS_CASE(4) VCA_WS_CASE(5)
  }
#undef VCA_WS_CASE
  return cudaErrorInvalidValue;
}

}  // namespace

// Tries the weights-stationary kernel.  Returns 1 if it was launched, 0 if the geometry does not fit (caller uses the
// streaming kernel), negative on error.  Arguments as conv_tc.cu::fwd_like.
int conv_ws_try(int NF, int IH, int IW, int Kdim, int OH, int OW, int Nout, int KH, int KW, int ph, int pw, int flip,
                    const void* x, const void* wpk, const float* bias, void* y, double* stats, const tc::EpiExtra* ex, cudaStream_t s,
                    const float* slope) {
  // x == nullptr: dry run -- 1 when this kernel would take the geometry (and, with stats != nullptr, emit the statistics)
  if (g_ws_mode == 0) return 0;
  const int taps = KH * KW, kchunks = (Kdim + KC - 1) / KC;
  if (taps < g_ws_min_taps || taps > 256 || kchunks > 2 || KW > 64) return 0;
  int bn = ((Nout + 15) / 16) * 16; if (bn > 256) bn = 256;
  // filter-row stacking: KW taps per MMA (N = KW * bn <= 256), combined by warp shuffles in the epilogue; not with the
  // inference epilogue (it reads the accumulator columns directly)
  int sk = (g_ws_stack && !ex && KW >= 2 && KW <= 5) ? KW : 1;
  int sth = 0, stw = 0, sP = 0;
  double sutil = 0;
  if (sk > 1) {
    sutil = choose_ws_tile_stacked(OH, OW, KW, sth, stw, sP);
    if (sutil < 0.5) sk = 1;
  }
  const size_t W_BUDGET = (size_t)g_ws_wbudget_kb * 1024;
  if (sk > 1) while (sk * bn > 256 && bn > 16) bn = (bn / 2 + 15) / 16 * 16;
  while ((size_t)taps * kchunks * bn * 128 > W_BUDGET && bn > 32) bn = (bn / 2 + 15) / 16 * 16;
  if ((size_t)taps * kchunks * bn * 128 > W_BUDGET) return 0;
  if (sk > 1 && sk * bn > 256) sk = 1;
  const int n_tiles = (Nout + bn - 1) / bn;
  if (g_ws_mode == 1) {
    // auto: only where the streaming kernel is bandwidth bound and re-reading the halo per n-tile stays cheap
    // (measured: ResNet layer 2, 128 -> 128 3x3 as two 64-channel CTA columns with stacked filter rows: 753 -> 897 TFLOP/s)
    if (Kdim > (sk > 1 ? 2 * g_ws_kmax : g_ws_kmax) || n_tiles > (sk > 1 ? 2 : 1)) return 0;
  }
  WsParams p;
  p.NF = NF; p.OH = OH; p.OW = OW; p.Cout = Nout;
  if (sk > 1) { p.th = sth; p.tw = stw; p.P = sP; }
  else { choose_ws_tile(OH, OW, KH, KW, p.th, p.tw); p.P = p.tw + KW - 1; }
  p.SK = sk;
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
  // TMA-store epilogue: rows of 64 / 128 bytes (the swizzle spans), whole channel tiles, two staging buffers
  const int rb = bn * 2;
  p.tma_out = g_ws_tma_out && !ex && (rb == 64 || rb == 128) && Nout % bn == 0;
  p.out_bytes = p.tma_out ? (((uint32_t)(p.th * p.tw) * (uint32_t)rb + 1023u) & ~1023u) : 0u;
  const size_t fixed = (size_t)p.w_bytes + 2 * (size_t)p.out_bytes + 1024 + 6144;   // + alignment + barriers/tables/statistics or epilogue vectors
  int sa = (int)((227 * 1024 - fixed) / p.a_stage_bytes);
  if (sa < 2 && p.tma_out) {                     // no room for the staging tiles: plain stores
    p.tma_out = 0; p.out_bytes = 0;
    sa = (int)((227 * 1024 - (size_t)p.w_bytes - 1024 - 6144) / p.a_stage_bytes);
  }
  if (sa > 4) sa = 4;
  if (sa < 2) return 0;
  p.sa = sa;
  p.tmem_cols = g_ws_dbg & 24 ? 512 : pow2_cols(2 * sk * bn);
  if (p.tmem_cols > 512) return 0;
  if (stats && bn > 64) return 0;          // the epilogue keeps the statistics of at most 64 columns in registers
  if (!x) return 1;
  p.base_off_mode = g_ws_base_off; p.dbg = g_ws_dbg;
  p.bias = bias; p.y = (bf16*)y; p.stats = stats; p.slope = slope;
  p.has_ex = ex != nullptr;
  if (ex) p.ex = *ex; else p.ex = EpiExtra{nullptr, nullptr, 0.f, 0, 0.f, nullptr};
  const size_t smem = (size_t)p.w_bytes + (size_t)sa * p.a_stage_bytes + 2 * (size_t)p.out_bytes + 1024 + 6144;

  CUtensorMap tmA, tmB, tmC;
  long long dA[4] = {Kdim, IW, IH, NF}; int bA[4] = {KC, p.P, p.th + KH - 1, 1};
  long long dB[3] = {Kdim, Nout, (long long)taps}; int bB[3] = {KC, bn, 1};
  if (bA[1] > 256 || bA[2] > 256) return 0;
  int rc = make_map(&tmA, x, 4, dA, bA); if (rc) return rc;
  rc = make_map(&tmB, wpk, 3, dB, bB); if (rc) return rc;
  if (p.tma_out) {
    long long dC[4] = {Nout, OW, OH, NF}; int bC[4] = {bn, p.tw, p.th, 1};
    rc = make_map(&tmC, y, 4, dC, bC, rb == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B);
    if (rc) return rc;
  } else tmC = tmA;
  int gx = vca_num_sms() / n_tiles; if (gx < 1) gx = 1; if (gx > p.num_tiles) gx = p.num_tiles;
  dim3 grid((unsigned)gx, (unsigned)n_tiles, 1);
  cudaError_t e;
  if (p.stats) e = launch_ws<1>(sk, grid, smem, s, tmA, tmB, tmC, p);
  else if (p.has_ex) e = launch_ws<2>(1, grid, smem, s, tmA, tmB, tmC, p);
  else e = launch_ws<0>(sk, grid, smem, s, tmA, tmB, tmC, p);
  if (e != cudaSuccess) { vca_set_error("conv_tc_ws_kernel launch set-up failed: %s", cudaGetErrorString(e)); return VCA_ERR_CUDA; }
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
  if (eq("ws_stack")) { g_ws_stack = value; return VCA_OK; }
  if (eq("ws_wbudget_kb")) { g_ws_wbudget_kb = value; return VCA_OK; }
  if (eq("ws_kmax")) { g_ws_kmax = value; return VCA_OK; }
  if (eq("ws_tma_out")) { g_ws_tma_out = value; return VCA_OK; }
  if (eq("ws_min_taps")) { g_ws_min_taps = value < 1 ? 1 : value; return VCA_OK; }
  if (eq("wgws_mode")) { g_wgws_mode = value; return VCA_OK; }
  if (eq("bn_vec")) { g_bn_vec = value; return VCA_OK; }
  if (eq("gru_cluster")) { g_gru_cluster = value; return VCA_OK; }
  if (eq("gru_bs")) { g_gru_bs = value; return VCA_OK; }
  if (eq("hs_mode")) { g_hs_mode = value; return VCA_OK; }
  if (eq("im2col_rb")) { g_im2col_rb = value < 0 ? 0 : value; return VCA_OK; }   // 0: the untiled reference kernel
  if (eq("wg_smem_kb")) { g_wg_smem_kb = value; return VCA_OK; }
  if (eq("fwd_smem_kb")) { g_fwd_smem_kb = value; return VCA_OK; }
  if (eq("wg_dbg")) { g_wg_dbg = value; return VCA_OK; }
  if (eq("wgws_mstack")) { g_wgws_mstack = value; return VCA_OK; }
  if (eq("wgws_waves")) { g_wgws_waves = value < 1 ? 1 : value; return VCA_OK; }
  if (eq("gl_fpw")) { g_gl_fpw = value < 1 ? 1 : value; return VCA_OK; }
  vca_set_error("vca_set_option: unknown key %s", key);
  return VCA_ERR_ARG;
}
}
