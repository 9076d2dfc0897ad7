// Multi-tap tcgen05 weight-gradient kernel for layers with <= 64 input channels (the 32/64-channel 5x5 layers at
// 80x300 / 40x150, the ResNet layer-1 3x3s, the (5,1) temporal stem conv).
//
//   dW[co, ci, tap] = sum_pixels dY[pixel, co] * X[pixel (+) tap, ci]
//
// conv_tc_wgrad_kernel gives every filter tap its own CTA and MMA (M = 128 padded output channels, N = Cin <= 64), so
// dY and X are re-fetched once per tap and every MMA is bound by its A-operand shared-memory read.  Here one CTA
// owns a whole GROUP of taps whose window shifts are equally spaced (the KW taps of one filter row: 1 pixel apart;
// or, for KW = 1, the KH taps: one pitch apart):
//   * the activation halo is fetched ONCE per pixel tile and kept in pitched pixel order (pitch P, a multiple of 8);
//   * dY is stored with the same pitch (one TMA per image row; the P - tw gap rows stay zero), so tile row m of both
//     operands is the same pixel;
//   * the B operand of ONE tcgen05.mma is MN-major with 64-channel atoms whose leading-dimension byte offset is the
//     tap spacing (128 B or P*128 B): atom g of the descriptor *is* tap g's shifted window.  A single MMA therefore
//     produces N = 64 * taps_in_group columns (up to 256) -- 4-5 taps per A-operand read instead of one.
// TMEM holds the group's accumulators (<= 320 columns); split-K over pixel tiles; fp32 red.add epilogue.
//
// Round 2 -- FILTER-ROW STACKING ALONG M for Cout <= 64: the MMA is 128 rows tall whatever Cout is, so a 64-channel layer
// left rows 64..127 computing garbage.  The A operand is MN-major with 64-channel atoms a leading-dimension offset apart:
// with that offset set to `mshift` image rows of the pitched dY tile, atom 1 IS the dY tile shifted down by mshift rows, and
//     D[64 + co] = sum_p dY[p + mshift rows][co] . X[p + tap]  =  sum_q dY[q][co] . X[q + tap - mshift rows]
// is the gradient of the tap `mshift` filter rows ABOVE the one rows 0..63 compute.  One pass therefore yields two filter
// rows (KW > 1) or 2 x 3 taps of a (5,1) filter: 3 -> 2 passes for 3x3, 5 -> 3 for 5x5, N = 320 -> 192 for the stem.
// The tile grid starts mshift rows above the image so that both halves see every dY row exactly once (rows outside the
// image are zero-filled by the TMA unit).
#include "tc_common.cuh"

using namespace tc;

namespace {

constexpr int KC = 64;
constexpr int ATOM_BYTES = 128 * 128;     // one 64-channel atom of a 128-row pitched tile

struct WgWsParams {
  int NF, OH, OW, Cout, Cin, taps;
  int KH, KW, ph, pw;
  int th, tw, P, tiles_w, tiles_h, num_tiles, tiles_per_split;
  int ci_tiles;          // 64-channel chunks of Cin (one CTA column each)
  int G;                 // taps per group
  int groups;            // number of groups (KH when grouping along kw, 1 when grouping along kh)
  int along_kh;          // 1: group = the KH taps of a (KH,1) filter, spacing P rows; 0: group = KW taps of row kh, spacing 1
  int x_rows;            // image rows in the X box
  int a_atoms;           // 64-channel atoms of dY actually loaded (1 or 2)
  int mstack;            // 1: rows 64..127 = the dY tile shifted by `mshift` image rows (Cout <= 64)
  int mshift;            // filter rows between the taps of the two halves
  int dy_rows;           // image rows of dY per tile (th + mshift when stacked)
  int x_row0[8];         // per group: first image row of the X box relative to (tile origin - ph)
  signed char top_tap[8][5], bot_tap[8][5];   // per group and N-slot: tap index whose gradient rows 0..63 / 64..127 hold (-1: none)
  int ksteps;            // ceil(th*P / 16)
  int stages;
  uint32_t x_stage_bytes, a_stage_bytes, x_tx_bytes, a_row_tx_bytes, tmem_cols;
  float* dw;
  int dbg;               // probe: 1 = no atomics (tools/conv_shapes.py --opt wg_dbg=1)
  int tm;                // 1: tap-major destination [taps][Cout][Cin], added to by TMA reduce (tmW)
};

__global__ void __launch_bounds__(192, 1) conv_tc_wgrad_ws_kernel(const __grid_constant__ CUtensorMap tmDY,
                                                                  const __grid_constant__ CUtensorMap tmX,
                                                                  const __grid_constant__ CUtensorMap tmW, const WgWsParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int S = p.stages;
  uint8_t* sA = smem;                                        // dY: [stage][atom][128 rows pitched][128 B]
  uint8_t* sX = smem + (size_t)S * p.a_stage_bytes;          // X halo: [stage][rows pitched][128 B]
  uint64_t* full = (uint64_t*)(sX + (size_t)S * p.x_stage_bytes);
  uint64_t* empty = full + S;
  uint64_t* accum_bar = empty + S;
  uint32_t* tmem_slot = (uint32_t*)(accum_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int grp = blockIdx.z;                                // tap group
  const int co0 = (blockIdx.y / p.ci_tiles) * 128;
  const int ci0 = (blockIdx.y % p.ci_tiles) * KC;
  const int t_beg = blockIdx.x * p.tiles_per_split;
  const int t_end = min(t_beg + p.tiles_per_split, p.num_tiles);
  const int iters = t_end - t_beg;

  {  // gap rows / never-written rows must read as zero (they are part of the K reduction)
    uint4 z = make_uint4(0, 0, 0, 0);
    uint4* ptr = (uint4*)smem;
    const size_t n16 = ((size_t)S * (p.a_stage_bytes + p.x_stage_bytes)) / 16;
    for (size_t i = threadIdx.x; i < n16; i += blockDim.x) ptr[i] = z;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 0 && lane == 0) {
    for (int i = 0; i < S; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(accum_bar, 1);
    mbar_init_fence();
  }
  if (warp == 1) tmem_alloc(tmem_slot, p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (iters > 0) {
    if (warp == 0) {
      if (lane == 0) {
        int stage = 0; uint32_t phase = 0;
        int tw_i, th_i, n;                                           // tile coordinates, advanced without divisions
        { int t = t_beg; tw_i = t % p.tiles_w; t /= p.tiles_w; th_i = t % p.tiles_h; n = t / p.tiles_h; }
        for (int it = 0; it < iters; ++it) {
          const int ow0 = tw_i * p.tw, oh0 = th_i * p.th - (p.mstack ? p.mshift : 0);
          const int n_cur = n;
          if (++tw_i == p.tiles_w) { tw_i = 0; if (++th_i == p.tiles_h) { th_i = 0; ++n; } }
          mbar_wait(&empty[stage], phase ^ 1);
          mbar_expect_tx(&full[stage], p.x_tx_bytes + (uint32_t)(p.a_atoms * p.dy_rows) * p.a_row_tx_bytes);
          // X halo: rows oh0 - ph + x_row0 ... (+ x_rows), columns ow0 - pw ... (+ P)
          tma_load_4d(sX + (size_t)stage * p.x_stage_bytes, &tmX, &full[stage], ci0, ow0 - p.pw, oh0 - p.ph + p.x_row0[grp], n_cur);
          // dY: one box per image row, written at pitch P so that tile row r*P + w is pixel (oh0 + r, ow0 + w)
          for (int a = 0; a < p.a_atoms; ++a)
            for (int r = 0; r < p.dy_rows; ++r)
              tma_load_4d(sA + (size_t)stage * p.a_stage_bytes + (size_t)a * ATOM_BYTES + (size_t)r * p.P * 128, &tmDY, &full[stage],
                          co0 + a * KC, ow0, oh0 + r, n_cur);
          if (++stage == S) { stage = 0; phase ^= 1; }
        }
      }
      __syncwarp();
    } else if (warp == 1) {
      if (lane == 0) {
        // N of the two MMA parts: up to 4 taps (256 columns) + the remaining taps
        const int g1 = p.G > 4 ? 4 : p.G, g2 = p.G - g1;
        const uint32_t idesc1 = make_idesc(128, g1 * KC, 1, 1);
        const uint32_t idesc2 = g2 > 0 ? make_idesc(128, g2 * KC, 1, 1) : 0u;
        const uint32_t tap_lbo = (uint32_t)(p.along_kh ? p.P : 1) * 128u;   // byte distance between consecutive taps' windows
        const uint32_t a_lbo = p.mstack ? (uint32_t)(p.mshift * p.P) * 128u : (uint32_t)ATOM_BYTES;   // distance between the two M atoms
        int stage = 0; uint32_t phase = 0, accum = 0;
        for (int it = 0; it < iters; ++it) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t a0 = smem_u32(sA + (size_t)stage * p.a_stage_bytes);
          const uint32_t x0 = smem_u32(sX + (size_t)stage * p.x_stage_bytes);
          for (int k = 0; k < p.ksteps; ++k) {
            const uint64_t ad = make_desc(a0 + k * 2048, a_lbo, 1024);
            const uint64_t bd1 = make_desc(x0 + k * 2048, tap_lbo, 1024);
            umma_bf16(tmem_base, ad, bd1, idesc1, accum);
            if (g2 > 0) {
              const uint64_t bd2 = make_desc(x0 + k * 2048 + (uint32_t)g1 * tap_lbo, tap_lbo, 1024);
              umma_bf16(tmem_base + (uint32_t)(g1 * KC), ad, bd2, idesc2, accum);
            }
            accum = 1;
          }
          umma_commit(&empty[stage]);
          if (++stage == S) { stage = 0; phase ^= 1; }
        }
        umma_commit(accum_bar);
      }
      __syncwarp();
    } else {
      const int q = warp & 3;
      const bool bottom = p.mstack && q >= 2;                          // rows 64..127: the shifted dY tile
      const int co = p.mstack ? (q & 1) * 32 + lane : co0 + q * 32 + lane;
      mbar_wait(accum_bar, 0);
      tc_fence_after();
      if (p.tm) {
        // tap-major destination: stage every tap's [128 x 64] block (two 32-column blocks) in the idle pipeline buffers,
        // then one TMA reduce-add box per (tap, 32 input channels, 64 output channels)
        for (int g = 0; g < p.G; ++g)
          for (int c = 0; c < KC; c += 16) {
            if (ci0 + c >= p.Cin) break;                                // warp-uniform
            float v[16];
            tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(g * KC + c), v);
            dw_stage16(smem + (size_t)(g * 2 + (c >> 5)) * 16384, q * 32 + lane, c & 16, v);
          }
        fence_async_smem();
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (threadIdx.x == 64 && !p.dbg) {
          for (int g = 0; g < p.G; ++g)
            for (int c = 0; c < KC && ci0 + c < p.Cin; c += 32)
              for (int h = 0; h < 2; ++h) {
                // rows 0..63 / 64..127 of the block: output channels co0 + 64 h of the top tap, or (stacked) all 64 output
                // channels of the top / bottom tap
                const int tap = p.mstack ? (h ? p.bot_tap[grp][g] : p.top_tap[grp][g]) : p.top_tap[grp][g];
                const int cob = p.mstack ? 0 : co0 + 64 * h;
                if (tap < 0 || cob >= p.Cout) continue;
                tma_reduce_add_3d(&tmW, smem + (size_t)(g * 2 + (c >> 5)) * 16384 + h * 8192, ci0 + c, cob, tap);
              }
          bulk_commit();
          bulk_wait_all();
        }
      } else
      for (int g = 0; g < p.G; ++g) {
        const int tap = bottom ? p.bot_tap[grp][g] : p.top_tap[grp][g];
        if (tap < 0) continue;                                         // warp-uniform
        for (int c = 0; c < KC; c += 16) {
          if (ci0 + c >= p.Cin) break;                                // warp-uniform
          float v[16];
          tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(g * KC + c), v);
          if (co < p.Cout) {
#pragma unroll
            for (int i = 0; i < 16; ++i)
              if (ci0 + c + i < p.Cin && !(p.dbg && v[i] != 123.456f)) atomicAdd(p.dw + ((long long)co * p.Cin + ci0 + c + i) * p.taps + tap, v[i]);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, p.tmem_cols);
}

}  // namespace

int g_wgws_mode = 1;   // 0 off, 1 auto, 2 any Cin
int g_wg_dbg = 0;       // probe switch shared by both wgrad kernels: 1 = skip the red.add epilogue
int g_wgws_mstack = 1; // filter-row stacking along M for Cout <= 64
int g_wgws_waves = 1;  // "wgws_waves": CTAs per SM over the kernel's life (more = shorter CTAs, friendlier to concurrent streams)

// 1 = launched, 0 = not applicable, < 0 error.  dw fp32 [Cout][Cin][taps], zero on entry.
int conv_wgrad_ws_try(const ConvGeom& g, const void* dy, const void* x, float* dw, cudaStream_t s, int tm) {
  if (!g_wgws_mode) return 0;
  const int taps = g.KH * g.KW;
  // Cin > 64 runs as 64-channel chunks (one CTA column each): the dY tile is then re-read per chunk, but a stage still
  // feeds G x 8 MMAs from ~50 KB, against 8 MMAs from 64 KB in conv_tc_wgrad_kernel.
  // Measured (tools/conv_shapes.py --wgws): faster for 128 ch (408 -> 513 TF/s 5x5, 505 -> 579 3x3), 256 ch 5x5
  // (711 -> 868) and 640 -> 512 (744 -> 857); slower where the 256-wide streaming tiles fit exactly (512 ch: 1054 vs
  // 700) and on small maps where the pitched tile wastes rows (7x7: 610 vs 326) -- hence the rule below.
  if (g.Cin > 4096 || taps < 2 || g.KW > 5 || g.KH > 8) return 0;
  WgWsParams p;
  p.NF = g.N; p.OH = g.OH; p.OW = g.OW; p.Cout = g.Cout; p.Cin = g.Cin; p.taps = taps;
  p.KH = g.KH; p.KW = g.KW; p.ph = g.ph; p.pw = g.pw;
  p.ci_tiles = (g.Cin + KC - 1) / KC;
  p.along_kh = g.KW == 1;
  p.G = p.along_kh ? g.KH : g.KW;
  p.groups = p.along_kh ? 1 : g.KH;
  if (p.G > 5) return 0;
  // tile: pitch P a multiple of 8 (every image row of the tile starts on a 1024-B swizzle period), th*P <= 128
  int best_tw = 0, best_th = 0; double best = -1;
  for (int P = 8; P <= 128; P += 8) {
    const int tw = P - (g.KW - 1);
    if (tw < 1) continue;
    int th = 128 / P; if (th > g.OH) th = g.OH; if (th > 8) th = 8; if (th < 1) continue;
    // useful output pixels per 128 MMA rows spent, over the whole image
    const double util = (double)g.OH * g.OW / ((double)((g.OW + tw - 1) / tw) * ((g.OH + th - 1) / th) * 128.0);
    if (util > best) { best = util; best_tw = tw; best_th = th; }
  }
  if (best_tw == 0) return 0;
  if (g_wgws_mode == 1 && g.Cin > 2 * KC && !(best >= 0.6 && (g.Cin <= 256 || g.Cin % 256 != 0))) return 0;
  p.tw = best_tw; p.th = best_th; p.P = p.tw + g.KW - 1;
  p.tiles_w = (g.OW + p.tw - 1) / p.tw; p.tiles_h = (g.OH + p.th - 1) / p.th;
  const long long nt = (long long)g.N * p.tiles_w * p.tiles_h;
  if (nt > 0x7fffffff) return 0;
  p.num_tiles = (int)nt;
  p.a_atoms = g.Cout > KC ? 2 : 1;
  // filter-row stacking along M (see the header): the groups and what each accumulator block holds
  // (the K range of an MMA is whole 16-row steps: the tile must end on one, or the top half would also sum the first
  //  pixels of the extra dY row)
  p.mstack = g_wgws_mstack && g.Cout <= KC && g.KH >= 2 && (p.th * p.P) % 16 == 0;
  p.mshift = 0;
  for (int gi = 0; gi < 8; ++gi) { p.x_row0[gi] = 0; for (int j = 0; j < 5; ++j) { p.top_tap[gi][j] = -1; p.bot_tap[gi][j] = -1; } }
  if (!p.mstack) {
    if (p.groups > 8) return 0;
    for (int gi = 0; gi < p.groups; ++gi) {
      p.x_row0[gi] = p.along_kh ? 0 : gi;
      for (int j = 0; j < p.G; ++j) p.top_tap[gi][j] = (signed char)(p.along_kh ? j * g.KW : gi * g.KW + j);
    }
  } else if (!p.along_kh) {
    // KW > 1: rows 0..63 filter row kt, rows 64..127 filter row kt - 1;  kt = 1, 3, ..., and KH - 1 alone when KH is odd
    p.mshift = 1;
    p.groups = (g.KH + 1) / 2;
    if (p.groups > 8) return 0;
    for (int gi = 0; gi < p.groups; ++gi) {
      const int kt = 2 * gi + 1 < g.KH ? 2 * gi + 1 : g.KH - 1;
      const bool pair = 2 * gi + 1 < g.KH;
      p.x_row0[gi] = kt;
      for (int j = 0; j < p.G; ++j) { p.top_tap[gi][j] = (signed char)(kt * g.KW + j); p.bot_tap[gi][j] = (signed char)(pair ? (kt - 1) * g.KW + j : -1); }
    }
  } else {
    // (KH,1) filter: N slots = taps t0 .. KH-1, rows 64..127 the taps t0 rows above them (0 .. t0-1)
    const int Gn = (g.KH + 1) / 2, t0 = g.KH - Gn;
    p.mshift = t0; p.G = Gn; p.groups = 1;
    p.x_row0[0] = t0;
    for (int j = 0; j < Gn; ++j) { p.top_tap[0][j] = (signed char)((t0 + j) * g.KW); p.bot_tap[0][j] = (signed char)(j < t0 ? j * g.KW : -1); }
  }
  p.dy_rows = p.th + p.mshift;
  p.x_rows = p.along_kh ? p.th + p.G - 1 : p.th;
  if (p.mstack) p.tiles_h = (g.OH + p.mshift + p.th - 1) / p.th;
  {
    const long long nt2 = (long long)g.N * p.tiles_w * p.tiles_h;
    if (nt2 > 0x7fffffff) return 0;
    p.num_tiles = (int)nt2;
  }
  p.ksteps = (p.th * p.P + 15) / 16;
  p.x_tx_bytes = (uint32_t)(p.P * p.x_rows) * 128u;
  p.a_row_tx_bytes = (uint32_t)p.tw * 128u;
  // rows any tap window may touch: ksteps*16 rows starting at the largest shift
  const int max_shift = (p.G - 1) * (p.along_kh ? p.P : 1);
  const uint32_t x_need = (uint32_t)(p.ksteps * 16 + max_shift) * 128u;
  p.x_stage_bytes = ((x_need > p.x_tx_bytes ? x_need : p.x_tx_bytes) + 1023u) & ~1023u;
  // Cout <= 64: only atom 0 is loaded; the M = 128 MMA then also reads the 16 KB behind it (the next stage / the X
  // buffers -- valid shared memory) into accumulator rows 64..127, which the epilogue never stores.
  p.a_stage_bytes = p.mstack ? (((uint32_t)(p.ksteps * 16 + p.mshift * p.P) * 128u + 1023u) & ~1023u) : (uint32_t)p.a_atoms * ATOM_BYTES;
  const size_t stage_bytes = (size_t)p.x_stage_bytes + p.a_stage_bytes;
  int stages = (int)((200 * 1024) / stage_bytes);
  if (stages > 8) stages = 8;
  if (stages < 2) return 0;
  p.stages = stages;
  p.tmem_cols = pow2_cols(p.G * KC);
  p.dw = dw; p.dbg = g_wg_dbg; p.tm = tm;
  const int co_tiles = (g.Cout + 127) / 128;
  const long long base_ctas = (long long)co_tiles * p.ci_tiles * p.groups;
  extern int g_wgws_waves;
  int split = (int)(g_wgws_waves * vca_num_sms() / base_ctas);   // one CTA per SM (smem-bound): whole waves only
  if (split < 1) split = 1; if (split > p.num_tiles) split = p.num_tiles;
  p.tiles_per_split = (p.num_tiles + split - 1) / split;
  split = (p.num_tiles + p.tiles_per_split - 1) / p.tiles_per_split;
  const size_t smem = (size_t)stages * stage_bytes + ATOM_BYTES + 1024 + 512;   // + one atom of read-only slack

  if (tm && (size_t)stages * stage_bytes < (size_t)p.G * 2 * 16384) return 0;   // no room to stage the tile: streaming kernel
  CUtensorMap tmDY, tmX, tmW;
  long long dY[4] = {g.Cout, g.OW, g.OH, g.N}; int bY[4] = {KC, p.tw, 1, 1};
  long long dX[4] = {g.Cin, g.IW, g.IH, g.N}; int bX[4] = {KC, p.P, p.x_rows, 1};
  if (bX[1] > 256 || bX[2] > 256 || bY[1] > 256) return 0;
  int rc = make_map(&tmDY, dy, 4, dY, bY); if (rc) return rc;
  rc = make_map(&tmX, x, 4, dX, bX); if (rc) return rc;
  if (tm) { rc = make_map_dw(&tmW, dw, g.Cin, g.Cout, taps); if (rc) return rc; } else tmW = tmX;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(conv_tc_wgrad_ws_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) {
      vca_set_error("cudaFuncSetAttribute(conv_tc_wgrad_ws_kernel) failed"); return VCA_ERR_CUDA;
    }
    attr_set = true;
  }
  if ((long long)co_tiles * p.ci_tiles > 65535) return 0;
  dim3 grid((unsigned)split, (unsigned)(co_tiles * p.ci_tiles), (unsigned)p.groups);
  conv_tc_wgrad_ws_kernel<<<grid, 192, smem, s>>>(tmDY, tmX, tmW, p);
  VCA_LAUNCH_CHECK();
  return 1;
}
