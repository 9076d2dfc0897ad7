// tcgen05 / TMEM / TMA implicit-GEMM convolution kernels for sm_100a (bf16 operands, fp32 accumulate).
//
// Forward and dgrad (stride 1) share one kernel:  for an output tile of up to 128 pixels (a 3-D box tn x th x tw of
// the channels-last activation) and BN output channels,
//     D[pixel, co] = sum_{tap} sum_{cin chunk of 64}  A_tap[pixel, 64] * W_tap[co, 64]^T
// where A_tap is the SAME box shifted by the tap offset -- fetched by one 4-D TMA box load whose out-of-bounds
// elements are zero-filled by the TMA unit (that is the convolution's zero padding; nothing is materialised) --
// and W_tap is a [BN x 64] K-major slab of the packed weights.  Both land in 128B-swizzled shared memory and are
// consumed by tcgen05.mma (UMMA 128 x BN x 16) into a TMEM accumulator; 4 epilogue warps read it back with
// tcgen05.ld, add the bias and store bf16 channels-last.
//
// Wgrad:  dW[tap][co, ci] = sum_pixels dY[pixel, co] * X[pixel (+) tap, ci]  -- the reduction runs over pixels, so
// both operands are MN-major (the pixel index is the slow smem dimension); split-K over pixel tiles, fp32 red.add
// into the parameter layout.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2..5 = epilogue
// (TMEM lane quarter = warp_id % 4).  Descriptor encodings follow the PTX ISA "tcgen05 shared memory descriptor"
// / "instruction descriptor" tables.
#include "tc_common.cuh"

using namespace tc;

// conv_tc_ws.cu: weights-stationary / halo-resident variant for <=128-channel layers (1 = launched, 0 = not applicable)
int conv_ws_try(int NF, int IH, int IW, int Kdim, int OH, int OW, int Nout, int KH, int KW, int ph, int pw, int flip,
                const void* x, const void* wpk, const float* bias, void* y, double* stats, const tc::EpiExtra* ex, cudaStream_t s,
                const float* slope = nullptr);
// conv_tc_hs.cu: halo-resident activations + streamed weights for 64..128 output channels (same return convention)
int conv_hs_try(int NF, int IH, int IW, int Kdim, int OH, int OW, int Nout, int KH, int KW, int ph, int pw, int flip,
                const void* x, const void* wpk, const float* bias, void* y, double* stats, const tc::EpiExtra* ex, cudaStream_t s);
// conv_tc_wgrad_ws.cu: multi-tap weight-gradient kernel for <= 64 input channels (same return convention)
int conv_wgrad_ws_try(const ConvGeom& g, const void* dy, const void* x, float* dw, cudaStream_t s, int tm);
extern int g_wg_dbg;
int g_wg_smem_kb = 190;    // shared-memory budget per CTA of the streaming wgrad kernel
int g_fwd_smem_kb = 100;   // shared-memory budget per CTA of the streaming forward kernel on multi-wave grids (100: two CTAs per SM)

namespace {

constexpr int KC = 64;              // bf16 channels per K chunk = 128 B = one SWIZZLE_128B row
constexpr int TILE_ROWS = 128;      // UMMA M
constexpr int A_STAGE_BYTES = TILE_ROWS * 128;

struct FwdParams {
  int NF, OH, OW, Cout;          // output tensor [NF, OH, OW, Cout]
  int IH, IW;                    // input spatial dims (to skip taps whose whole shifted box is padding)
  int tn, th, tw;                // pixel box
  int tiles_w, tiles_h;          // tile counts along W and H (tiles along NF = gridDim.x / (tiles_w*tiles_h))
  int KH, KW, ph, pw;            // input coord = output coord + k - p
  int kchunks;                   // ceil(Kdim / 64)
  int ksteps_last;               // 16-channel MMA steps in the last K chunk (4 unless Kdim % 64 != 0)
  int BN;                      // output channels per CTA (multiple of 16, <= 256)
  int stages;
  int flip;                      // 1: weight tap index is mirrored (dgrad)
  uint32_t a_bytes, b_bytes;     // bytes the two TMA loads of one stage deliver
  uint32_t tmem_cols;
  int splits, taps_per_split;    // split-K over the filter taps: blockIdx.z takes taps [z*per, (z+1)*per)
  float* ws;                     // split-K only: fp32 [splits][pixels][Cout] partial sums, added in order by splitk_finish_kernel
  const float* bias;             // [Cout] or null
  double* stats;                 // [2 * Cout] BatchNorm sum / sum-of-squares accumulators (fp64, added to) or null
  EpiExtra ex;                   // inference epilogue (scale / residual / activation); has_ex = 0: plain bias epilogue
  int has_ex;
  bf16* y;
};

__global__ void __launch_bounds__(192) conv_tc_fwd_kernel(const __grid_constant__ CUtensorMap tmA,
                                                          const __grid_constant__ CUtensorMap tmB, const FwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int S = p.stages;
  const uint32_t b_stage = (uint32_t)p.BN * 128u;
  uint8_t* sA = smem;
  uint8_t* sB = smem + (size_t)S * A_STAGE_BYTES;
  uint64_t* full = (uint64_t*)(sB + (size_t)S * b_stage);
  uint64_t* empty = full + S;
  uint64_t* accum_bar = empty + S;
  uint32_t* tmem_slot = (uint32_t*)(accum_bar + 1);
  float* s_sum = (float*)(tmem_slot + 2);   // [BN] + [BN]: per-channel sum / sum of squares of this tile (BatchNorm statistics)
  float* s_sq = s_sum + p.BN;               // (the same space holds the [3][BN] epilogue vectors of an inference launch)
  if (p.stats) for (int i = threadIdx.x; i < 2 * p.BN; i += blockDim.x) s_sum[i] = 0.f;
  if (p.has_ex) epi_stage(s_sum, p.BN, blockIdx.y * p.BN, p.Cout, p.ex, p.bias);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // tile coordinates
  int t = blockIdx.x;
  const int tw_i = t % p.tiles_w; t /= p.tiles_w;
  const int th_i = t % p.tiles_h; const int tn_i = t / p.tiles_h;
  const int ow0 = tw_i * p.tw, oh0 = th_i * p.th, n0 = tn_i * p.tn;
  const int co0 = blockIdx.y * p.BN;
  const int tap_beg = blockIdx.z * p.taps_per_split;
  const int tap_end = min(tap_beg + p.taps_per_split, p.KH * p.KW);

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < S; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(accum_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(tmem_slot, p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      int a = tap_beg / p.KW, b = tap_beg - a * p.KW;            // no division inside the loops
      for (int tap = tap_beg; tap < tap_end; ++tap) {
        const int r0 = oh0 + a - p.ph, c0 = ow0 + b - p.pw;
        if (!(r0 >= p.IH || r0 + p.th <= 0 || c0 >= p.IW || c0 + p.tw <= 0)) {   // else: the whole box is zero padding
          const int wtap = p.flip ? (p.KH - 1 - a) * p.KW + (p.KW - 1 - b) : tap;
          for (int kc = 0; kc < p.kchunks; ++kc) {
            mbar_wait(&empty[stage], phase ^ 1);
            mbar_expect_tx(&full[stage], p.a_bytes + p.b_bytes);
            tma_load_4d(sA + (size_t)stage * A_STAGE_BYTES, &tmA, &full[stage], kc * KC, c0, r0, n0);
            tma_load_3d(sB + (size_t)stage * b_stage, &tmB, &full[stage], kc * KC, co0, wtap);
            if (++stage == S) { stage = 0; phase ^= 1; }
          }
        }
        if (++b == p.KW) { b = 0; ++a; }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc(TILE_ROWS, p.BN, 0, 0);
      int stage = 0; uint32_t phase = 0, started = 0;
      int a = tap_beg / p.KW, b = tap_beg - a * p.KW;
      for (int tap = tap_beg; tap < tap_end; ++tap) {
        // same skip rule as the producer (dgrad onto a map much taller than dY: most filter rows see only padding)
        const int r0 = oh0 + a - p.ph, c0 = ow0 + b - p.pw;
        const bool dead = r0 >= p.IH || r0 + p.th <= 0 || c0 >= p.IW || c0 + p.tw <= 0;
        if (++b == p.KW) { b = 0; ++a; }
        if (dead) continue;
        for (int kc = 0; kc < p.kchunks; ++kc) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t a0 = smem_u32(sA + (size_t)stage * A_STAGE_BYTES);
          const uint32_t b0 = smem_u32(sB + (size_t)stage * b_stage);
          const int ksteps = (kc == p.kchunks - 1) ? p.ksteps_last : KC / 16;   // skip the zero-padded tail of a K chunk
#pragma unroll
          for (int k = 0; k < KC / 16; ++k) {
            if (k < ksteps) {
              const uint64_t ad = make_desc(a0 + k * 32, 0, 1024);
              const uint64_t bd = make_desc(b0 + k * 32, 0, 1024);
              umma_bf16(tmem_base, ad, bd, idesc, (started | (uint32_t)k) != 0);
            }
          }
          started = 1;
          umma_commit(&empty[stage]);  // implies tcgen05.fence::before_thread_sync
          if (++stage == S) { stage = 0; phase ^= 1; }
        }
      }
      umma_commit(accum_bar);
    }
    __syncwarp();
  } else {
    // ---- epilogue: TMEM -> registers -> (+bias) -> bf16 -> global (channels-last)
    const int q = warp & 3;                 // TMEM lane quarter this warp may access
    const int r = q * 32 + lane;            // tile row = pixel index inside the box
    const int hw = p.th * p.tw;
    const int in = r / hw, rem = r - in * hw;
    const int ih = rem / p.tw, iw = rem - ih * p.tw;
    const int n = n0 + in, oh = oh0 + ih, ow = ow0 + iw;
    const bool row_ok = (r < p.tn * hw) && n < p.NF && oh < p.OH && ow < p.OW;
    bf16* yrow = p.y + (((long long)n * p.OH + oh) * p.OW + ow) * p.Cout + co0;
    if (p.has_ex && p.ex.res && row_ok) epi_prefetch_row(p.ex.res + (yrow - p.y), min(p.BN, p.Cout - co0) * 2);
    mbar_wait(accum_bar, 0);
    tc_fence_after();
    if (p.ws) {
      // split-K: this CTA's partial sums go to ITS OWN slab ws[split][pixel][Cout] with plain vector stores (zeros if every
      // tap of the split was padding); splitk_finish_kernel adds the slabs in split order -- deterministic, no atomics, no
      // memset (fp32 red.add into one shared slab made identical forward passes differ in the last bit, which the bf16
      // rounding and the train-mode BatchNorms downstream amplify)
      bool any = false;
      for (int tap = tap_beg; tap < tap_end; ++tap) {
        const int a = tap / p.KW, b = tap % p.KW;
        const int r0 = oh0 + a - p.ph, c0 = ow0 + b - p.pw;
        any = any || !(r0 >= p.IH || r0 + p.th <= 0 || c0 >= p.IW || c0 + p.tw <= 0);
      }
      float* wrow = p.ws + ((long long)blockIdx.z * p.NF * p.OH * p.OW + (((long long)n * p.OH + oh) * p.OW + ow)) * p.Cout + co0;
      for (int c = 0; c < p.BN; c += 16) {
        float v[16];
        if (any) tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c, v);
        else {
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = 0.f;
        }
        if (row_ok && co0 + c < p.Cout) {
          if (co0 + c + 16 <= p.Cout && (p.Cout & 3) == 0) {
#pragma unroll
            for (int i = 0; i < 4; ++i) *reinterpret_cast<float4*>(wrow + c + 4 * i) = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
          } else {
            for (int i = 0; i < 16 && co0 + c + i < p.Cout; ++i) wrow[c + i] = v[i];
          }
        }
      }
    } else if (p.has_ex) {
      // inference epilogue (scale / shift / residual / activation), 64 columns at a time
      const bf16* rrow = p.ex.res ? p.ex.res + (yrow - p.y) : nullptr;
      for (int c = 0; c < p.BN; c += 64)
        epi_group64(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c, s_sum + c, p.BN, p.ex.res_scale, co0 + c, p.BN - c, p.Cout,
                    yrow + c, rrow ? rrow + c : nullptr, row_ok);
    } else {
    for (int c = 0; c < p.BN; c += 16) {
      float v[16];
      tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c, v);
      if (p.bias && co0 + c < p.Cout) {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] += (co0 + c + i < p.Cout) ? __ldg(p.bias + co0 + c + i) : 0.f;
      }
      if (p.stats) epi_stats16(v, row_ok, co0 + c, p.Cout, s_sum + c, s_sq + c, lane);
      if (row_ok && co0 + c < p.Cout) {
        if (co0 + c + 16 <= p.Cout) {
          uint4 o0, o1;
          __nv_bfloat162 h;
          uint32_t w[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) { h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]); w[i] = *reinterpret_cast<uint32_t*>(&h); }
          o0 = make_uint4(w[0], w[1], w[2], w[3]); o1 = make_uint4(w[4], w[5], w[6], w[7]);
          *reinterpret_cast<uint4*>(yrow + c) = o0;
          *reinterpret_cast<uint4*>(yrow + c + 8) = o1;
        } else {
          for (int i = 0; i < 16 && co0 + c + i < p.Cout; ++i) yrow[c + i] = __float2bfloat16_rn(v[i]);
        }
      }
    }
    if (p.stats) epi_stats_flush(s_sum, s_sq, p.BN, co0, p.Cout, p.stats, (int)threadIdx.x - 64);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, p.tmem_cols);
}

// y[m][c] = bf16(sum_s ws[s][m][c] + bias[c]), slabs added in split order: finishes a split-K convolution
__global__ void splitk_finish_kernel(const float* __restrict__ ws, const float* __restrict__ bias, bf16* __restrict__ y,
                                     long long total, int Cout, int splits) {
  for (long long i = (blockIdx.x * (long long)blockDim.x + threadIdx.x) * 8; i < total; i += (long long)gridDim.x * blockDim.x * 8) {
    if (i + 8 <= total && Cout % 8 == 0) {
      float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      for (int sp = 0; sp < splits; ++sp) {
        const float4 a = *reinterpret_cast<const float4*>(ws + sp * total + i), b = *reinterpret_cast<const float4*>(ws + sp * total + i + 4);
        v[0] += a.x; v[1] += a.y; v[2] += a.z; v[3] += a.w; v[4] += b.x; v[5] += b.y; v[6] += b.z; v[7] += b.w;
      }
      const int c = (int)(i % Cout);
      uint32_t w[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * k] + (bias ? bias[c + 2 * k] : 0.f), v[2 * k + 1] + (bias ? bias[c + 2 * k + 1] : 0.f));
        w[k] = *reinterpret_cast<uint32_t*>(&h);
      }
      *reinterpret_cast<uint4*>(y + i) = make_uint4(w[0], w[1], w[2], w[3]);
    } else {
      for (long long j = i; j < total && j < i + 8; ++j) {
        float a = 0.f;
        for (int sp = 0; sp < splits; ++sp) a += ws[sp * total + j];
        y[j] = __float2bfloat16_rn(a + (bias ? bias[j % Cout] : 0.f));
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// wgrad
// ---------------------------------------------------------------------------------------------------------------
struct WgradParams {
  int NF, OH, OW;                // dY spatial dims
  int Cout, Cin, taps;
  int tn, th, tw, tiles_w, tiles_h, num_ptiles;
  int KW, ph, pw;
  int BNc;                       // ci per CTA (multiple of 16, <= 128)
  int ci_tiles;
  int stages;
  int ptiles_per_split;
  int a_atoms, b_atoms;          // 64-channel atoms actually loaded for A (co) and B (ci)
  uint32_t atom_bytes;           // bytes one TMA box delivers (box pixels * 128)
  uint32_t ksteps;               // ceil(box pixels / 16)
  uint32_t tmem_cols;
  float* dw;                     // [Cout][Cin][taps] fp32, accumulated with red.add
  int dbg;
  int tm;                        // 1: the destination is TAP-MAJOR [taps][Cout][Cin] and is added to by TMA reduce (tmW)
};

__global__ void __launch_bounds__(192) conv_tc_wgrad_kernel(const __grid_constant__ CUtensorMap tmDY,
                                                            const __grid_constant__ CUtensorMap tmX,
                                                            const __grid_constant__ CUtensorMap tmW, const WgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int S = p.stages;
  const uint32_t a_stage = 2u * A_STAGE_BYTES;                       // co: 2 atoms of 64
  const uint32_t b_stage = (uint32_t)((p.BNc + 63) / 64) * A_STAGE_BYTES;  // ci atoms
  uint8_t* sA = smem;
  uint8_t* sB = smem + (size_t)S * a_stage;
  uint64_t* full = (uint64_t*)(sB + (size_t)S * b_stage);
  uint64_t* empty = full + S;
  uint64_t* accum_bar = empty + S;
  uint32_t* tmem_slot = (uint32_t*)(accum_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tap = blockIdx.z;
  const int kh = tap / p.KW, kw = tap % p.KW;
  const int co0 = (blockIdx.y / p.ci_tiles) * TILE_ROWS;
  const int ci0 = (blockIdx.y % p.ci_tiles) * p.BNc;
  const int pt_beg = blockIdx.x * p.ptiles_per_split;
  const int pt_end = min(pt_beg + p.ptiles_per_split, p.num_ptiles);
  const int iters = pt_end - pt_beg;

  // Rows a TMA box never writes (box pixels < 128, or an atom that is never loaded) must read as zero.
  {
    uint4 z = make_uint4(0, 0, 0, 0);
    uint4* ptr = (uint4*)smem;
    const size_t n16 = ((size_t)S * (a_stage + b_stage)) / 16;
    for (size_t i = threadIdx.x; i < n16; i += blockDim.x) ptr[i] = z;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 0 && lane == 0) {
    for (int i = 0; i < S; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(accum_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
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
        int tw_i, th_i, tn_i;                                        // tile coordinates, advanced without divisions
        { int t = pt_beg; tw_i = t % p.tiles_w; t /= p.tiles_w; th_i = t % p.tiles_h; tn_i = t / p.tiles_h; }
        for (int it = 0; it < iters; ++it) {
          const int ow0 = tw_i * p.tw, oh0 = th_i * p.th, n0 = tn_i * p.tn;
          if (++tw_i == p.tiles_w) { tw_i = 0; if (++th_i == p.tiles_h) { th_i = 0; ++tn_i; } }
          mbar_wait(&empty[stage], phase ^ 1);
          mbar_expect_tx(&full[stage], (uint32_t)(p.a_atoms + p.b_atoms) * p.atom_bytes);
          for (int a = 0; a < p.a_atoms; ++a)
            tma_load_4d(sA + (size_t)stage * a_stage + (size_t)a * A_STAGE_BYTES, &tmDY, &full[stage], co0 + a * KC, ow0, oh0, n0);
          for (int b = 0; b < p.b_atoms; ++b)
            tma_load_4d(sB + (size_t)stage * b_stage + (size_t)b * A_STAGE_BYTES, &tmX, &full[stage], ci0 + b * KC,
                        ow0 + kw - p.pw, oh0 + kh - p.ph, n0);
          if (++stage == S) { stage = 0; phase ^= 1; }
        }
      }
      __syncwarp();
    } else if (warp == 1) {
      if (lane == 0) {
        const uint32_t idesc = make_idesc(TILE_ROWS, p.BNc, 1, 1);
        int stage = 0; uint32_t phase = 0;
        for (int it = 0; it < iters; ++it) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t a0 = smem_u32(sA + (size_t)stage * a_stage);
          const uint32_t b0 = smem_u32(sB + (size_t)stage * b_stage);
          for (uint32_t k = 0; k < p.ksteps; ++k) {
            // MN-major SW128: 16 pixels (K) = 16 rows of 128 B; atoms of 64 channels are A_STAGE_BYTES apart (LBO)
            const uint64_t ad = make_desc(a0 + k * 2048, A_STAGE_BYTES, 1024);
            const uint64_t bd = make_desc(b0 + k * 2048, A_STAGE_BYTES, 1024);
            umma_bf16(tmem_base, ad, bd, idesc, (it | (int)k) != 0);
          }
          umma_commit(&empty[stage]);
          if (++stage == S) { stage = 0; phase ^= 1; }
        }
        umma_commit(accum_bar);
      }
      __syncwarp();
    } else {
      const int q = warp & 3;
      const int co = co0 + q * 32 + lane;
      mbar_wait(accum_bar, 0);
      tc_fence_after();
      if (p.tm) {
        // tap-major destination: stage the tile in the (now idle) pipeline buffers, add it with TMA reduce boxes
        for (int c = 0; c < p.BNc; c += 16) {
          float v[16];
          tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c, v);
          dw_stage16(smem + (size_t)(c >> 5) * 16384, q * 32 + lane, c & 16, v);
        }
        fence_async_smem();
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (threadIdx.x == 64 && !p.dbg) {
          for (int c = 0; c < p.BNc && ci0 + c < p.Cin; c += 32)
            for (int h = 0; h < 2 && co0 + 64 * h < p.Cout; ++h)
              tma_reduce_add_3d(&tmW, smem + (size_t)(c >> 5) * 16384 + h * 8192, ci0 + c, co0 + 64 * h, tap);
          bulk_commit();
          bulk_wait_all();
        }
      } else
      for (int c = 0; c < p.BNc; c += 16) {
        float v[16];
        tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c, v);
        if (co < p.Cout) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int ci = ci0 + c + i;
            if (ci < p.Cin && !(p.dbg && v[i] != 123.456f)) atomicAdd(p.dw + ((long long)co * p.Cin + ci) * p.taps + tap, v[i]);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, p.tmem_cols);
}

// ---------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------
// choose the pixel box (tn, th, tw), tn*th*tw <= 128, that minimises the number of tiles
void choose_box(int NF, int H, int W, int& tn, int& th, int& tw, int max_th = 1 << 30) {
  long long best = -1; tn = th = tw = 1;
  for (int w = 1; w <= W && w <= 128; ++w) {
    for (int h = 1; h <= H && h <= max_th && w * h <= 128; ++h) {
      int n = 128 / (w * h); if (n > NF) n = NF; if (n < 1) n = 1;
      long long tiles = (long long)((W + w - 1) / w) * ((H + h - 1) / h) * ((NF + n - 1) / n);
      if (best < 0 || tiles < best || (tiles == best && w > tw)) { best = tiles; tn = n; th = h; tw = w; }
    }
  }
}

// Split-K plan of the streaming kernel for small-grid, long-K problems (the 5x5 heads of the discriminators on 5x18
// maps: 32 CTAs x 200 stages): number of K splits (1 = none) for a grid of `ctas` CTAs and `taps` filter taps.
int plan_splits(long long ctas, int taps) {
  if (ctas * 2 > vca_num_sms() || taps < 8) return 1;
  int sp = (int)(vca_num_sms() / ctas);
  if (sp > 8) sp = 8;
  if (sp > taps / 2) sp = taps / 2;
  return sp < 2 ? 1 : sp;
}

int fwd_like(int NF, int IH, int IW, int Kdim, int OH, int OW, int Nout, int KH, int KW, int ph, int pw, int flip,
             const void* x, const void* wpk, const float* bias, void* y, cudaStream_t s, float* ws = nullptr,
             size_t ws_bytes = 0, size_t* ws_need = nullptr, double* stats = nullptr, const EpiExtra* ex = nullptr) {
  if (ws_need) *ws_need = 0;
  if (!ws_need) {
    int r = conv_ws_try(NF, IH, IW, Kdim, OH, OW, Nout, KH, KW, ph, pw, flip, x, wpk, bias, y, stats, ex, s);
    if (r == 0) r = conv_hs_try(NF, IH, IW, Kdim, OH, OW, Nout, KH, KW, ph, pw, flip, x, wpk, bias, y, stats, ex, s);
    if (r != 0) return r < 0 ? r : VCA_OK;
  }
  FwdParams p;
  p.NF = NF; p.OH = OH; p.OW = OW; p.Cout = Nout;
  p.IH = IH; p.IW = IW;
  // An input shorter than the filter (dgrad of the discriminator heads: dY is 1 x 14, dX 5 x 18) means every output
  // row sees a single filter row; one-row boxes let the kernel skip the other KH-1 (all-padding) taps.
  choose_box(NF, OH, OW, p.tn, p.th, p.tw, IH < KH ? 1 : 1 << 30);
  p.tiles_w = (OW + p.tw - 1) / p.tw; p.tiles_h = (OH + p.th - 1) / p.th;
  int tiles_n = (NF + p.tn - 1) / p.tn;
  p.KH = KH; p.KW = KW; p.ph = ph; p.pw = pw; p.flip = flip;
  p.kchunks = (Kdim + KC - 1) / KC;
  p.ksteps_last = (Kdim - (p.kchunks - 1) * KC + 15) / 16;
  int bn = Nout >= 256 ? 256 : ((Nout + 15) / 16) * 16;
  // keep >= ~1 wave of CTAs when the problem is small: halve BN
  long long ptiles = (long long)p.tiles_w * p.tiles_h * tiles_n;
  // (with enough filter taps and a workspace on offer, splitting K keeps the MMA N wide instead: a 64-wide tile caps
  //  the tensor pipe at 50 %)
  const int planned0 = plan_splits(ptiles * ((Nout + bn - 1) / bn), KH * KW);
  const bool can_split = (ws_need != nullptr || ws != nullptr) && planned0 > 1 &&
                         (ws_need != nullptr || ws_bytes >= (size_t)planned0 * NF * OH * OW * Nout * sizeof(float));
  while (!can_split && bn > 64 && ptiles * ((Nout + bn - 1) / bn) < vca_num_sms() && bn % 32 == 0) bn /= 2;
  p.BN = bn;
  p.a_bytes = (uint32_t)(p.tn * p.th * p.tw) * 128u;
  p.b_bytes = (uint32_t)bn * 128u;
  p.tmem_cols = pow2_cols(bn);
  p.bias = bias; p.y = (bf16*)y; p.stats = stats;
  p.has_ex = ex != nullptr;
  if (ex) p.ex = *ex; else p.ex = EpiExtra{nullptr, nullptr, 0.f, 0, 0.f, nullptr};
  const size_t stage_bytes = A_STAGE_BYTES + (size_t)bn * 128;
  // Two CTAs per SM (100 KB each) hide each other's TMA latency on big grids.  A grid that cannot even fill the SMs
  // once (small feature maps of the discriminator heads) gets one deep pipeline per CTA instead: a stage is only
  // 4 MMAs (~0.15-0.3 us) while a TMA round trip is ~1.5 us.
  long long total_ctas = ptiles * ((Nout + bn - 1) / bn);
  int splits = plan_splits(total_ctas, KH * KW);
  const size_t need = (size_t)splits * NF * OH * OW * Nout * sizeof(float);      // one fp32 slab per split
  if (ws_need) { *ws_need = splits > 1 ? need : 0; return VCA_OK; }   // planning query only
  if (splits > 1 && (!ws || ws_bytes < need)) splits = 1;
  const int taps_per_split = (KH * KW + splits - 1) / splits;
  splits = (KH * KW + taps_per_split - 1) / taps_per_split;
  p.splits = splits; p.taps_per_split = taps_per_split; p.ws = splits > 1 ? ws : nullptr;
  total_ctas *= splits;
  // multi-wave grids: two CTAs per SM (100 KB each) hide each other's pipeline fill / epilogue; tiles of <= 128 columns leave
  // TMEM room for THREE (72 KB each), of <= 64 for FOUR (52 KB) -- measured: 128-channel 5x5 at 20 x 75 600 -> 756 TFLOP/s,
  // 128 -> 64 5x5 at 40 x 150 533 -> 670; wider tiles unchanged
  const size_t budget = total_ctas <= vca_num_sms() ? 196 * 1024 : (size_t)(g_fwd_smem_kb != 100 ? g_fwd_smem_kb : bn <= 64 ? 52 : bn <= 128 ? 72 : 100) * 1024;
  int stages = (int)(budget / stage_bytes);
  const int max_stages = total_ctas <= vca_num_sms() ? 12 : 6;
  if (stages > max_stages) stages = max_stages; if (stages < 2) stages = 2;
  p.stages = stages;
  const size_t smem = stages * stage_bytes + 1024 + 256 + 3072;   // alignment + barriers + statistic accumulators / epilogue vectors

  CUtensorMap tmA, tmB;
  long long dA[4] = {Kdim, IW, IH, NF}; int bA[4] = {KC, p.tw, p.th, p.tn};
  long long dB[3] = {Kdim, Nout, (long long)KH * KW}; int bB[3] = {KC, bn, 1};
  int rc = make_map(&tmA, x, 4, dA, bA); if (rc) return rc;
  rc = make_map(&tmB, wpk, 3, dB, bB); if (rc) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(conv_tc_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) {
      vca_set_error("cudaFuncSetAttribute(conv_tc_fwd_kernel) failed"); return VCA_ERR_CUDA;
    }
    attr_set = true;
  }
  if (p.ws) {
    if (stats || ex) { vca_set_error("conv forward: statistics / fused epilogues are not available on the split-K path"); return VCA_ERR_UNSUPPORTED; }
    p.bias = nullptr;   // added by the finishing pass
  }
  dim3 grid((unsigned)ptiles, (unsigned)((Nout + bn - 1) / bn), (unsigned)splits);
  conv_tc_fwd_kernel<<<grid, 192, smem, s>>>(tmA, tmB, p);
  VCA_LAUNCH_CHECK();
  if (p.ws) {
    const long long total = (long long)NF * OH * OW * Nout;
    splitk_finish_kernel<<<vca_grid_1d(total, 256, 8), 256, 0, s>>>(ws, bias, (bf16*)y, total, Nout, splits);
    VCA_LAUNCH_CHECK();
  }
  return VCA_OK;
}

bool tc_geom_ok(const ConvGeom& g) {
  return g.ID == 1 && g.OD == 1 && g.KD == 1 && g.pd == 0 && g.sd == 1 && g.sh == 1 && g.sw == 1 && g.Cin % 8 == 0 &&
         g.Cout % 8 == 0 && g.KH * g.KW <= 65535 && g.OH == g.IH + 2 * g.ph - g.KH + 1 && g.OW == g.IW + 2 * g.pw - g.KW + 1 &&
         g.OH > 0 && g.OW > 0 && g.N > 0;
}

}  // namespace

extern "C" {

// kind: 0 forward, 1 dgrad, 2 wgrad.  1 when the tcgen05 path handles this geometry.
int vca_conv_tc_supported(const ConvGeom* g, int kind) {
  if (!g || !tc_geom_ok(*g)) return 0;
  if (kind == 0) return g->Cin >= 32;
  if (kind == 1) return g->Cout >= 32;
  if (kind == 2) return g->Cin >= 16 && g->Cout >= 16;
  return 0;
}

// Bytes of fp32 workspace with which vca_conv_{fwd,dgrad}_tc_ws would run this geometry split-K (0 = no split: the
// plain entry points do the same work).  kind 0 = forward, 1 = dgrad.
int vca_conv_tc_workspace(const ConvGeom* g, int kind) {
  if (!g || kind < 0 || kind > 1 || !vca_conv_tc_supported(g, kind)) return 0;
  size_t need = 0;
  if (kind == 0) fwd_like(g->N, g->IH, g->IW, g->Cin, g->OH, g->OW, g->Cout, g->KH, g->KW, g->ph, g->pw, 0, nullptr, nullptr, nullptr,
                          nullptr, 0, nullptr, 0, &need);
  else fwd_like(g->N, g->OH, g->OW, g->Cout, g->IH, g->IW, g->Cin, g->KH, g->KW, g->KH - 1 - g->ph, g->KW - 1 - g->pw, 1, nullptr,
                nullptr, nullptr, nullptr, 0, nullptr, 0, &need);
  return need > 0x7fffffff ? 0 : (int)need;
}
// x [N,IH,IW,Cin] bf16; wd = packed [taps][Cout][Cin] bf16 (vca_pack_conv_weight "wd"); y [N,OH,OW,Cout] bf16.
// ws / ws_bytes: optional fp32 workspace (see vca_conv_tc_workspace); null = never split.
int vca_conv_fwd_tc_ws(const ConvGeom* g, const void* x, const void* wd, const float* bias, void* y, float* ws, long long ws_bytes,
                       cudaStream_t s) {
  VCA_CHECK_ARG(g && x && wd && y && ws_bytes >= 0 && vca_conv_tc_supported(g, 0));
  return fwd_like(g->N, g->IH, g->IW, g->Cin, g->OH, g->OW, g->Cout, g->KH, g->KW, g->ph, g->pw, 0, x, wd, bias, y, s, ws,
                  (size_t)ws_bytes);
}
int vca_conv_fwd_tc(const ConvGeom* g, const void* x, const void* wd, const float* bias, void* y, cudaStream_t s) {
  return vca_conv_fwd_tc_ws(g, x, wd, bias, y, nullptr, 0, s);
}
// Inference forward with a fused epilogue:  y = act(conv(x, w) * scale[c] + shift[c] + res * res_scale)   (never split-K).
//   scale / shift: per-output-channel fp32 (either may be null: 1 / 0) -- an eval-mode BatchNorm folded into the conv;
//   res: bf16 tensor of y's shape or null;  act: 0 none, 1 LeakyReLU(slope), 2 PReLU(prelu_w[Cout]), 3 ReLU.
int vca_conv_fwd_tc_epi(const ConvGeom* g, const void* x, const void* wd, const float* scale, const float* shift, const void* res,
                        float res_scale, int act, float slope, const float* prelu_w, void* y, cudaStream_t s) {
  VCA_CHECK_ARG(g && x && wd && y && vca_conv_tc_supported(g, 0) && act >= 0 && act <= 3 && (act != 2 || prelu_w));
  VCA_CHECK_ARG(!res || (g->Cout % 8 == 0 && (reinterpret_cast<uintptr_t>(res) & 15) == 0));
  EpiExtra ex{scale, (const bf16*)res, res_scale, act, slope, prelu_w};
  return fwd_like(g->N, g->IH, g->IW, g->Cin, g->OH, g->OW, g->Cout, g->KH, g->KW, g->ph, g->pw, 0, x, wd, shift, y, s, nullptr, 0, nullptr,
                  nullptr, &ex);
}
// Inference forward of a convolution whose eval-mode BatchNorm has been folded into the weights (scale) and the bias (shift):
//     y = a(conv(x, wd) + shift),   a(v) = v > 0 ? v : v * slope[c]     (slope per output channel: LeakyReLU / PReLU / ReLU = 0)
// Runs only where the weights-stationary kernel takes the geometry (its stacked MMAs and TMA-store epilogue are untouched:
// the activation is one select per value on the accumulator row); VCA_ERR_UNSUPPORTED otherwise -- the caller then keeps the
// separate normalisation pass.  1 / 0 from the _supported query.
int vca_conv_fwd_tc_act_supported(const ConvGeom* g) {
  if (!g || !vca_conv_tc_supported(g, 0)) return 0;
  return conv_ws_try(g->N, g->IH, g->IW, g->Cin, g->OH, g->OW, g->Cout, g->KH, g->KW, g->ph, g->pw, 0, nullptr, nullptr, nullptr, nullptr,
                     nullptr, nullptr, nullptr) == 1 ? 1 : 0;
}
int vca_conv_fwd_tc_act(const ConvGeom* g, const void* x, const void* wd, const float* shift, const float* slope, void* y, cudaStream_t s) {
  VCA_CHECK_ARG(g && x && wd && y && slope && vca_conv_tc_supported(g, 0));
  const int r = conv_ws_try(g->N, g->IH, g->IW, g->Cin, g->OH, g->OW, g->Cout, g->KH, g->KW, g->ph, g->pw, 0, x, wd, shift, y, nullptr, nullptr, s, slope);
  if (r == 0) { vca_set_error("vca_conv_fwd_tc_act: the weights-stationary kernel does not take this geometry"); return VCA_ERR_UNSUPPORTED; }
  return r < 0 ? r : VCA_OK;
}
// 1 when vca_conv_fwd_tc_stats takes this geometry AND the statistics come (almost) for free: the weights-stationary
// persistent kernel (<= 64 channels in and out: the stem, ResNet layer 1, the 40x150 / 80x300 generator stages -- the
// large tensors, where the separate statistics pass costs most).  The streaming / halo-resident kernels can emit them too
// (measured: their one-tile-at-a-time epilogues pay more for the column reduction than the saved pass is worth).
int vca_conv_fwd_tc_stats_supported(const ConvGeom* g) {
  if (!g || !vca_conv_tc_supported(g, 0)) return 0;
  static double dummy;
  return conv_ws_try(g->N, g->IH, g->IW, g->Cin, g->OH, g->OW, g->Cout, g->KH, g->KW, g->ph, g->pw, 0, nullptr, nullptr, nullptr, nullptr,
                     &dummy, nullptr, nullptr) == 1 ? 1 : 0;
}
// Forward convolution that also accumulates the per-output-channel sum and sum of squares of y (as stored, i.e. bf16
// rounded) into stats[0 .. Cout) and stats[Cout .. 2 Cout) (fp64, ADDED to): the batch statistics of a BatchNorm that
// follows (vca_bn_finalize_stats turns them into mean / invstd).  Never runs split-K.
int vca_conv_fwd_tc_stats(const ConvGeom* g, const void* x, const void* wd, const float* bias, void* y, double* stats, cudaStream_t s) {
  VCA_CHECK_ARG(g && x && wd && y && stats && vca_conv_tc_supported(g, 0));
  return fwd_like(g->N, g->IH, g->IW, g->Cin, g->OH, g->OW, g->Cout, g->KH, g->KW, g->ph, g->pw, 0, x, wd, bias, y, s, nullptr, 0, nullptr,
                  stats);
}
// dy [N,OH,OW,Cout] bf16; wf = packed [taps][Cin][Cout] bf16 (vca_pack_conv_weight "wf"); dx [N,IH,IW,Cin] bf16.
int vca_conv_dgrad_tc_ws(const ConvGeom* g, const void* dy, const void* wf, void* dx, float* ws, long long ws_bytes, cudaStream_t s) {
  VCA_CHECK_ARG(g && dy && wf && dx && ws_bytes >= 0 && vca_conv_tc_supported(g, 1));
  return fwd_like(g->N, g->OH, g->OW, g->Cout, g->IH, g->IW, g->Cin, g->KH, g->KW, g->KH - 1 - g->ph, g->KW - 1 - g->pw, 1, dy, wf,
                  nullptr, dx, s, ws, (size_t)ws_bytes);
}
int vca_conv_dgrad_tc(const ConvGeom* g, const void* dy, const void* wf, void* dx, cudaStream_t s) {
  return vca_conv_dgrad_tc_ws(g, dy, wf, dx, nullptr, 0, s);
}
static int wgrad_tc(const ConvGeom* g, const void* dy, const void* x, float* dw, int tm, cudaStream_t s) {
  VCA_CHECK_ARG(g && dy && x && dw && vca_conv_tc_supported(g, 2));
  VCA_CHECK_ARG(!tm || (g->Cin % 4 == 0 && (reinterpret_cast<uintptr_t>(dw) & 15) == 0));
  {
    const int r = conv_wgrad_ws_try(*g, dy, x, dw, s, tm);   // multi-tap halo-resident kernel for <= 64 input channels
    if (r != 0) return r < 0 ? r : VCA_OK;
  }
  WgradParams p;
  p.NF = g->N; p.OH = g->OH; p.OW = g->OW; p.Cout = g->Cout; p.Cin = g->Cin; p.taps = g->KH * g->KW;
  choose_box(g->N, g->OH, g->OW, p.tn, p.th, p.tw);
  p.tiles_w = (g->OW + p.tw - 1) / p.tw; p.tiles_h = (g->OH + p.th - 1) / p.th;
  p.num_ptiles = p.tiles_w * p.tiles_h * ((g->N + p.tn - 1) / p.tn);
  p.KW = g->KW; p.ph = g->ph; p.pw = g->pw;
  // ci per CTA: the kernel is L2->SM bound, and a 256-wide tile re-reads the dY atoms half as often (12 instead of 16
  // atoms per pixel tile at Cin = 512); taken when it does not pad Cin by more than a fifth.
  p.BNc = g->Cin >= 128 ? 128 : ((g->Cin + 15) / 16) * 16;
  if (g->Cin >= 256 && ((g->Cin + 255) / 256) * 256 * 5 <= g->Cin * 6) p.BNc = 256;
  p.ci_tiles = (g->Cin + p.BNc - 1) / p.BNc;
  const int co_tiles = (g->Cout + TILE_ROWS - 1) / TILE_ROWS;
  p.a_atoms = g->Cout >= TILE_ROWS ? 2 : (g->Cout + KC - 1) / KC;   // per co tile; tail tiles rely on TMA zero fill
  if (p.a_atoms > 2) p.a_atoms = 2;
  p.b_atoms = (p.BNc + KC - 1) / KC;
  const int box_px = p.tn * p.th * p.tw;
  p.atom_bytes = (uint32_t)box_px * 128u;
  p.ksteps = (uint32_t)((box_px + 15) / 16);
  p.tmem_cols = pow2_cols(p.BNc);
  p.dw = dw; p.dbg = g_wg_dbg; p.tm = tm;
  const size_t stage_bytes = 2 * A_STAGE_BYTES + (size_t)p.b_atoms * A_STAGE_BYTES;
  int stages = (int)(((size_t)g_wg_smem_kb * 1024) / stage_bytes);
  if (stages > 4) stages = 4; if (stages < 2) stages = 2;
  p.stages = stages;
  const size_t smem = stages * stage_bytes + 1024 + 256;
  const long long base_ctas = (long long)co_tiles * p.ci_tiles * p.taps;
  int split = (int)((2ll * vca_num_sms() + base_ctas - 1) / base_ctas);
  if (split < 1) split = 1; if (split > p.num_ptiles) split = p.num_ptiles;
  p.ptiles_per_split = (p.num_ptiles + split - 1) / split;
  split = (p.num_ptiles + p.ptiles_per_split - 1) / p.ptiles_per_split;

  CUtensorMap tmDY, tmX, tmW;
  long long dY[4] = {g->Cout, g->OW, g->OH, g->N}; int bx[4] = {KC, p.tw, p.th, p.tn};
  long long dX[4] = {g->Cin, g->IW, g->IH, g->N};
  int rc = make_map(&tmDY, dy, 4, dY, bx); if (rc) return rc;
  rc = make_map(&tmX, x, 4, dX, bx); if (rc) return rc;
  if (tm) {
    if ((size_t)stages * stage_bytes < (size_t)((p.BNc + 31) / 32) * 16384) { vca_set_error("wgrad: staging does not fit"); return VCA_ERR_UNSUPPORTED; }
    rc = make_map_dw(&tmW, dw, g->Cin, g->Cout, p.taps); if (rc) return rc;
  } else tmW = tmX;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(conv_tc_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024) != cudaSuccess) {
      vca_set_error("cudaFuncSetAttribute(conv_tc_wgrad_kernel) failed"); return VCA_ERR_CUDA;
    }
    attr_set = true;
  }
  VCA_CHECK_ARG((long long)co_tiles * p.ci_tiles <= 65535 && p.taps <= 65535);
  dim3 grid((unsigned)split, (unsigned)(co_tiles * p.ci_tiles), (unsigned)p.taps);
  conv_tc_wgrad_kernel<<<grid, 192, smem, s>>>(tmDY, tmX, tmW, p);
  VCA_LAUNCH_CHECK();
  return VCA_OK;
}
// dw fp32 [Cout][Cin][taps] (the parameter layout), ADDED to with red.add across pixel splits.
int vca_conv_wgrad_tc(const ConvGeom* g, const void* dy, const void* x, float* dw, cudaStream_t s) { return wgrad_tc(g, dy, x, dw, 0, s); }
// Same gradient ADDED to a TAP-MAJOR fp32 tensor [taps][Cout][Cin] (16-byte aligned, Cin % 4 == 0) through shared memory and
// TMA reduce-add boxes -- full-line adds in L2 instead of one scattered 4-byte atomic per element.  For a pointwise
// conv / linear layer (taps = 1) the two layouts coincide; otherwise vca_grad_unslab_batched folds the slabs of an
// optimizer group back into the parameter layout in one launch.
int vca_conv_wgrad_tc_tm(const ConvGeom* g, const void* dy, const void* x, float* dw_tm, cudaStream_t s) { return wgrad_tc(g, dy, x, dw_tm, 1, s); }

}  // extern "C"
