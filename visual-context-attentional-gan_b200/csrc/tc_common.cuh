// PTX wrappers shared by the tcgen05 / TMEM / TMA kernels (sm_100a): mbarrier, TMA tensor loads, TMEM alloc,
// tcgen05.mma / commit / ld, shared-memory matrix descriptors, instruction descriptors, host-side tensor maps.
#pragma once
#include "common.cuh"
#include <cuda.h>

namespace tc {

constexpr uint32_t SPIN_LIMIT = 1u << 24;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0, ok = 0;
  const uint32_t addr = smem_u32(bar);
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
    if (ok) break;
    if (++spins > SPIN_LIMIT) __trap();  // a lost arrival must fail loudly, never hang the GPU
  }
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// Shared-memory matrix descriptor (SM100 "version 1"), SWIZZLE_128B.
//   bits [0,14) start address >> 4 | [16,30) leading-dim byte offset >> 4 | [32,46) stride-dim byte offset >> 4
//   bits [46,48) version = 1 | [49,52) base offset (row phase of a start address that is not 1024-B aligned)
//   bits [61,64) layout type (2 = SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t base_off = 0) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(base_off & 7) << 49;
  d |= (uint64_t)2 << 61;
  return d;
}
// Instruction descriptor, kind::f16: D = f32, A = B = bf16; a_major bit 15, b_major bit 16 (0 = K-major, 1 = MN-major),
// N>>3 at bits [17,23), M>>4 at bits [24,29).
__host__ __device__ inline uint32_t make_idesc(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}

// ---- BatchNorm batch statistics fused into the convolution epilogues ---------------------------------------------------
// The conv kernels' epilogue threads hold one output pixel (tile row) per thread and 16 output channels at a time.  The
// per-channel sum and sum of squares over the tile's rows are taken with a butterfly of warp shuffles (16 shuffles per
// 16-column chunk and quantity instead of 16 x 5), added into shared accumulators (4 epilogue warps), and flushed with
// one fp64 atomic per channel and tile into the BatchNorm's [2 * Cout] scratch -- the buffer bn_finalize_kernel reads.
// Statistics are taken of the bf16-ROUNDED values, i.e. of exactly the tensor the normalisation pass will read.
__device__ __forceinline__ float bf16_round(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }
// after the call, a[0] of lane L is the sum over all 32 lanes of column epi_col(L) (lanes 2k and 2k+1 hold the same column)
__device__ __forceinline__ int epi_col(int lane) { return ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1); }
__device__ __forceinline__ float colsum16(float (&a)[16], int lane) {
  const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4, b1 = lane & 2;
#pragma unroll
  for (int i = 0; i < 8; ++i) { const float send = b4 ? a[i] : a[i + 8], keep = b4 ? a[i + 8] : a[i]; a[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16); }
#pragma unroll
  for (int i = 0; i < 4; ++i) { const float send = b3 ? a[i] : a[i + 4], keep = b3 ? a[i + 4] : a[i]; a[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8); }
#pragma unroll
  for (int i = 0; i < 2; ++i) { const float send = b2 ? a[i] : a[i + 2], keep = b2 ? a[i + 2] : a[i]; a[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4); }
  { const float send = b1 ? a[0] : a[1], keep = b1 ? a[1] : a[0]; a[0] = keep + __shfl_xor_sync(0xffffffffu, send, 2); }
  a[0] += __shfl_xor_sync(0xffffffffu, a[0], 1);
  return a[0];
}
// v: the 16 (bias-added) outputs of this thread's row for channels [co, co+16); s_sum / s_sq point at the chunk's 16 slots.
// Must be called by all 32 lanes of the warp (rows that do not exist pass row_ok = false).
__device__ __forceinline__ void epi_stats16(const float (&v)[16], bool row_ok, int co, int Cout, float* s_sum, float* s_sq, int lane) {
  float a[16], b[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) { const float r = row_ok ? bf16_round(v[i]) : 0.f; a[i] = r; b[i] = r * r; }
  const float sa = colsum16(a, lane), sb = colsum16(b, lane);
  const int col = epi_col(lane);
  if (!(lane & 1) && co + col < Cout) { atomicAdd(s_sum + col, sa); atomicAdd(s_sq + col, sb); }
}
// All epilogue threads (tid128 = 0..nthr-1): publish the tile's sums and leave the shared accumulators zeroed.
__device__ __forceinline__ void epi_stats_flush(float* s_sum, float* s_sq, int BN, int co0, int Cout, double* sums, int tid128, int nthr = 128) {
  asm volatile("bar.sync 1, %0;" ::"r"(nthr) : "memory");
  for (int c = tid128; c < BN; c += nthr) {
    if (co0 + c < Cout) { atomicAdd(sums + co0 + c, (double)s_sum[c]); atomicAdd(sums + Cout + co0 + c, (double)s_sq[c]); }
    s_sum[c] = 0.f; s_sq[c] = 0.f;
  }
  asm volatile("bar.sync 1, %0;" ::"r"(nthr) : "memory");
}

// ---- inference epilogue: y = act(acc * scale[c] + shift[c] + res * res_scale) ------------------------------------------
// Eval-mode BatchNorm is a fixed per-channel affine map: it folds into the producing convolution as
// scale / shift, the activation that follows and the residual add of a BasicBlock / the shortcut + 1/sqrt(2) of a GenResBlk
// run on the accumulator row while it is in registers -- the separate normalisation / activation / add passes disappear.
struct EpiExtra {
  const float* scale;      // [Cout] or null (1)
  const bf16* res;         // [pixels][Cout] (the output's layout) or null
  float res_scale;
  int act;                 // 0 none, 1 LeakyReLU(slope), 2 PReLU(prelu_w[c]), 3 ReLU
  float slope;
  const float* prelu_w;    // [Cout] (act == 2)
};
// v: accumulator values of channels [co, co + 16) of one output pixel; `shift` is the conv's bias pointer (or null);
// res_row = res + pixel * Cout + co (dereferenced only when row_ok)
__device__ __forceinline__ void epi_apply16(float (&v)[16], const EpiExtra& e, const float* shift, int co, int Cout, const bf16* res_row,
                                            bool row_ok) {
  if (co >= Cout) return;
  const bool full = co + 16 <= Cout;
  if (e.scale) {
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] *= (full || co + i < Cout) ? __ldg(e.scale + co + i) : 1.f;
  }
  if (shift) {
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] += (full || co + i < Cout) ? __ldg(shift + co + i) : 0.f;
  }
  if (e.res && row_ok) {
    if (full) {
      const uint4 r0 = *reinterpret_cast<const uint4*>(res_row), r1 = *reinterpret_cast<const uint4*>(res_row + 8);
      const uint32_t w[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        v[2 * i] = fmaf(__uint_as_float(w[i] << 16), e.res_scale, v[2 * i]);
        v[2 * i + 1] = fmaf(__uint_as_float(w[i] & 0xffff0000u), e.res_scale, v[2 * i + 1]);
      }
    } else {
      for (int i = 0; i < 16 && co + i < Cout; ++i) v[i] = fmaf(__bfloat162float(res_row[i]), e.res_scale, v[i]);
    }
  }
  if (e.act == 1) {
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = v[i] > 0.f ? v[i] : v[i] * e.slope;
  } else if (e.act == 2) {
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = v[i] > 0.f ? v[i] : v[i] * ((full || co + i < Cout) ? __ldg(e.prelu_w + co + i) : 0.f);
  } else if (e.act == 3) {
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.f);
  }
}

// Per-channel epilogue vectors staged in shared memory (s_epi = [3][BN] floats: scale, shift, negative-side slope): the
// epilogue threads read them with broadcast LDS instead of dependent global loads -- with ~200 KB of the SM's memory
// configured as shared there is next to no L1 left, and a global load per channel and chunk made the epilogue, not the
// MMAs, the critical path of the short-K layers (measured: 125 instead of 550 TFLOP/s on ResNet layer 1).
// Called by all threads of the CTA before the first __syncthreads().
__device__ __forceinline__ void epi_stage(float* s_epi, int BN, int co0, int Cout, const EpiExtra& e, const float* shift) {
  for (int i = threadIdx.x; i < BN; i += blockDim.x) {
    const int c = co0 + i;
    const bool ok = c < Cout;
    s_epi[i] = (e.scale && ok) ? e.scale[c] : 1.f;
    s_epi[BN + i] = (shift && ok) ? shift[c] : 0.f;
    s_epi[2 * BN + i] = e.act == 0 ? 1.f : (e.act == 1 ? e.slope : (e.act == 2 ? (ok ? e.prelu_w[c] : 0.f) : 0.f));
  }
}
// One 64-column group of the inference epilogue for this thread's output row.  The residual's 128 bytes are requested with
// eight back-to-back 16-byte loads BEFORE the accumulator columns are read, so their latency is paid once per group.
// taddr = TMEM address of the group's first column (this thread's lane); co = output channel of that column; bn_left =
// columns of the CTA's tile from there on; se = s_epi + (column offset of the group inside the tile); BN = tile width.
__device__ __forceinline__ void epi_group64(uint32_t taddr, const float* se, int BN, float res_scale, int co, int bn_left, int Cout,
                                            bf16* ydst, const bf16* rdst, bool row_ok) {
  uint4 rr[8];
  const bool res_vec = rdst != nullptr && row_ok && co + 64 <= Cout && bn_left >= 64;
  if (res_vec) {
#pragma unroll
    for (int j = 0; j < 8; ++j) rr[j] = *reinterpret_cast<const uint4*>(rdst + j * 8);
  }
#pragma unroll
  for (int cc = 0; cc < 4; ++cc) {
    const int c = cc * 16;
    if (c < bn_left) {
      float v[16];
      tmem_ld16(taddr + (uint32_t)c, v);
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = fmaf(v[i], se[c + i], se[BN + c + i]);
      if (res_vec) {
        const uint4 r0 = rr[2 * cc], r1 = rr[2 * cc + 1];
        const uint32_t w8[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          v[2 * i] = fmaf(__uint_as_float(w8[i] << 16), res_scale, v[2 * i]);
          v[2 * i + 1] = fmaf(__uint_as_float(w8[i] & 0xffff0000u), res_scale, v[2 * i + 1]);
        }
      } else if (rdst != nullptr && row_ok) {
        for (int i = 0; i < 16 && co + c + i < Cout; ++i) v[i] = fmaf(__bfloat162float(rdst[c + i]), res_scale, v[i]);
      }
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = v[i] > 0.f ? v[i] : v[i] * se[2 * BN + c + i];
      if (row_ok && co + c < Cout) {
        if (co + c + 16 <= Cout) {
          uint32_t w[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) { __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]); w[i] = *reinterpret_cast<uint32_t*>(&h); }
          *reinterpret_cast<uint4*>(ydst + c) = make_uint4(w[0], w[1], w[2], w[3]);
          *reinterpret_cast<uint4*>(ydst + c + 8) = make_uint4(w[4], w[5], w[6], w[7]);
        } else {
          for (int i = 0; i < 16 && co + c + i < Cout; ++i) ydst[c + i] = __float2bfloat16_rn(v[i]);
        }
      }
    }
  }
}

// Residual rows are read by the epilogue threads with a dependent load per 16-column chunk; issued only after the
// accumulator is complete, their DRAM latency would sit on the critical path.  Called BEFORE waiting for the accumulator, this
// pulls the row (nbytes, from 128-byte aligned-ish `row`) into L2 while the MMAs are still running.
__device__ __forceinline__ void epi_prefetch_row(const bf16* row, int nbytes) {
  const char* p = reinterpret_cast<const char*>(row);
  for (int o = 0; o < nbytes; o += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(p + o));
}

// ---- weight-gradient epilogue through shared memory + TMA reduce-add (tap-major destination) --------------------------
// The accumulator tile (rows = output channels, columns = input channels of ONE filter tap) is staged as fp32 in blocks of
// 32 columns: block b = [128 rows][128 B], 16-byte chunks XOR-swizzled with (row & 7) -- the SWIZZLE_128B image of two
// {32 ci, 64 co, 1 tap} boxes -- and added to global memory by cp.reduce.async.bulk.tensor (full-line fp32 adds performed
// in L2) instead of one scattered 4-byte red.add per element: measured 15-40 % of a wgrad kernel's time on the big layers.
__device__ __forceinline__ void dw_stage16(uint8_t* blk, int row, int c16, const float (&v)[16]) {
  uint8_t* r = blk + row * 128;
#pragma unroll
  for (int i = 0; i < 4; ++i)
    *reinterpret_cast<float4*>(r + ((((c16 >> 2) + i) ^ (row & 7)) << 4)) = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
}
__device__ __forceinline__ void tma_reduce_add_3d(const CUtensorMap* map, const void* src, int c0, int c1, int c2) {
  asm volatile("cp.reduce.async.bulk.tensor.3d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"((uint64_t)map), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- host ----
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static inline EncodeTiledFn get_encode() {
  // The driver entry point needs the primary context bound to THIS thread.  A fresh thread that has made no context-binding
  // runtime call yet -- autograd's backward worker when one of our kernels is the first thing it runs -- otherwise gets
  // CUDA_ERROR_INVALID_CONTEXT (201) from cuTensorMapEncodeTiled.
  static thread_local bool ctx_bound = false;
  if (!ctx_bound) { cudaFree(nullptr); ctx_bound = true; }
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}
// bf16 tensor map (default: 64-channel = 128 B inner box, SWIZZLE_128B), zero OOB fill; dims/box innermost-first.
static inline int make_map(CUtensorMap* m, const void* base, int rank, const long long* dims, const int* box,
                           CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_128B) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { vca_set_error("cuTensorMapEncodeTiled entry point unavailable"); return VCA_ERR_CUDA; }
  cuuint64_t gd[5]; cuuint64_t gs[4]; cuuint32_t bx[5]; cuuint32_t es[5];
  long long stride = 2;
  for (int i = 0; i < rank; ++i) {
    gd[i] = (cuuint64_t)dims[i]; bx[i] = (cuuint32_t)box[i]; es[i] = 1;
    if (i > 0) gs[i - 1] = (cuuint64_t)stride;
    stride *= dims[i];
  }
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gd, gs, bx, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { vca_set_error("cuTensorMapEncodeTiled failed (CUresult %d)", (int)r); return VCA_ERR_CUDA; }
  return VCA_OK;
}
// fp32 tensor map of a weight gradient in TAP-MAJOR order [taps][Cout][Cin]: boxes of 32 input channels (128 B, SWIZZLE_128B)
// x 64 output channels, the unit of the TMA reduce-add epilogue of the wgrad kernels
static inline int make_map_dw(CUtensorMap* m, const float* base, int Cin, int Cout, int taps) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { vca_set_error("cuTensorMapEncodeTiled entry point unavailable"); return VCA_ERR_CUDA; }
  cuuint64_t gd[3] = {(cuuint64_t)Cin, (cuuint64_t)Cout, (cuuint64_t)taps};
  cuuint64_t gs[2] = {(cuuint64_t)Cin * 4, (cuuint64_t)Cin * Cout * 4};
  cuuint32_t bx[3] = {32, 64, 1}, es[3] = {1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { vca_set_error("cuTensorMapEncodeTiled(dw) failed (CUresult %d)", (int)r); return VCA_ERR_CUDA; }
  return VCA_OK;
}
static inline uint32_t pow2_cols(int n) { uint32_t c = 32; while ((int)c < n) c <<= 1; return c; }

}  // namespace tc
