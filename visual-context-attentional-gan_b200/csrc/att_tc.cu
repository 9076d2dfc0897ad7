// Visual-context attention (generator.py:154-171) as ONE tcgen05 kernel per (batch entry, 128-query tile):
//
//     S = Q K^T / sqrt(d)   ->   key mask  att[i, :, len[i]:] = -inf   ->   P = softmax_keys(S)   ->   O = P V
//
//   phase 1   Q tile [128 x 256] and K [SP x 256] arrive by TMA in four 64-channel chunks (SWIZZLE_128B, K-major);
//             16 tcgen05.mma 128 x SP x 16 accumulate the scores in TMEM columns [0, SP).
//   softmax   4 epilogue warps, one query row per thread: tcgen05.ld the row, scale, mask keys >= len, max, exp2, sum
//             (fp32, in registers); the normalised probabilities are written as bf16 (a) to shared memory in the
//             K-major SWIZZLE_128B layout a TMA load would have produced -- they are the A operand of phase 2 -- and
//             (b) to global memory for the backward pass.
//   phase 2   V [SP x 256] arrives by TMA into the space K occupied (the MN-major B operand: keys = K dimension, 64-channel
//             atoms SP*128 B apart); SP/16 tcgen05.mma 128 x 256 x 16 accumulate O in TMEM columns [256, 512).
//   epilogue  O -> bf16 -> global, channels-last rows.
//
// S <= 256 keys (GRID: 75; LRS: <= 250 -- SURVEY section 5); SP = S rounded up to 16; d = 256.  Rows / keys beyond the
// extents of a batch entry are zero-filled by the TMA unit (the batch is the outermost tensor-map dimension).
// Shared memory: 64 KB (Q chunks, then P) + 128 KB (K chunks, then V).  TMEM: all 512 columns.
#include "tc_common.cuh"

using namespace tc;

namespace {

constexpr int D = 256;                 // out_dim of AVAttention (generator.py:349,351)
constexpr int QCHUNK = 128 * 128;      // one 64-channel chunk of the Q tile / one 64-key chunk of the P tile

struct AttParams {
  int Tq, S, SP;
  float scale_log2;                    // (1/sqrt(d)) * log2(e)
  const int* lens;                     // [B] valid keys per batch entry (clamped to [0, S])
  bf16* O;                             // [B][Tq][256]
  bf16* P;                             // [B][Tq][SP]
};

__global__ void __launch_bounds__(192, 1) att_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                                                            const __grid_constant__ CUtensorMap tmV, const AttParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sQ = smem;                                  // phase 1: 4 x [128][128 B];  phase 2: P, 64-key chunks [128][128 B]
  uint8_t* sK = smem + 4 * QCHUNK;                     // phase 1: 4 x [SP][128 B];   phase 2: V, 4 x [SP][128 B] (64-channel atoms)
  uint64_t* bars = (uint64_t*)(sK + 4 * 256 * 128);
  uint64_t* qk_full = bars;                            // [4]
  uint64_t* s_done = bars + 4;                         // scores complete (tcgen05.commit)
  uint64_t* v_full = bars + 5;
  uint64_t* p_ready = bars + 6;                        // 128 epilogue threads have written P
  uint64_t* o_done = bars + 7;
  uint32_t* tmem_slot = (uint32_t*)(bars + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 128, b = blockIdx.y;
  const uint32_t kchunk_bytes = (uint32_t)p.SP * 128u;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < 4; ++i) mbar_init(&qk_full[i], 1);
    mbar_init(s_done, 1); mbar_init(v_full, 1); mbar_init(p_ready, 128); mbar_init(o_done, 1);
    mbar_init_fence();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      for (int c = 0; c < 4; ++c) {
        mbar_expect_tx(&qk_full[c], (uint32_t)QCHUNK + kchunk_bytes);
        tma_load_3d(sQ + (size_t)c * QCHUNK, &tmQ, &qk_full[c], c * 64, q0, b);
        tma_load_3d(sK + (size_t)c * kchunk_bytes, &tmK, &qk_full[c], c * 64, 0, b);
      }
      mbar_wait(s_done, 0);                            // the score MMAs have consumed Q and K: their space is free
      mbar_expect_tx(v_full, 4u * kchunk_bytes);
      for (int t = 0; t < 4; ++t) tma_load_3d(sK + (size_t)t * kchunk_bytes, &tmV, v_full, t * 64, 0, b);
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc1 = make_idesc(128, p.SP, 0, 0);
      for (int c = 0; c < 4; ++c) {
        mbar_wait(&qk_full[c], 0);
        tc_fence_after();
        const uint32_t a0 = smem_u32(sQ + (size_t)c * QCHUNK), b0 = smem_u32(sK + (size_t)c * kchunk_bytes);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16(tmem_base, make_desc(a0 + k * 32, 0, 1024), make_desc(b0 + k * 32, 0, 1024), idesc1, (c | k) != 0);
      }
      umma_commit(s_done);
      // phase 2: O = P V
      mbar_wait(v_full, 0);
      mbar_wait(p_ready, 0);
      tc_fence_after();
      const uint32_t idesc2 = make_idesc(128, D, 0, 1);
      const uint32_t p0 = smem_u32(sQ), v0 = smem_u32(sK);
      const int ksteps = p.SP / 16;
      for (int kk = 0; kk < ksteps; ++kk) {
        const uint64_t ad = make_desc(p0 + (uint32_t)(kk >> 2) * QCHUNK + (uint32_t)(kk & 3) * 32, 0, 1024);
        const uint64_t bd = make_desc(v0 + (uint32_t)kk * 2048, kchunk_bytes, 1024);
        umma_bf16(tmem_base + 256, ad, bd, idesc2, kk != 0);
      }
      umma_commit(o_done);
    }
    __syncwarp();
  } else {
    const int q = warp & 3;
    const int r = q * 32 + lane;                       // query row of the tile = TMEM lane
    const int tq = q0 + r;
    const bool row_ok = tq < p.Tq;
    int len = p.lens[b];
    len = len < 0 ? 0 : (len > p.S ? p.S : len);
    const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
    mbar_wait(s_done, 0);
    tc_fence_after();
    // pass 1: row maximum over the valid keys
    float mx = -INFINITY;
    for (int c = 0; c < p.SP; c += 16) {
      float v[16];
      tmem_ld16(trow + (uint32_t)c, v);
#pragma unroll
      for (int i = 0; i < 16; ++i)
        if (c + i < len) mx = fmaxf(mx, v[i] * p.scale_log2);
    }
    // pass 2: sum of exponentials
    float sum = 0.f;
    for (int c = 0; c < p.SP; c += 16) {
      float v[16];
      tmem_ld16(trow + (uint32_t)c, v);
#pragma unroll
      for (int i = 0; i < 16; ++i)
        if (c + i < len) sum += exp2f(v[i] * p.scale_log2 - mx);
    }
    const float inv = 1.f / sum;                       // len == 0: 1/0 * exp2(nan) -> NaN rows, as the reference produces
    // pass 3: probabilities -> bf16 -> shared memory (A operand of P V) and global memory (saved for backward)
    bf16* prow = p.P + ((long long)b * p.Tq + tq) * p.SP;
    for (int c = 0; c < p.SP; c += 16) {
      float v[16];
      tmem_ld16(trow + (uint32_t)c, v);
      uint32_t w[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float a = (c + 2 * i < len) ? exp2f(v[2 * i] * p.scale_log2 - mx) * inv : 0.f;
        float bb = (c + 2 * i + 1 < len) ? exp2f(v[2 * i + 1] * p.scale_log2 - mx) * inv : 0.f;
        if (len == 0) a = bb = __int_as_float(0x7fc00000);
        __nv_bfloat162 h = __floats2bfloat162_rn(a, bb);
        w[i] = *reinterpret_cast<uint32_t*>(&h);
      }
      // K-major SWIZZLE_128B: 64-key chunk (c >> 6), row r, 16-byte unit j of the 128-byte row stored at unit j ^ (r & 7)
      uint8_t* base = sQ + (size_t)(c >> 6) * QCHUNK + (size_t)r * 128;
      const int j = (c & 63) >> 3;
      *reinterpret_cast<uint4*>(base + (((j) ^ (r & 7)) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
      *reinterpret_cast<uint4*>(base + (((j + 1) ^ (r & 7)) << 4)) = make_uint4(w[4], w[5], w[6], w[7]);
      if (row_ok) {
        *reinterpret_cast<uint4*>(prow + c) = make_uint4(w[0], w[1], w[2], w[3]);
        *reinterpret_cast<uint4*>(prow + c + 8) = make_uint4(w[4], w[5], w[6], w[7]);
      }
    }
    tc_fence_before();                                            // order the tcgen05.ld's before the MMA that overwrites nothing of S, but reuses TMEM reads
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy smem writes -> visible to the tensor core's async proxy
    mbar_arrive(p_ready);
    // epilogue: O
    mbar_wait(o_done, 0);
    tc_fence_after();
    bf16* orow = p.O + ((long long)b * p.Tq + tq) * D;
    for (int c = 0; c < D; c += 16) {
      float v[16];
      tmem_ld16(trow + 256u + (uint32_t)c, v);
      if (row_ok) {
        uint32_t w[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) { __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]); w[i] = *reinterpret_cast<uint32_t*>(&h); }
        *reinterpret_cast<uint4*>(orow + c) = make_uint4(w[0], w[1], w[2], w[3]);
        *reinterpret_cast<uint4*>(orow + c + 8) = make_uint4(w[4], w[5], w[6], w[7]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// dS = scale * P o (dP - rowsum(dP o P)):  one warp per (batch, query) row; P bf16 [rows][SP], dP fp32 [rows][SP]
__global__ void att_softmax_bwd_kernel(const bf16* __restrict__ P, const float* __restrict__ dP, bf16* __restrict__ dS, long long rows,
                                       int S, int SP, float scale) {
  const long long row = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const bf16* p = P + row * SP; const float* g = dP + row * SP; bf16* o = dS + row * SP;
  float dot = 0.f;
  for (int j = lane; j < S; j += 32) dot += __bfloat162float(p[j]) * g[j];
  dot = warp_sum(dot);
  for (int j = lane; j < SP; j += 32) o[j] = __float2bfloat16_rn(j < S ? scale * __bfloat162float(p[j]) * (g[j] - dot) : 0.f);
}

int make_map3(CUtensorMap* m, const void* base, long long inner, long long rows, long long Z, long long ld, long long stride, int box_rows) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { vca_set_error("cuTensorMapEncodeTiled entry point unavailable"); return VCA_ERR_CUDA; }
  cuuint64_t gd[3] = {(cuuint64_t)inner, (cuuint64_t)rows, (cuuint64_t)Z};
  cuuint64_t gs[2] = {(cuuint64_t)ld * 2, (cuuint64_t)stride * 2};
  cuuint32_t bx[3] = {64, (cuuint32_t)box_rows, 1};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { vca_set_error("attention: cuTensorMapEncodeTiled failed (CUresult %d)", (int)r); return VCA_ERR_CUDA; }
  return VCA_OK;
}

}  // namespace

extern "C" {

// 1 when vca_att_fwd_tc handles (S keys, d channels)
int vca_att_tc_supported(int S, int d) { return d == D && S >= 1 && S <= 256; }

// Q [B][Tq][256], K, V [B][S][256] bf16 contiguous; lens int32 [B]; O [B][Tq][256] bf16; P [B][Tq][SP] bf16 with
// SP = (S + 15) / 16 * 16 (softmax probabilities, zero at masked / padded keys).  scale = 1/sqrt(out_dim).
int vca_att_fwd_tc(const void* Q, const void* K, const void* V, const int* lens, void* O, void* P, int B, int Tq, int S, float scale,
                   cudaStream_t s) {
  VCA_CHECK_ARG(Q && K && V && lens && O && P && B > 0 && Tq > 0 && vca_att_tc_supported(S, D));
  AttParams p;
  p.Tq = Tq; p.S = S; p.SP = (S + 15) / 16 * 16;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.lens = lens; p.O = (bf16*)O; p.P = (bf16*)P;
  CUtensorMap tmQ, tmK, tmV;
  int rc = make_map3(&tmQ, Q, D, Tq, B, D, (long long)Tq * D, 128); if (rc) return rc;
  rc = make_map3(&tmK, K, D, S, B, D, (long long)S * D, p.SP); if (rc) return rc;
  rc = make_map3(&tmV, V, D, S, B, D, (long long)S * D, p.SP); if (rc) return rc;
  const size_t smem = 4 * QCHUNK + 4 * 256 * 128 + 1024 + 256;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(att_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) {
      vca_set_error("cudaFuncSetAttribute(att_fwd_tc_kernel) failed"); return VCA_ERR_CUDA;
    }
    attr_set = true;
  }
  dim3 grid((unsigned)((Tq + 127) / 128), (unsigned)B);
  att_fwd_tc_kernel<<<grid, 192, smem, s>>>(tmQ, tmK, tmV, p);
  VCA_LAUNCH_CHECK();
  return VCA_OK;
}

// dS [rows][SP] bf16 = scale * P o (dP - rowsum(dP o P));  P bf16, dP fp32, both [rows][SP]; columns >= S are written as 0
int vca_att_softmax_bwd(const void* P, const float* dP, void* dS, long long rows, int S, int SP, float scale, cudaStream_t s) {
  VCA_CHECK_ARG(P && dP && dS && rows > 0 && S > 0 && SP >= S);
  const long long threads = rows * 32;
  att_softmax_bwd_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, s>>>((const bf16*)P, dP, (bf16*)dS, rows, S, SP, scale);
  VCA_LAUNCH_CHECK();
  return VCA_OK;
}

}  // extern "C"
