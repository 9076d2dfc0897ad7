// Degenerate-channel convolutions and the weight re-pack -- HBM/latency-bound kernels that the tiled GEMM core would run
// with 63/64 of its tile empty:
//   * Cin = 1 forward / wgrad: the 5x5 (1 -> 32) stems of the three discriminators and the 3x3/s2 (1 -> 128) stem of the
//     sync discriminator, all reading the 80x300 mel (generator.py:272, 323),
//   * Cout = 1 pointwise (to_mel heads, generator.py:208-225): forward, dgrad, wgrad,
//   * parameter layout [Cout][Cin][taps] fp32 -> the two packed K-major slabs, as a shared-memory tile transpose.
#include "vec.cuh"

namespace {

constexpr int SMALL_MAX_W = 8192;

// ---- Cin = 1 forward: one thread = one output pixel x 8 output channels; weights [tap][Cout] in shared memory.
template <class T>
__global__ void __launch_bounds__(256) cin1_fwd_kernel(ConvGeom g, const T* __restrict__ x, const T* __restrict__ wf,
                                                       const float* __restrict__ bias, T* __restrict__ y, long long M) {
  __shared__ __align__(16) float sw[SMALL_MAX_W];
  const int taps = g.KD * g.KH * g.KW;
  for (int i = threadIdx.x; i < taps * g.Cout; i += blockDim.x) sw[i] = to_f(wf[i]);
  __syncthreads();
  const int CG = g.Cout >> 3;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < M * CG; i += (long long)gridDim.x * blockDim.x) {
    const int cg = (int)(i % CG); long long mm = i / CG;
    const int ow = (int)(mm % g.OW); mm /= g.OW;
    const int oh = (int)(mm % g.OH); mm /= g.OH;
    const int od = (int)(mm % g.OD); const long long n = mm / g.OD;
    float acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = bias ? bias[cg * 8 + k] : 0.f;
    for (int kd = 0; kd < g.KD; ++kd) {
      const int id = od * g.sd - g.pd + kd;
      if ((unsigned)id >= (unsigned)g.ID) continue;
      for (int kh = 0; kh < g.KH; ++kh) {
        const int ih = oh * g.sh - g.ph + kh;
        if ((unsigned)ih >= (unsigned)g.IH) continue;
        const T* xr = x + ((n * g.ID + id) * g.IH + ih) * (long long)g.IW;
        const float* wr = sw + ((kd * g.KH + kh) * g.KW) * g.Cout + cg * 8;
        for (int kw = 0; kw < g.KW; ++kw) {
          const int iw = ow * g.sw - g.pw + kw;
          if ((unsigned)iw >= (unsigned)g.IW) continue;
          const float xv = to_f(xr[iw]);
          const float4 w0 = *reinterpret_cast<const float4*>(wr + kw * g.Cout);
          const float4 w1 = *reinterpret_cast<const float4*>(wr + kw * g.Cout + 4);
          acc[0] = fmaf(xv, w0.x, acc[0]); acc[1] = fmaf(xv, w0.y, acc[1]); acc[2] = fmaf(xv, w0.z, acc[2]); acc[3] = fmaf(xv, w0.w, acc[3]);
          acc[4] = fmaf(xv, w1.x, acc[4]); acc[5] = fmaf(xv, w1.y, acc[5]); acc[6] = fmaf(xv, w1.z, acc[6]); acc[7] = fmaf(xv, w1.w, acc[7]);
        }
      }
    }
    T* o = y + (i / CG) * g.Cout + cg * 8;
    if (sizeof(T) == 2) Vec<bf16>::store(reinterpret_cast<bf16*>(o), acc);
    else { Vec<float>::store(reinterpret_cast<float*>(o), acc); Vec<float>::store(reinterpret_cast<float*>(o) + 4, acc + 4); }
  }
}

// ---- Cin = 1 wgrad: dw[co][tap] = sum_p dY[p][co] * x[p (+) tap].  A CTA stages 64 output pixels (dY rows and the
// taps' input samples) in shared memory; thread j accumulates the (tap, co) pairs j, j+256, ... and adds them
// atomically at the end.
constexpr int WG_PIX = 64, WG_MAXPAIR = 8;
template <class T>
__global__ void __launch_bounds__(256) cin1_wgrad_kernel(ConvGeom g, const T* __restrict__ dy, const T* __restrict__ x,
                                                         float* __restrict__ dw, long long M) {
  extern __shared__ float sm[];
  const int taps = g.KD * g.KH * g.KW, C = g.Cout, npair = taps * C;
  float* sdy = sm;                      // [WG_PIX][C]
  float* sx = sm + WG_PIX * C;          // [taps][WG_PIX]
  float acc[WG_MAXPAIR];
#pragma unroll
  for (int k = 0; k < WG_MAXPAIR; ++k) acc[k] = 0.f;
  const long long ntiles = (M + WG_PIX - 1) / WG_PIX;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long m0 = tile * WG_PIX;
    __syncthreads();
    for (int i = threadIdx.x; i < WG_PIX * C; i += blockDim.x) {
      const long long m = m0 + i / C;
      sdy[i] = m < M ? to_f(dy[m * C + (i % C)]) : 0.f;
    }
    for (int i = threadIdx.x; i < taps * WG_PIX; i += blockDim.x) {
      const int t = i / WG_PIX, p = i - t * WG_PIX;
      long long mm = m0 + p;
      float v = 0.f;
      if (mm < M) {
        const int ow = (int)(mm % g.OW); mm /= g.OW;
        const int oh = (int)(mm % g.OH); mm /= g.OH;
        const int od = (int)(mm % g.OD); const long long n = mm / g.OD;
        int tt = t;
        const int kw = tt % g.KW; tt /= g.KW; const int kh = tt % g.KH; const int kd = tt / g.KH;
        const int id = od * g.sd - g.pd + kd, ih = oh * g.sh - g.ph + kh, iw = ow * g.sw - g.pw + kw;
        if ((unsigned)id < (unsigned)g.ID && (unsigned)ih < (unsigned)g.IH && (unsigned)iw < (unsigned)g.IW)
          v = to_f(x[((n * g.ID + id) * g.IH + ih) * (long long)g.IW + iw]);
      }
      sx[i] = v;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < WG_MAXPAIR; ++k) {
      const int pair = threadIdx.x + k * 256;
      if (pair < npair) {
        const int t = pair / C, co = pair - t * C;
        const float* xs = sx + t * WG_PIX;
        float a = 0.f;
#pragma unroll 8
        for (int p = 0; p < WG_PIX; ++p) a = fmaf(sdy[p * C + co], xs[p], a);
        acc[k] += a;
      }
    }
  }
#pragma unroll
  for (int k = 0; k < WG_MAXPAIR; ++k) {
    const int pair = threadIdx.x + k * 256;
    if (pair < npair) { const int t = pair / C, co = pair - t * C; atomicAdd(&dw[(long long)co * taps + t], acc[k]); }
  }
}

// ---- Cout = 1 pointwise convolution (1x1, stride 1): y[p] = sum_c x[p][c] w[c] + b
template <class T, class VT>
__global__ void pw1_fwd_kernel(const T* __restrict__ x, const T* __restrict__ w, const float* __restrict__ bias, T* __restrict__ y,
                               long long M, int C) {
  extern __shared__ float sw1[];
  for (int i = threadIdx.x; i < C; i += blockDim.x) sw1[i] = to_f(w[i]);
  __syncthreads();
  constexpr int V = VT::N;
  const float b = bias ? bias[0] : 0.f;
  for (long long m = blockIdx.x * (long long)blockDim.x + threadIdx.x; m < M; m += (long long)gridDim.x * blockDim.x) {
    float acc = b;
    for (int c = 0; c < C; c += V) {
      float v[V];
      VT::load(x + m * C + c, v);
#pragma unroll
      for (int k = 0; k < V; ++k) acc = fmaf(v[k], sw1[c + k], acc);
    }
    y[m] = from_f<T>(acc);
  }
}
// dx[p][c] = dy[p] * w[c]
template <class T, class VT>
__global__ void pw1_dgrad_kernel(const T* __restrict__ dy, const T* __restrict__ w, T* __restrict__ dx, long long M, int C) {
  constexpr int V = VT::N;
  const int CV = C / V;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < M * CV; i += (long long)gridDim.x * blockDim.x) {
    const int cv = (int)(i % CV); const long long m = i / CV;
    const float g = to_f(dy[m]);
    float v[V];
#pragma unroll
    for (int k = 0; k < V; ++k) v[k] = g * to_f(w[cv * V + k]);
    VT::store(dx + i * V, v);
  }
}
// dw[c] = sum_p dy[p] * x[p][c]   (fp32 atomics; dw zeroed by the caller)
template <class T, class VT>
__global__ void __launch_bounds__(256) pw1_wgrad_kernel(const T* __restrict__ dy, const T* __restrict__ x, float* __restrict__ dw,
                                                        long long M, int C) {
  constexpr int V = VT::N;
  __shared__ float sh[256 * V];
  const int cv = blockIdx.y * blockDim.x + threadIdx.x;
  const bool active = cv < C / V;
  float a[V];
#pragma unroll
  for (int k = 0; k < V; ++k) a[k] = 0.f;
  if (active) {
    for (long long r = (long long)blockIdx.x * blockDim.y + threadIdx.y; r < M; r += (long long)gridDim.x * blockDim.y) {
      const float g = to_f(dy[r]);
      float v[V];
      VT::load(x + r * C + cv * V, v);
#pragma unroll
      for (int k = 0; k < V; ++k) a[k] = fmaf(g, v[k], a[k]);
    }
  }
  const int tx = threadIdx.x, ty = threadIdx.y, TX = blockDim.x, TY = blockDim.y;
#pragma unroll
  for (int k = 0; k < V; ++k) sh[(ty * TX + tx) * V + k] = a[k];
  __syncthreads();
  if (ty == 0 && active) {
#pragma unroll
    for (int k = 0; k < V; ++k) {
      float s = 0.f;
      for (int yy = 0; yy < TY; ++yy) s += sh[(yy * TX + tx) * V + k];
      atomicAdd(&dw[cv * V + k], s);
    }
  }
}

// ---- weight re-pack as a tile transpose: CTA = 32 co x 32 ci, taps in chunks of 8
constexpr int PK = 32, PK_T = 8;
template <class T>
__global__ void __launch_bounds__(256) pack_tiled_kernel(const float* __restrict__ w, T* __restrict__ wf, T* __restrict__ wd,
                                                         int Cout, int Cin, int taps) {
  __shared__ float s[PK_T][PK][PK + 1];
  const int co0 = blockIdx.y * PK, ci0 = blockIdx.x * PK;
  {
    const int t0 = blockIdx.z * PK_T;            // one tap chunk per CTA (grid.z) for parallelism on small layers
    const int tc = min(PK_T, taps - t0);
    for (int i = threadIdx.x; i < PK * PK * tc; i += blockDim.x) {
      const int t = i % tc; int r = i / tc; const int ci = r % PK; const int co = r / PK;
      float v = 0.f;
      if (co0 + co < Cout && ci0 + ci < Cin) v = w[((long long)(co0 + co) * Cin + ci0 + ci) * taps + t0 + t];
      s[t][co][ci] = v;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < PK * PK * tc; i += blockDim.x) {
      const int a = i % PK; int r = i / PK; const int b = r % PK; const int t = r / PK;
      // wd[t][co=b][ci=a]: ci fastest;   wf[t][ci=b][co=a]: co fastest
      if (wd && co0 + b < Cout && ci0 + a < Cin) wd[((long long)(t0 + t) * Cout + co0 + b) * Cin + ci0 + a] = from_f<T>(s[t][b][a]);
      if (wf && ci0 + b < Cin && co0 + a < Cout) wf[((long long)(t0 + t) * Cin + ci0 + b) * Cout + co0 + a] = from_f<T>(s[t][a][b]);
    }
  }
}

// ---- batched re-pack: every conv weight of an optimizer group in ONE launch -------------------------------------
// Job table (device, 8 x int64 per job): w (fp32 [Cout][Cin][taps]), wf, wd (either may be 0), Cout, Cin, taps,
// first CTA of the job, unused.  A CTA finds its job by binary search on the CTA offsets and then does what a CTA of
// pack_tiled_kernel does.  Replaces ~75 launches per optimizer step (the step re-packs every weight Adam just changed).
template <class T>
__global__ void __launch_bounds__(256) pack_batched_kernel(const long long* __restrict__ jobs, int njobs) {
  __shared__ float s[PK_T][PK][PK + 1];
  int lo = 0, hi = njobs - 1;
  const long long me = blockIdx.x;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (jobs[(long long)mid * 8 + 6] <= me) lo = mid; else hi = mid - 1;
  }
  const long long* j = jobs + (long long)lo * 8;
  const float* __restrict__ w = reinterpret_cast<const float*>(j[0]);
  T* __restrict__ wf = reinterpret_cast<T*>(j[1]);
  T* __restrict__ wd = reinterpret_cast<T*>(j[2]);
  const int Cout = (int)j[3], Cin = (int)j[4], taps = (int)j[5];
  int local = (int)(me - j[6]);
  const int tiles_ci = (Cin + PK - 1) / PK, tiles_co = (Cout + PK - 1) / PK;
  const int bx = local % tiles_ci; local /= tiles_ci;
  const int by = local % tiles_co; const int bz = local / tiles_co;
  const int co0 = by * PK, ci0 = bx * PK, t0 = bz * PK_T;
  const int tc = min(PK_T, taps - t0);
  for (int i = threadIdx.x; i < PK * PK * tc; i += blockDim.x) {
    const int t = i % tc; int r = i / tc; const int ci = r % PK; const int co = r / PK;
    float v = 0.f;
    if (co0 + co < Cout && ci0 + ci < Cin) v = w[((long long)(co0 + co) * Cin + ci0 + ci) * taps + t0 + t];
    s[t][co][ci] = v;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < PK * PK * tc; i += blockDim.x) {
    const int a = i % PK; int r = i / PK; const int b = r % PK; const int t = r / PK;
    if (wd && co0 + b < Cout && ci0 + a < Cin) wd[((long long)(t0 + t) * Cout + co0 + b) * Cin + ci0 + a] = from_f<T>(s[t][b][a]);
    if (wf && ci0 + b < Cin && co0 + a < Cout) wf[((long long)(t0 + t) * Cin + ci0 + b) * Cout + co0 + a] = from_f<T>(s[t][a][b]);
  }
}

// ---- batched un-slab: tap-major gradient slabs folded back into the parameter layout, ONE launch per optimizer group ----
// The tcgen05 wgrad kernels add a filter's gradient to a TAP-MAJOR fp32 slab [taps][Cout][Cin] (vca_conv_wgrad_tc_tm: TMA
// reduce-add boxes instead of scattered 4-byte atomics).  Job table as above with j[0] = the parameter's .grad
// ([Cout][Cin][taps], ADDED to), j[1] = its slab (read, then left ZEROED for the next step), j[2] unused.
// A CTA owns UCO output channels x UCI input channels x ALL taps of one filter: the slab side is read in 128-byte runs (one
// per tap and output channel), the gradient side is ONE contiguous run of UCI * taps floats per output channel (the
// [Cout][Cin][taps] layout keeps (ci, tap) adjacent) -- the first version wrote 8-tap (32-byte) fragments and took three
// passes over every line: 2 TB/s.  All loads of a phase are issued before the first dependent store.
constexpr int UCO = 8, UCI = 32;
__global__ void __launch_bounds__(256) unslab_batched_kernel(const long long* __restrict__ jobs, int njobs) {
  extern __shared__ float us[];                  // [UCO][UCI * taps]
  int lo = 0, hi = njobs - 1;
  const long long me = blockIdx.x;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (jobs[(long long)mid * 8 + 6] <= me) lo = mid; else hi = mid - 1;
  }
  const long long* j = jobs + (long long)lo * 8;
  float* __restrict__ grad = reinterpret_cast<float*>(j[0]);
  float* __restrict__ slab = reinterpret_cast<float*>(j[1]);
  const int Cout = (int)j[3], Cin = (int)j[4], taps = (int)j[5];
  const int local = (int)(me - j[6]);
  const int tiles_ci = (Cin + UCI - 1) / UCI;
  const int ci0 = (local % tiles_ci) * UCI, co0 = (local / tiles_ci) * UCO;
  const int a = threadIdx.x % UCI, bq = threadIdx.x / UCI;          // this thread's (ci, co) inside the tile
  const bool ok = co0 + bq < Cout && ci0 + a < Cin;
  const int row = UCI * taps;
  float* q = slab + ((long long)(co0 + bq)) * Cin + ci0 + a;         // + t * Cout * Cin per tap
  const long long tstride = (long long)Cout * Cin;
  for (int t0 = 0; t0 < taps; t0 += 8) {
    float v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = (ok && t0 + k < taps) ? q[(t0 + k) * tstride] : 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k)
      if (t0 + k < taps) {
        us[bq * row + a * taps + t0 + k] = v[k];
        if (ok) q[(t0 + k) * tstride] = 0.f;
      }
  }
  __syncthreads();
  const int nci = min(UCI, Cin - ci0);
  const int n = nci * taps;                                            // valid floats per output-channel row
  for (int b = 0; b < UCO && co0 + b < Cout; ++b) {
    float* g = grad + ((long long)(co0 + b) * Cin + ci0) * taps;
    for (int i0 = 0; i0 < n; i0 += 256 * 4) {
      float v[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) { const int i = i0 + k * 256 + (int)threadIdx.x; v[k] = i < n ? g[i] : 0.f; }
#pragma unroll
      for (int k = 0; k < 4; ++k) { const int i = i0 + k * 256 + (int)threadIdx.x; if (i < n) g[i] = v[k] + us[b * row + i]; }
    }
  }
}

}  // namespace

extern "C" {
// jobs: device table (8 x int64 per job: grad, slab, 0, Cout, Cin, taps, first CTA, 0); total_ctas = sum of
// vca_unslab_job_ctas over the jobs; max_taps = the largest taps of any job (sizes the shared-memory tile).  grad += slab (transposed), slab = 0.  Must be ordered after every kernel that adds to
// either buffer (it is a plain read-modify-write).
int vca_grad_unslab_batched(const long long* jobs, int njobs, long long total_ctas, int max_taps, cudaStream_t s) {
  VCA_CHECK_ARG(jobs && njobs > 0 && total_ctas > 0 && total_ctas < 0x7fffffffLL && max_taps > 0);
  const size_t smem = (size_t)UCO * UCI * max_taps * sizeof(float);
  if (smem > 200 * 1024) { vca_set_error("vca_grad_unslab_batched: %d taps do not fit in shared memory", max_taps); return VCA_ERR_UNSUPPORTED; }
  if (smem > 48 * 1024) {
    static size_t attr = 0;
    if (smem > attr) {
      if (cudaFuncSetAttribute(unslab_batched_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
        vca_set_error("cudaFuncSetAttribute(unslab_batched_kernel) failed"); return VCA_ERR_CUDA;
      }
      attr = smem;
    }
  }
  unslab_batched_kernel<<<(unsigned)total_ctas, 256, smem, s>>>(jobs, njobs);
  VCA_LAUNCH_CHECK();
  return VCA_OK;
}
// CTAs a job of this shape occupies in vca_grad_unslab_batched
int vca_unslab_job_ctas(int Cout, int Cin, int taps) {
  if (Cout <= 0 || Cin <= 0 || taps <= 0) return 0;
  return ((Cin + UCI - 1) / UCI) * ((Cout + UCO - 1) / UCO);
}
// CTAs a job of this shape occupies in vca_pack_conv_weights_batched (the host builds the CTA offsets of the table with it)
int vca_pack_job_ctas(int Cout, int Cin, int taps) {
  if (Cout <= 0 || Cin <= 0 || taps <= 0) return 0;
  return ((Cin + PK - 1) / PK) * ((Cout + PK - 1) / PK) * ((taps + PK_T - 1) / PK_T);
}
// jobs: device table described above; total_ctas = sum of vca_pack_job_ctas over the jobs.
int vca_pack_conv_weights_batched(int dtype, const long long* jobs, int njobs, long long total_ctas, cudaStream_t s) {
  VCA_CHECK_ARG(jobs && njobs > 0 && total_ctas > 0 && total_ctas < 0x7fffffffLL);
  if (dtype == VCA_F32) pack_batched_kernel<float><<<(unsigned)total_ctas, 256, 0, s>>>(jobs, njobs);
  else pack_batched_kernel<bf16><<<(unsigned)total_ctas, 256, 0, s>>>(jobs, njobs);
  VCA_LAUNCH_CHECK();
  return VCA_OK;
}
}

#define SMALL_T(dtype, CALL_F32, CALL_BF16) do { if ((dtype) == VCA_F32) { CALL_F32; } else { CALL_BF16; } } while (0)

// 1 = handled, 0 = not applicable, <0 error
// conv_c32.cu: lane = output channel kernels for Cin = 1 -> Cout = 32, stride 1, 5x5 / 3x3
int conv_c32_fwd(int dtype, const ConvGeom& g, const void* x, const void* wf, const float* bias, void* y, cudaStream_t s);
int conv_c32_wgrad(int dtype, const ConvGeom& g, const void* dy, const void* x, float* dw, cudaStream_t s);

int conv_cin1_fwd(int dtype, const ConvGeom& g, const void* x, const void* wf, const float* bias, void* y, cudaStream_t s) {
  if (const int r = conv_c32_fwd(dtype, g, x, wf, bias, y, s)) return r;
  const int taps = g.KD * g.KH * g.KW;
  if (g.Cin != 1 || g.Cout % 8 || taps * g.Cout > SMALL_MAX_W || !vca_aligned16(y)) return 0;
  const long long M = (long long)g.N * g.OD * g.OH * g.OW;
  unsigned grid = vca_grid_1d(M * (g.Cout / 8), 256);
  SMALL_T(dtype, (cin1_fwd_kernel<float><<<grid, 256, 0, s>>>(g, (const float*)x, (const float*)wf, bias, (float*)y, M)),
          (cin1_fwd_kernel<bf16><<<grid, 256, 0, s>>>(g, (const bf16*)x, (const bf16*)wf, bias, (bf16*)y, M)));
  if (cudaGetLastError() != cudaSuccess) { vca_set_error("cin1_fwd_kernel launch failed"); return VCA_ERR_CUDA; }
  return 1;
}
int conv_cin1_wgrad(int dtype, const ConvGeom& g, const void* dy, const void* x, float* dw, cudaStream_t s) {
  if (const int r = conv_c32_wgrad(dtype, g, dy, x, dw, s)) return r;
  const int taps = g.KD * g.KH * g.KW;
  if (g.Cin != 1 || taps * g.Cout > 256 * WG_MAXPAIR) return 0;
  const long long M = (long long)g.N * g.OD * g.OH * g.OW;
  const size_t smem = (size_t)(WG_PIX * g.Cout + taps * WG_PIX) * sizeof(float);
  if (smem > 48 * 1024) return 0;
  long long ntiles = (M + WG_PIX - 1) / WG_PIX;
  unsigned grid = (unsigned)(ntiles < 4LL * vca_num_sms() ? ntiles : 4LL * vca_num_sms());
  SMALL_T(dtype, (cin1_wgrad_kernel<float><<<grid, 256, smem, s>>>(g, (const float*)dy, (const float*)x, dw, M)),
          (cin1_wgrad_kernel<bf16><<<grid, 256, smem, s>>>(g, (const bf16*)dy, (const bf16*)x, dw, M)));
  if (cudaGetLastError() != cudaSuccess) { vca_set_error("cin1_wgrad_kernel launch failed"); return VCA_ERR_CUDA; }
  return 1;
}
static bool pw1_geom(const ConvGeom& g) {
  return g.Cout == 1 && g.KD * g.KH * g.KW == 1 && g.sd == 1 && g.sh == 1 && g.sw == 1 && g.pd == 0 && g.ph == 0 && g.pw == 0 &&
         g.Cin <= 4096;
}
template <class T>
static bool pw1_vec(const ConvGeom& g, const void* p) { return g.Cin % Vec<T>::N == 0 && vca_aligned16(p); }

int conv_pw1_fwd(int dtype, const ConvGeom& g, const void* x, const void* w, const float* bias, void* y, cudaStream_t s) {
  if (!pw1_geom(g)) return 0;
  const long long M = (long long)g.N * g.OD * g.OH * g.OW;
  unsigned grid = vca_grid_1d(M, 256);
  const size_t sm = g.Cin * sizeof(float);
  if (dtype == VCA_F32) {
    if (pw1_vec<float>(g, x)) pw1_fwd_kernel<float, Vec<float>><<<grid, 256, sm, s>>>((const float*)x, (const float*)w, bias, (float*)y, M, g.Cin);
    else pw1_fwd_kernel<float, Vec1<float>><<<grid, 256, sm, s>>>((const float*)x, (const float*)w, bias, (float*)y, M, g.Cin);
  } else {
    if (pw1_vec<bf16>(g, x)) pw1_fwd_kernel<bf16, Vec<bf16>><<<grid, 256, sm, s>>>((const bf16*)x, (const bf16*)w, bias, (bf16*)y, M, g.Cin);
    else pw1_fwd_kernel<bf16, Vec1<bf16>><<<grid, 256, sm, s>>>((const bf16*)x, (const bf16*)w, bias, (bf16*)y, M, g.Cin);
  }
  if (cudaGetLastError() != cudaSuccess) { vca_set_error("pw1_fwd_kernel launch failed"); return VCA_ERR_CUDA; }
  return 1;
}
int conv_pw1_dgrad(int dtype, const ConvGeom& g, const void* dy, const void* w, void* dx, cudaStream_t s) {
  if (!pw1_geom(g)) return 0;
  const long long M = (long long)g.N * g.ID * g.IH * g.IW;
  if (dtype == VCA_F32) {
    if (pw1_vec<float>(g, dx)) pw1_dgrad_kernel<float, Vec<float>><<<vca_grid_1d(M * (g.Cin / 4), 256), 256, 0, s>>>((const float*)dy, (const float*)w, (float*)dx, M, g.Cin);
    else pw1_dgrad_kernel<float, Vec1<float>><<<vca_grid_1d(M * g.Cin, 256), 256, 0, s>>>((const float*)dy, (const float*)w, (float*)dx, M, g.Cin);
  } else {
    if (pw1_vec<bf16>(g, dx)) pw1_dgrad_kernel<bf16, Vec<bf16>><<<vca_grid_1d(M * (g.Cin / 8), 256), 256, 0, s>>>((const bf16*)dy, (const bf16*)w, (bf16*)dx, M, g.Cin);
    else pw1_dgrad_kernel<bf16, Vec1<bf16>><<<vca_grid_1d(M * g.Cin, 256), 256, 0, s>>>((const bf16*)dy, (const bf16*)w, (bf16*)dx, M, g.Cin);
  }
  if (cudaGetLastError() != cudaSuccess) { vca_set_error("pw1_dgrad_kernel launch failed"); return VCA_ERR_CUDA; }
  return 1;
}
int conv_pw1_wgrad(int dtype, const ConvGeom& g, const void* dy, const void* x, float* dw, cudaStream_t s) {
  if (!pw1_geom(g)) return 0;
  const long long M = (long long)g.N * g.OD * g.OH * g.OW;
  if (dtype == VCA_F32) {
    if (!pw1_vec<float>(g, x)) return 0;
    RowColGrid rc = row_col_grid(M, g.Cin / 4, 148 * 4);
    pw1_wgrad_kernel<float, Vec<float>><<<rc.grid, rc.block, 0, s>>>((const float*)dy, (const float*)x, dw, M, g.Cin);
  } else {
    if (!pw1_vec<bf16>(g, x)) return 0;
    RowColGrid rc = row_col_grid(M, g.Cin / 8, 148 * 4);
    pw1_wgrad_kernel<bf16, Vec<bf16>><<<rc.grid, rc.block, 0, s>>>((const bf16*)dy, (const bf16*)x, dw, M, g.Cin);
  }
  if (cudaGetLastError() != cudaSuccess) { vca_set_error("pw1_wgrad_kernel launch failed"); return VCA_ERR_CUDA; }
  return 1;
}
int pack_weight_tiled(int dtype, const float* w, void* wf, void* wd, int Cout, int Cin, int taps, cudaStream_t s) {
  dim3 grid((Cin + PK - 1) / PK, (Cout + PK - 1) / PK, (taps + PK_T - 1) / PK_T);
  if (grid.y > 65535u || grid.z > 65535u) return 0;
  SMALL_T(dtype, (pack_tiled_kernel<float><<<grid, 256, 0, s>>>(w, (float*)wf, (float*)wd, Cout, Cin, taps)),
          (pack_tiled_kernel<bf16><<<grid, 256, 0, s>>>(w, (bf16*)wf, (bf16*)wd, Cout, Cin, taps)));
  if (cudaGetLastError() != cudaSuccess) { vca_set_error("pack_tiled_kernel launch failed"); return VCA_ERR_CUDA; }
  return 1;
}

// ---- pixel-pair merge of a small-channel convolution -----------------------------------------------------------
// A (Cin -> Cout, KH x 5, pad 2, stride 1) convolution over [N,H,W,Cin] is the SAME linear map as a
// (2Cin -> 2Cout, KH x 3, pad 1) convolution over the free view [N,H,W/2,2Cin] (two neighbouring pixels = one
// 2Cin-channel pixel) with the Toeplitz-expanded weight
//     w2[p*Cout+co][s*Cin+ci][kh][jj] = w[co][ci][kh][2jj+s-p]   (0 outside 0..4),
// p / s = position of the output / input pixel inside its pair.  For Cin = Cout = 32 that doubles the MMA N (the
// tensor pipe of the SS-mode kernels is bound by the A-operand read, so N = 32 caps it at 25 %) for 1.2x the MACs.
namespace {
__global__ void pair_expand_kernel(const float* __restrict__ w, float* __restrict__ w2, int Cout, int Cin, int KH) {
  const long long total = 4LL * Cout * Cin * KH * 3;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long r = i;
    const int jj = (int)(r % 3); r /= 3;
    const int kh = (int)(r % KH); r /= KH;
    const int ci2 = (int)(r % (2 * Cin)); const int co2 = (int)(r / (2 * Cin));
    const int p = co2 / Cout, co = co2 - p * Cout, s = ci2 / Cin, ci = ci2 - s * Cin;
    const int kw = 2 * jj + s - p;
    w2[i] = (kw >= 0 && kw < 5) ? w[(((long long)co * Cin + ci) * KH + kh) * 5 + kw] : 0.f;
  }
}
// adjoint: dw[co][ci][kh][kw] (+)= sum over (p, s, jj) with 2jj+s-p == kw of dw2[p*Cout+co][s*Cin+ci][kh][jj]
__global__ void pair_contract_kernel(const float* __restrict__ dw2, float* __restrict__ dw, int Cout, int Cin, int KH,
                                     int accumulate) {
  const long long total = (long long)Cout * Cin * KH * 5;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long r = i;
    const int kw = (int)(r % 5); r /= 5;
    const int kh = (int)(r % KH); r /= KH;
    const int ci = (int)(r % Cin); const int co = (int)(r / Cin);
    float a = 0.f;
#pragma unroll
    for (int p = 0; p < 2; ++p)
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        const int t = kw + p - s;
        if (t >= 0 && (t & 1) == 0 && t / 2 < 3)
          a += dw2[((((long long)(p * Cout + co)) * (2 * Cin) + s * Cin + ci) * KH + kh) * 3 + t / 2];
      }
    if (accumulate) atomicAdd(&dw[i], a);   // .grad sinks may be hit from several streams at once (real / fake pass): never a plain RMW
    else dw[i] = a;
  }
}
}  // namespace

extern "C" {
// w [Cout][Cin][KH][5] fp32 -> w2 [2Cout][2Cin][KH][3] fp32
int vca_pair_expand_weight(const float* w, float* w2, int Cout, int Cin, int KH, cudaStream_t s) {
  VCA_CHECK_ARG(w && w2 && Cout > 0 && Cin > 0 && KH > 0);
  pair_expand_kernel<<<vca_grid_1d(4LL * Cout * Cin * KH * 3, 256), 256, 0, s>>>(w, w2, Cout, Cin, KH);
  VCA_LAUNCH_CHECK();
  return VCA_OK;
}
// dw2 [2Cout][2Cin][KH][3] -> dw [Cout][Cin][KH][5] (accumulate != 0: added to the existing contents)
int vca_pair_contract_wgrad(const float* dw2, float* dw, int Cout, int Cin, int KH, int accumulate, cudaStream_t s) {
  VCA_CHECK_ARG(dw2 && dw && Cout > 0 && Cin > 0 && KH > 0);
  pair_contract_kernel<<<vca_grid_1d((long long)Cout * Cin * KH * 5, 256), 256, 0, s>>>(dw2, dw, Cout, Cin, KH, accumulate);
  VCA_LAUNCH_CHECK();
  return VCA_OK;
}
}
