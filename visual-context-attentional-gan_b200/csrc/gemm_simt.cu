// Generic fp32-accumulate SIMT implicit-GEMM core.  This is the *exact* path: it serves the fp32 parity
// mode (<=1e-4 vs the CPU oracle) for every contraction of the hot path, and in bf16 mode the layers the
// tcgen05 kernels (conv_tc.cu) do not cover (Cin=1 stems, stride-2 convs, tiny GEMMs).
//
// One tiled kernel, three problem families plugged in as functors:
//   conv forward   M = N*OD*OH*OW  N = Cout      K = taps*Cin   (src/models: every nn.Conv{1,2,3}d)
//   conv dgrad     M = N*ID*IH*IW  N = Cin       K = taps*Cout
//   conv wgrad     M = Cout        N = taps*Cin  K = pixels (split-K, fp32 atomics)
//   strided batched GEMM (nn.Linear, torch.bmm sites: generator.py:147-171, 336-357; GRU gates)
#include "common.cuh"

// conv_small.cu: degenerate-channel kernels (1 = handled, 0 = not applicable, < 0 error)
int conv_cin1_fwd(int dtype, const ConvGeom& g, const void* x, const void* wf, const float* bias, void* y, cudaStream_t s);
int conv_cin1_wgrad(int dtype, const ConvGeom& g, const void* dy, const void* x, float* dw, cudaStream_t s);
int conv_pw1_fwd(int dtype, const ConvGeom& g, const void* x, const void* w, const float* bias, void* y, cudaStream_t s);
int conv_pw1_dgrad(int dtype, const ConvGeom& g, const void* dy, const void* w, void* dx, cudaStream_t s);
int conv_pw1_wgrad(int dtype, const ConvGeom& g, const void* dy, const void* x, float* dw, cudaStream_t s);
int pack_weight_tiled(int dtype, const float* w, void* wf, void* wd, int Cout, int Cin, int taps, cudaStream_t s);
// conv_c32.cu: Cin = 1 -> Cout = 32 dgrad with lane = output channel
int conv_c32_dgrad(int dtype, const ConvGeom& g, const void* dy, const void* wd, void* dx, cudaStream_t s);

namespace {

constexpr int BM = 64, BN = 64, BK = 32, NT = 256;

template <class P>
__global__ void __launch_bounds__(NT) gemm_core(P p) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  __shared__ typename P::MInfo mi[BM];
  __shared__ typename P::NInfo ni[BN];
  __shared__ typename P::KInfo ki[BK];

  const int tid = threadIdx.x;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int z = blockIdx.z;
  int kbeg, kend;
  p.krange(z, kbeg, kend);

  if (tid < BM) mi[tid] = p.minfo(m0 + tid, z);
  else if (tid < BM + BN) ni[tid - BM] = p.ninfo(n0 + tid - BM, z);
  __syncthreads();

  float acc[4][4], comp[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = comp[i][j] = 0.f;

  const int tx = tid & 15, ty = tid >> 4;

  for (int k0 = kbeg; k0 < kend; k0 += BK) {
    __syncthreads();
    if (tid < BK) ki[tid] = p.kinfo(k0 + tid, kend, z);
    __syncthreads();
    // ---- A tile (BM x BK)
    if (P::A_K_FAST) {
      const int k = tid & 31;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int m = (tid >> 5) + 8 * i;
        As[k][m] = p.loadA(mi[m], ki[k]);
      }
    } else {
      const int m = tid & 63;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int k = (tid >> 6) + 4 * i;
        As[k][m] = p.loadA(mi[m], ki[k]);
      }
    }
    // ---- B tile (BK x BN)
    if (P::B_N_FAST) {
      const int n = tid & 63;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int k = (tid >> 6) + 4 * i;
        Bs[k][n] = p.loadB(ki[k], ni[n]);
      }
    } else {
      const int k = tid & 31;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int n = (tid >> 5) + 8 * i;
        Bs[k][n] = p.loadB(ki[k], ni[n]);
      }
    }
    __syncthreads();
    // two-level accumulation: a fresh partial sum per K tile, folded into the running sum afterwards.  Keeps the
    // fp32 rounding error of long reductions (K up to 16 000 taps*channels, or 1e5 pixels in wgrad) at the level of
    // a blocked CPU summation instead of growing like sqrt(K).
    float part[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) part[i][j] = 0.f;
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w};
      const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) part[i][j] = fmaf(av[i], bv[j], part[i][j]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {   // Kahan-compensated fold of the tile partial
        const float y = part[i][j] - comp[i][j];
        const float t = acc[i][j] + y;
        comp[i][j] = (t - acc[i][j]) - y;
        acc[i][j] = t;
      }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) p.store(mi[ty * 4 + i], ni[tx * 4 + j], acc[i][j]);
}

// ---------------------------------------------------------------------------------------------
// convolution forward
// ---------------------------------------------------------------------------------------------
template <class T>
struct ConvFwdP {
  static constexpr bool A_K_FAST = true, B_N_FAST = true;
  ConvGeom g;
  const T* x;      // [N,ID,IH,IW,Cin]
  const T* w;      // packed [taps][Cin][Cout]
  const float* bias;  // [Cout] or null
  T* y;            // [N,OD,OH,OW,Cout]
  int M, K;
  struct MInfo { long long xbase; long long yoff; int id0, ih0, iw0; int valid; };
  struct NInfo { int co; int valid; };
  struct KInfo { int kd, kh, kw, ci; int woff; int valid; };
  __device__ void krange(int, int& b, int& e) const { b = 0; e = K; }
  __device__ MInfo minfo(int m, int) const {
    MInfo r; r.valid = m < M;
    int mm = r.valid ? m : 0;
    int ow = mm % g.OW; mm /= g.OW;
    int oh = mm % g.OH; mm /= g.OH;
    int od = mm % g.OD; int n = mm / g.OD;
    r.xbase = (long long)n * g.ID * g.IH * g.IW;
    r.id0 = od * g.sd - g.pd; r.ih0 = oh * g.sh - g.ph; r.iw0 = ow * g.sw - g.pw;
    r.yoff = (long long)(r.valid ? m : 0) * g.Cout;
    return r;
  }
  __device__ NInfo ninfo(int n, int) const { NInfo r; r.valid = n < g.Cout; r.co = r.valid ? n : 0; return r; }
  __device__ KInfo kinfo(int k, int kend, int) const {
    KInfo r; r.valid = k < kend;
    int kk = r.valid ? k : 0;
    r.ci = kk % g.Cin; int tap = kk / g.Cin;
    r.kw = tap % g.KW; tap /= g.KW; r.kh = tap % g.KH; r.kd = tap / g.KH;
    r.woff = kk * g.Cout;
    return r;
  }
  __device__ float loadA(const MInfo& m, const KInfo& k) const {
    int id = m.id0 + k.kd, ih = m.ih0 + k.kh, iw = m.iw0 + k.kw;
    if (!(m.valid && k.valid) || (unsigned)id >= (unsigned)g.ID || (unsigned)ih >= (unsigned)g.IH ||
        (unsigned)iw >= (unsigned)g.IW) return 0.f;
    return to_f(x[(m.xbase + ((long long)id * g.IH + ih) * g.IW + iw) * g.Cin + k.ci]);
  }
  __device__ float loadB(const KInfo& k, const NInfo& n) const {
    return (k.valid && n.valid) ? to_f(w[k.woff + n.co]) : 0.f;
  }
  __device__ void store(const MInfo& m, const NInfo& n, float v) const {
    if (m.valid && n.valid) y[m.yoff + n.co] = from_f<T>(v + (bias ? bias[n.co] : 0.f));
  }
};

// ---------------------------------------------------------------------------------------------
// convolution dgrad: dX[n,id,ih,iw,ci] = sum_{tap,co} dY[n,(id+pd-kd)/sd,...,co] * W[co,ci,tap]
// ---------------------------------------------------------------------------------------------
template <class T>
struct ConvDgradP {
  static constexpr bool A_K_FAST = true, B_N_FAST = true;
  ConvGeom g;
  const T* dy;   // [N,OD,OH,OW,Cout]
  const T* w;    // packed [taps][Cout][Cin]
  T* dx;         // [N,ID,IH,IW,Cin]
  int M, K;
  struct MInfo { long long ybase; long long xoff; int id, ih, iw; int valid; };
  struct NInfo { int ci; int valid; };
  struct KInfo { int kd, kh, kw, co; int woff; int valid; };
  __device__ void krange(int, int& b, int& e) const { b = 0; e = K; }
  __device__ MInfo minfo(int m, int) const {
    MInfo r; r.valid = m < M;
    int mm = r.valid ? m : 0;
    r.iw = mm % g.IW; mm /= g.IW; r.ih = mm % g.IH; mm /= g.IH; r.id = mm % g.ID; int n = mm / g.ID;
    r.ybase = (long long)n * g.OD * g.OH * g.OW;
    r.xoff = (long long)(r.valid ? m : 0) * g.Cin;
    return r;
  }
  __device__ NInfo ninfo(int n, int) const { NInfo r; r.valid = n < g.Cin; r.ci = r.valid ? n : 0; return r; }
  __device__ KInfo kinfo(int k, int kend, int) const {
    KInfo r; r.valid = k < kend;
    int kk = r.valid ? k : 0;
    r.co = kk % g.Cout; int tap = kk / g.Cout;
    r.kw = tap % g.KW; tap /= g.KW; r.kh = tap % g.KH; r.kd = tap / g.KH;
    r.woff = kk * g.Cin;
    return r;
  }
  __device__ float loadA(const MInfo& m, const KInfo& k) const {
    if (!(m.valid && k.valid)) return 0.f;
    int td = m.id + g.pd - k.kd, th = m.ih + g.ph - k.kh, tw = m.iw + g.pw - k.kw;
    if (td < 0 || th < 0 || tw < 0) return 0.f;
    int od = td / g.sd, oh = th / g.sh, ow = tw / g.sw;
    if (od * g.sd != td || oh * g.sh != th || ow * g.sw != tw) return 0.f;
    if (od >= g.OD || oh >= g.OH || ow >= g.OW) return 0.f;
    return to_f(dy[(m.ybase + ((long long)od * g.OH + oh) * g.OW + ow) * g.Cout + k.co]);
  }
  __device__ float loadB(const KInfo& k, const NInfo& n) const {
    return (k.valid && n.valid) ? to_f(w[k.woff + n.ci]) : 0.f;
  }
  __device__ void store(const MInfo& m, const NInfo& n, float v) const {
    if (m.valid && n.valid) dx[m.xoff + n.ci] = from_f<T>(v);
  }
};

// ---------------------------------------------------------------------------------------------
// convolution wgrad (split-K over output pixels, fp32 atomics into the parameter layout)
//   dW[co,ci,tap] = sum_pix dY[pix,co] * X[pix (+) tap, ci]
// ---------------------------------------------------------------------------------------------
template <class T>
struct ConvWgradP {
  static constexpr bool A_K_FAST = false, B_N_FAST = true;
  ConvGeom g;
  const T* dy;  // [pixels][Cout]
  const T* x;   // [N,ID,IH,IW,Cin]
  float* dw;    // [Cout][Cin][taps]  (zero-initialised by the caller)
  int Ncols, Kpix, kchunk, taps;
  struct MInfo { int co; int valid; };
  struct NInfo { int kd, kh, kw, ci, tap; int valid; };
  struct KInfo { long long xbase; long long yoff; int id0, ih0, iw0; int valid; };
  __device__ void krange(int z, int& b, int& e) const {
    b = z * kchunk; e = b + kchunk; if (e > Kpix) e = Kpix; if (b > e) b = e;
  }
  __device__ MInfo minfo(int m, int) const { MInfo r; r.valid = m < g.Cout; r.co = r.valid ? m : 0; return r; }
  __device__ NInfo ninfo(int n, int) const {
    NInfo r; r.valid = n < Ncols;
    int nn = r.valid ? n : 0;
    r.ci = nn % g.Cin; int tap = nn / g.Cin; r.tap = tap;
    r.kw = tap % g.KW; tap /= g.KW; r.kh = tap % g.KH; r.kd = tap / g.KH;
    return r;
  }
  __device__ KInfo kinfo(int k, int kend, int) const {
    KInfo r; r.valid = k < kend;
    int mm = r.valid ? k : 0;
    r.yoff = (long long)mm * g.Cout;
    int ow = mm % g.OW; mm /= g.OW; int oh = mm % g.OH; mm /= g.OH; int od = mm % g.OD; int n = mm / g.OD;
    r.xbase = (long long)n * g.ID * g.IH * g.IW;
    r.id0 = od * g.sd - g.pd; r.ih0 = oh * g.sh - g.ph; r.iw0 = ow * g.sw - g.pw;
    return r;
  }
  __device__ float loadA(const MInfo& m, const KInfo& k) const {
    return (m.valid && k.valid) ? to_f(dy[k.yoff + m.co]) : 0.f;
  }
  __device__ float loadB(const KInfo& k, const NInfo& n) const {
    int id = k.id0 + n.kd, ih = k.ih0 + n.kh, iw = k.iw0 + n.kw;
    if (!(k.valid && n.valid) || (unsigned)id >= (unsigned)g.ID || (unsigned)ih >= (unsigned)g.IH ||
        (unsigned)iw >= (unsigned)g.IW) return 0.f;
    return to_f(x[(k.xbase + ((long long)id * g.IH + ih) * g.IW + iw) * g.Cin + n.ci]);
  }
  __device__ void store(const MInfo& m, const NInfo& n, float v) const {
    if (m.valid && n.valid) atomicAdd(&dw[((long long)m.co * g.Cin + n.ci) * taps + n.tap], v);
  }
};

// ---------------------------------------------------------------------------------------------
// strided batched GEMM:  C[z,m,n] = alpha * sum_k A[z,m,k] B[z,k,n] + bias[n] + beta*C[z,m,n]
// ---------------------------------------------------------------------------------------------
template <class TA, class TB, class TC, bool AK, bool BNF>
struct GemmP {
  static constexpr bool A_K_FAST = AK, B_N_FAST = BNF;
  const TA* A; const TB* B; TC* C; const float* bias;
  long long sAz, sAm, sAk, sBz, sBk, sBn, sCz, sCm, sCn;
  int M, N, K; float alpha, beta;
  struct MInfo { long long a, c; int valid; };
  struct NInfo { long long b, c; int n; int valid; };
  struct KInfo { long long a, b; int valid; };
  __device__ void krange(int, int& b, int& e) const { b = 0; e = K; }
  __device__ MInfo minfo(int m, int z) const {
    MInfo r; r.valid = m < M; int mm = r.valid ? m : 0;
    r.a = z * sAz + mm * sAm; r.c = z * sCz + mm * sCm; return r;
  }
  __device__ NInfo ninfo(int n, int z) const {
    NInfo r; r.valid = n < N; int nn = r.valid ? n : 0; r.n = nn;
    r.b = z * sBz + nn * sBn; r.c = nn * sCn; return r;
  }
  __device__ KInfo kinfo(int k, int kend, int) const {
    KInfo r; r.valid = k < kend; int kk = r.valid ? k : 0; r.a = kk * sAk; r.b = kk * sBk; return r;
  }
  __device__ float loadA(const MInfo& m, const KInfo& k) const { return (m.valid && k.valid) ? to_f(A[m.a + k.a]) : 0.f; }
  __device__ float loadB(const KInfo& k, const NInfo& n) const { return (k.valid && n.valid) ? to_f(B[k.b + n.b]) : 0.f; }
  __device__ void store(const MInfo& m, const NInfo& n, float v) const {
    if (!(m.valid && n.valid)) return;
    float r = alpha * v + (bias ? bias[n.n] : 0.f);
    TC* p = C + m.c + n.c;
    if (beta != 0.f) r += beta * to_f(*p);
    *p = from_f<TC>(r);
  }
};

template <class P>
int launch(const P& p, int M, int N, int Z, cudaStream_t s) {
  if (M <= 0 || N <= 0 || Z <= 0) return VCA_OK;
  dim3 grid((M + BM - 1) / BM, (N + BN - 1) / BN, Z);
  if (grid.y > 65535u || grid.z > 65535u) { vca_set_error("gemm_core: N or batch too large for the grid"); return VCA_ERR_UNSUPPORTED; }
  gemm_core<P><<<grid, NT, 0, s>>>(p);
  VCA_LAUNCH_CHECK();
  return VCA_OK;
}

// ---- weight (re)packing: param layout [Cout][Cin][taps] (fp32) -> packed compute layouts
template <class T>
__global__ void pack_weight_kernel(const float* __restrict__ w, T* __restrict__ wf, T* __restrict__ wd, int Cout, int Cin,
                                   int taps) {
  long long total = (long long)Cout * Cin * taps;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    // i enumerates the packed-forward layout [tap][ci][co] so that writes to wf are coalesced
    int co = (int)(i % Cout); long long r = i / Cout; int ci = (int)(r % Cin); int tap = (int)(r / Cin);
    float v = w[((long long)co * Cin + ci) * taps + tap];
    if (wf) wf[i] = from_f<T>(v);
    if (wd) wd[((long long)tap * Cout + co) * Cin + ci] = from_f<T>(v);
  }
}

// dgrad of a Cin = 1 convolution (the mel inputs of the discriminators / sync discriminator: the R1 gradient and the
// generator's adversarial gradient end here).  N = 1 makes it a per-pixel dot product of length taps*Cout, so the
// tiled GEMM above would waste 63/64 of its tile: one thread per input pixel, weights broadcast from shared memory,
// dY rows read with 128-bit loads.
constexpr int CIN1_MAX_W = 8192;
template <class T, int V>
__global__ void __launch_bounds__(256) conv_cin1_dgrad_kernel(ConvGeom g, const T* __restrict__ dy, const T* __restrict__ wd,
                                                              T* __restrict__ dx, long long M) {
  __shared__ float sw[CIN1_MAX_W];
  const int taps = g.KD * g.KH * g.KW;
  for (int i = threadIdx.x; i < taps * g.Cout; i += blockDim.x) sw[i] = to_f(wd[i]);   // [tap][co]
  __syncthreads();
  for (long long m = blockIdx.x * (long long)blockDim.x + threadIdx.x; m < M; m += (long long)gridDim.x * blockDim.x) {
    long long mm = m;
    const int iw = (int)(mm % g.IW); mm /= g.IW;
    const int ih = (int)(mm % g.IH); mm /= g.IH;
    const int id = (int)(mm % g.ID); const long long n = mm / g.ID;
    float acc = 0.f;
    for (int kd = 0; kd < g.KD; ++kd) {
      const int td = id + g.pd - kd;
      if (td < 0 || td % g.sd) continue;
      const int od = td / g.sd;
      if (od >= g.OD) continue;
      for (int kh = 0; kh < g.KH; ++kh) {
        const int th = ih + g.ph - kh;
        if (th < 0 || th % g.sh) continue;
        const int oh = th / g.sh;
        if (oh >= g.OH) continue;
        for (int kw = 0; kw < g.KW; ++kw) {
          const int tw = iw + g.pw - kw;
          if (tw < 0 || tw % g.sw) continue;
          const int ow = tw / g.sw;
          if (ow >= g.OW) continue;
          const T* row = dy + (((n * g.OD + od) * g.OH + oh) * g.OW + ow) * (long long)g.Cout;
          const float* wrow = sw + ((kd * g.KH + kh) * g.KW + kw) * g.Cout;
          if (V > 1) {
            for (int c = 0; c < g.Cout; c += V) {
              float v[V > 1 ? V : 1];
              if (V == 8) {
                const uint4 t = *reinterpret_cast<const uint4*>(row + c);
                const uint32_t u[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) { v[2 * i] = __uint_as_float(u[i] << 16); v[2 * i + 1] = __uint_as_float(u[i] & 0xffff0000u); }
              } else {
                const float4 t = *reinterpret_cast<const float4*>(row + c);
                v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
              }
#pragma unroll
              for (int i = 0; i < V; ++i) acc = fmaf(v[i], wrow[c + i], acc);
            }
          } else {
            for (int c = 0; c < g.Cout; ++c) acc = fmaf(to_f(row[c]), wrow[c], acc);
          }
        }
      }
    }
    dx[m] = from_f<T>(acc);
  }
}

}  // namespace

static bool geom_ok(const ConvGeom& g) {
  return g.N > 0 && g.Cin > 0 && g.Cout > 0 && g.KD > 0 && g.KH > 0 && g.KW > 0 && g.sd > 0 && g.sh > 0 && g.sw > 0 &&
         g.OD == (g.ID + 2 * g.pd - g.KD) / g.sd + 1 && g.OH == (g.IH + 2 * g.ph - g.KH) / g.sh + 1 &&
         g.OW == (g.IW + 2 * g.pw - g.KW) / g.sw + 1 && g.OD > 0 && g.OH > 0 && g.OW > 0;
}

template <class T>
static int conv_fwd_t(const ConvGeom& g, const void* x, const void* wf, const float* bias, void* y, cudaStream_t s) {
  {
    const int dt = sizeof(T) == 4 ? VCA_F32 : VCA_BF16;
    int r = conv_cin1_fwd(dt, g, x, wf, bias, y, s);
    if (r == 0) r = conv_pw1_fwd(dt, g, x, wf, bias, y, s);
    if (r != 0) return r < 0 ? r : VCA_OK;
  }
  ConvFwdP<T> p; p.g = g; p.x = (const T*)x; p.w = (const T*)wf; p.bias = bias; p.y = (T*)y;
  long long M = (long long)g.N * g.OD * g.OH * g.OW;
  VCA_CHECK_ARG(M < (1ll << 31));
  p.M = (int)M; p.K = g.KD * g.KH * g.KW * g.Cin;
  return launch(p, p.M, g.Cout, 1, s);
}
template <class T>
static int conv_dgrad_t(const ConvGeom& g, const void* dy, const void* wd, void* dx, cudaStream_t s) {
  {
    int r = conv_pw1_dgrad(sizeof(T) == 4 ? VCA_F32 : VCA_BF16, g, dy, wd, dx, s);
    if (r == 0) r = conv_c32_dgrad(sizeof(T) == 4 ? VCA_F32 : VCA_BF16, g, dy, wd, dx, s);
    if (r != 0) return r < 0 ? r : VCA_OK;
  }
  long long M = (long long)g.N * g.ID * g.IH * g.IW;
  if (g.Cin == 1 && g.KD * g.KH * g.KW * g.Cout <= CIN1_MAX_W) {
    constexpr int V = sizeof(T) == 2 ? 8 : 4;
    const bool vec = g.Cout % V == 0 && ((reinterpret_cast<uintptr_t>(dy) & 15) == 0);
    unsigned grid = vca_grid_1d(M, 256);
    if (vec) conv_cin1_dgrad_kernel<T, V><<<grid, 256, 0, s>>>(g, (const T*)dy, (const T*)wd, (T*)dx, M);
    else conv_cin1_dgrad_kernel<T, 1><<<grid, 256, 0, s>>>(g, (const T*)dy, (const T*)wd, (T*)dx, M);
    VCA_LAUNCH_CHECK();
    return VCA_OK;
  }
  ConvDgradP<T> p; p.g = g; p.dy = (const T*)dy; p.w = (const T*)wd; p.dx = (T*)dx;
  VCA_CHECK_ARG(M < (1ll << 31));
  p.M = (int)M; p.K = g.KD * g.KH * g.KW * g.Cout;
  return launch(p, p.M, g.Cin, 1, s);
}
template <class T>
static int conv_wgrad_t(const ConvGeom& g, const void* dy, const void* x, float* dw, cudaStream_t s) {
  {
    const int dt = sizeof(T) == 4 ? VCA_F32 : VCA_BF16;
    int r = conv_cin1_wgrad(dt, g, dy, x, dw, s);
    if (r == 0) r = conv_pw1_wgrad(dt, g, dy, x, dw, s);
    if (r != 0) return r < 0 ? r : VCA_OK;
  }
  ConvWgradP<T> p; p.g = g; p.dy = (const T*)dy; p.x = (const T*)x; p.dw = dw;
  long long Kp = (long long)g.N * g.OD * g.OH * g.OW;
  VCA_CHECK_ARG(Kp < (1ll << 31));
  p.taps = g.KD * g.KH * g.KW; p.Ncols = p.taps * g.Cin; p.Kpix = (int)Kp;
  int tiles = ((g.Cout + BM - 1) / BM) * ((p.Ncols + BN - 1) / BN);
  int want = (4 * vca_num_sms() + tiles - 1) / tiles;          // fill the chip ~4 CTAs/SM
  int maxsplit = (p.Kpix + 4 * BK - 1) / (4 * BK);
  int split = want < 1 ? 1 : (want > maxsplit ? maxsplit : want);
  if (split > 4096) split = 4096;
  p.kchunk = (((p.Kpix + split - 1) / split) + BK - 1) / BK * BK;
  split = (p.Kpix + p.kchunk - 1) / p.kchunk;
  return launch(p, g.Cout, p.Ncols, split, s);
}

extern "C" {

// Convolution forward on channels-last tensors; wf = packed [taps][Cin][Cout] (see vca_pack_conv_weight).
int vca_conv_fwd_simt(int dtype, const ConvGeom* g, const void* x, const void* wf, const float* bias, void* y,
                      cudaStream_t s) {
  VCA_CHECK_ARG(g && x && wf && y && geom_ok(*g));
  return dtype == VCA_F32 ? conv_fwd_t<float>(*g, x, wf, bias, y, s) : conv_fwd_t<bf16>(*g, x, wf, bias, y, s);
}
int vca_conv_dgrad_simt(int dtype, const ConvGeom* g, const void* dy, const void* wd, void* dx, cudaStream_t s) {
  VCA_CHECK_ARG(g && dy && wd && dx && geom_ok(*g));
  return dtype == VCA_F32 ? conv_dgrad_t<float>(*g, dy, wd, dx, s) : conv_dgrad_t<bf16>(*g, dy, wd, dx, s);
}
// dw: fp32 [Cout][Cin][taps], must be zero on entry (accumulated with atomics).
int vca_conv_wgrad_simt(int dtype, const ConvGeom* g, const void* dy, const void* x, float* dw, cudaStream_t s) {
  VCA_CHECK_ARG(g && dy && x && dw && geom_ok(*g));
  return dtype == VCA_F32 ? conv_wgrad_t<float>(*g, dy, x, dw, s) : conv_wgrad_t<bf16>(*g, dy, x, dw, s);
}
// w: fp32 parameter [Cout][Cin][taps]; wf: [taps][Cin][Cout]; wd: [taps][Cout][Cin] (either may be null).
int vca_pack_conv_weight(int dtype, const float* w, void* wf, void* wd, int Cout, int Cin, int taps, cudaStream_t s) {
  VCA_CHECK_ARG(w && (wf || wd) && Cout > 0 && Cin > 0 && taps > 0);
  if (Cout >= 8 && Cin >= 8) {   // tile-transpose kernel (coalesced on both sides); tiny layers use the simple one below
    const int r = pack_weight_tiled(dtype, w, wf, wd, Cout, Cin, taps, s);
    if (r != 0) return r < 0 ? r : VCA_OK;
  }
  long long total = (long long)Cout * Cin * taps;
  unsigned grid = vca_grid_1d(total, 256);
  if (dtype == VCA_F32) pack_weight_kernel<float><<<grid, 256, 0, s>>>(w, (float*)wf, (float*)wd, Cout, Cin, taps);
  else pack_weight_kernel<bf16><<<grid, 256, 0, s>>>(w, (bf16*)wf, (bf16*)wd, Cout, Cin, taps);
  VCA_LAUNCH_CHECK();
  return VCA_OK;
}

// Strided batched GEMM; element strides; dtA/dtB/dtC in {VCA_F32, VCA_BF16}; fp32 accumulate.
int vca_gemm_simt(int dtA, int dtB, int dtC, const void* A, const void* B, void* C, const float* bias, int Z, int M, int N,
                  int K, long long sAz, long long sAm, long long sAk, long long sBz, long long sBk, long long sBn,
                  long long sCz, long long sCm, long long sCn, float alpha, float beta, cudaStream_t s) {
  VCA_CHECK_ARG(A && B && C && K >= 0);
#define VCA_GEMM_GO(TA, TB, TC, AK, BNF)                                                    \
  {                                                                                          \
    GemmP<TA, TB, TC, AK, BNF> p;                                                            \
    p.A = (const TA*)A; p.B = (const TB*)B; p.C = (TC*)C; p.bias = bias;                     \
    p.sAz = sAz; p.sAm = sAm; p.sAk = sAk; p.sBz = sBz; p.sBk = sBk; p.sBn = sBn;           \
    p.sCz = sCz; p.sCm = sCm; p.sCn = sCn; p.M = M; p.N = N; p.K = K; p.alpha = alpha; p.beta = beta; \
    return launch(p, M, N, Z, s);                                                            \
  }
#define VCA_GEMM_FAST(TA, TB, TC)                                     \
  {                                                                   \
    bool ak = (sAk == 1) || (sAm != 1);                               \
    bool bn = (sBn == 1) || (sBk != 1);                               \
    if (ak && bn) VCA_GEMM_GO(TA, TB, TC, true, true)                 \
    else if (ak) VCA_GEMM_GO(TA, TB, TC, true, false)                 \
    else if (bn) VCA_GEMM_GO(TA, TB, TC, false, true)                 \
    else VCA_GEMM_GO(TA, TB, TC, false, false)                        \
  }
  if (dtA == VCA_F32 && dtB == VCA_F32 && dtC == VCA_F32) VCA_GEMM_FAST(float, float, float)
  if (dtA == VCA_BF16 && dtB == VCA_BF16 && dtC == VCA_BF16) VCA_GEMM_FAST(bf16, bf16, bf16)
  if (dtA == VCA_BF16 && dtB == VCA_BF16 && dtC == VCA_F32) VCA_GEMM_FAST(bf16, bf16, float)
  if (dtA == VCA_BF16 && dtB == VCA_F32 && dtC == VCA_BF16) VCA_GEMM_FAST(bf16, float, bf16)
  if (dtA == VCA_BF16 && dtB == VCA_F32 && dtC == VCA_F32) VCA_GEMM_FAST(bf16, float, float)
  if (dtA == VCA_F32 && dtB == VCA_BF16 && dtC == VCA_F32) VCA_GEMM_FAST(float, bf16, float)
  if (dtA == VCA_F32 && dtB == VCA_F32 && dtC == VCA_BF16) VCA_GEMM_FAST(float, float, bf16)
  vca_set_error("vca_gemm_simt: unsupported dtype combination %d %d %d", dtA, dtB, dtC);
  return VCA_ERR_UNSUPPORTED;
}

}  // extern "C"
