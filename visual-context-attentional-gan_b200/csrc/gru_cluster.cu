// GRU recurrence (nn.GRU of visual_front.py:20,33-34) on thread-block clusters with distributed shared memory, fp32.
//
// The T time steps are strictly sequential and each is a tiny GEMM (B x H by H x 3H), so the step time is all
// latency: where does h_{t-1} live and what does it cost to hand h_t to whoever needs it next.  Here one cluster of 8
// CTAs owns one (direction, 8-row batch slice): batch rows never interact, so clusters are completely independent
// and there is NO grid-wide barrier.  Inside a cluster every CTA owns H/8 hidden units, keeps their 3 rows of W_hh
// (forward) / columns of W_hh (backward) in shared memory for the whole sequence, and keeps a full copy of the
// slice's h_{t-1} (backward: of this step's gate gradients) in its own shared memory.  A step is
//     matvec from shared memory (register tile 3 gates x 4 batch rows, K split over warps)
//     -> gate math -> the new values are PUSHED into all 8 CTAs' shared memory (st.shared::cluster)
//     -> one hardware cluster barrier (release/acquire),
// double-buffered so a single barrier per step is enough.  Nothing on the recurrence path touches L2 or HBM: the
// input projections / saved gates are prefetched one step ahead and the outputs are fire-and-forget stores.
#include "common.cuh"
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

int g_gru_cluster = 1;   // "gru_cluster" in vca_set_option: 0 = use the cooperative-grid kernels of gru_persistent.cu

namespace {

constexpr int CS = 8;     // CTAs per cluster (portable maximum)
constexpr int BS = 8;     // batch rows per cluster
constexpr int NT = 256;   // threads per CTA

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + __expf(-x)); }

#define FMA4(ACC, W, X) ACC = fmaf(W.x, X.x, ACC); ACC = fmaf(W.y, X.y, ACC); ACC = fmaf(W.z, X.z, ACC); ACC = fmaf(W.w, X.w, ACC);

// gi [ndir][T][B][3H] (incl. b_ih); whh [ndir][3H][H]; bhh [ndir][3H]; out [T][B][ndir*H];
// gates [ndir][T][B][4H] = r, z, n, hn.  Grid = ndir * nbs clusters of CS CTAs.
__global__ void __launch_bounds__(NT, 1)
gru_cluster_fwd_kernel(const float* __restrict__ gi, const float* __restrict__ whh, const float* __restrict__ bhh,
                       float* __restrict__ out, float* __restrict__ gates, int ndir, int nbs, int T, int B, int H, int KS,
                       int per) {
  cg::cluster_group cluster = cg::this_cluster();
  extern __shared__ __align__(16) float sm[];
  const int HP = H + 4, UPC = H / CS, tiles = UPC * 2;
  float* sW = sm;                      // [3*UPC][HP]   row = gate * UPC + unit
  float* sH = sW + 3 * UPC * HP;       // [2][BS][HP]   ping-pong copies of the slice's h
  float* sP = sH + 2 * BS * HP;        // [KS][tiles][12] K-split partial sums
  const int rank = (int)cluster.block_rank();
  const int cid = blockIdx.x / CS;
  const int d = cid / nbs, bbase = (cid % nbs) * BS;
  const int j0 = rank * UPC;
  const int tid = threadIdx.x;
  const float* wd = whh + (size_t)d * 3 * H * H;
  for (int i = tid; i < 3 * UPC * (H >> 2); i += NT) {
    const int r = i / (H >> 2), c4 = i - r * (H >> 2);
    const int g = r / UPC, u = r - g * UPC;
    *reinterpret_cast<float4*>(sW + r * HP + c4 * 4) =
        *reinterpret_cast<const float4*>(wd + (size_t)(g * H + j0 + u) * H + c4 * 4);
  }
  for (int i = tid; i < 2 * BS * HP; i += NT) sH[i] = 0.f;   // h_0 = 0
  // matvec role: tile = (unit mu, batch rows {bg, 2+bg, 4+bg, 6+bg}), K range [k4lo, k4hi) in float4 units
  const int tile = tid % tiles, ks = tid / tiles;
  const bool mv = ks < KS;
  const int mu = tile >> 1, bg = tile & 1;
  const int k4lo = ks * per, k4hi = min(k4lo + per, H >> 2);
  // gate role: one thread per (unit fu, batch row fb)
  const bool fin = tid < UPC * BS;
  const int fu = tid % UPC, fb = tid / UPC;
  const int j = j0 + fu, b = bbase + fb;
  const bool live = fin && b < B;
  float br = 0.f, bz = 0.f, bn = 0.f;
  if (fin) { br = bhh[d * 3 * H + j]; bz = bhh[d * 3 * H + H + j]; bn = bhh[d * 3 * H + 2 * H + j]; }
  float* rH[CS];
#pragma unroll
  for (int r = 0; r < CS; ++r) rH[r] = cluster.map_shared_rank(sH, r);
  float g_r = 0.f, g_z = 0.f, g_n = 0.f;   // input projections of the current step (prefetched)
  if (live) {
    const float* gp = gi + (((size_t)d * T + (d == 0 ? 0 : T - 1)) * B + b) * 3 * H;
    g_r = gp[j]; g_z = gp[H + j]; g_n = gp[2 * H + j];
  }
  cluster.sync();   // every CTA of the cluster is running and has initialised its shared memory
  for (int s = 0; s < T; ++s) {
    const int t = d == 0 ? s : T - 1 - s;
    const float* hc = sH + (s & 1) * BS * HP;
    float n_r = 0.f, n_z = 0.f, n_n = 0.f;
    if (live && s + 1 < T) {
      const float* gp = gi + (((size_t)d * T + (d == 0 ? s + 1 : T - 2 - s)) * B + b) * 3 * H;
      n_r = gp[j]; n_z = gp[H + j]; n_n = gp[2 * H + j];
    }
    if (mv) {
      float acc[12];
#pragma unroll
      for (int i = 0; i < 12; ++i) acc[i] = 0.f;
      const float4* w0 = reinterpret_cast<const float4*>(sW + (0 * UPC + mu) * HP);
      const float4* w1 = reinterpret_cast<const float4*>(sW + (1 * UPC + mu) * HP);
      const float4* w2 = reinterpret_cast<const float4*>(sW + (2 * UPC + mu) * HP);
      const float4* h0 = reinterpret_cast<const float4*>(hc + (0 + bg) * HP);
      const float4* h1 = reinterpret_cast<const float4*>(hc + (2 + bg) * HP);
      const float4* h2 = reinterpret_cast<const float4*>(hc + (4 + bg) * HP);
      const float4* h3 = reinterpret_cast<const float4*>(hc + (6 + bg) * HP);
#pragma unroll 2
      for (int k = k4lo; k < k4hi; ++k) {
        const float4 a = w0[k], bq = w1[k], c = w2[k];
        const float4 x0 = h0[k], x1 = h1[k], x2 = h2[k], x3 = h3[k];
        FMA4(acc[0], a, x0) FMA4(acc[1], a, x1) FMA4(acc[2], a, x2) FMA4(acc[3], a, x3)
        FMA4(acc[4], bq, x0) FMA4(acc[5], bq, x1) FMA4(acc[6], bq, x2) FMA4(acc[7], bq, x3)
        FMA4(acc[8], c, x0) FMA4(acc[9], c, x1) FMA4(acc[10], c, x2) FMA4(acc[11], c, x3)
      }
      float4* pp = reinterpret_cast<float4*>(sP + ((size_t)ks * tiles + tile) * 12);
      pp[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
      pp[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
      pp[2] = make_float4(acc[8], acc[9], acc[10], acc[11]);
    }
    __syncthreads();
    if (fin) {
      float ar = 0.f, az = 0.f, an = 0.f;
      const int ftile = fu * 2 + (fb & 1), fi = fb >> 1;
      for (int q = 0; q < KS; ++q) {
        const float* pp = sP + ((size_t)q * tiles + ftile) * 12;
        ar += pp[fi]; az += pp[4 + fi]; an += pp[8 + fi];
      }
      float h = 0.f;
      if (live) {
        const float rr = sigmoidf_(g_r + ar + br);
        const float zz = sigmoidf_(g_z + az + bz);
        const float hn = an + bn;
        const float nn = tanhf(g_n + rr * hn);
        h = (1.f - zz) * nn + zz * hc[fb * HP + j];
        out[((size_t)t * B + b) * (ndir * H) + d * H + j] = h;
        float* gs = gates + (((size_t)d * T + t) * B + b) * 4 * H;
        gs[j] = rr; gs[H + j] = zz; gs[2 * H + j] = nn; gs[3 * H + j] = hn;
      }
      const int off = ((s + 1) & 1) * BS * HP + fb * HP + j;
#pragma unroll
      for (int r = 0; r < CS; ++r) rH[r][off] = h;
    }
    g_r = n_r; g_z = n_z; g_n = n_n;
    cluster.sync();   // h_t is complete in every CTA; everyone is done reading h_{t-1}
  }
}

// Backward through time.  dout [T][B][ndir*H]; dgi/dgh [ndir][T][B][3H].
__global__ void __launch_bounds__(NT, 1)
gru_cluster_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ whh, const float* __restrict__ gates,
                       const float* __restrict__ out, float* __restrict__ dgi, float* __restrict__ dgh, int ndir, int nbs,
                       int T, int B, int H, int KS, int per) {
  cg::cluster_group cluster = cg::this_cluster();
  extern __shared__ __align__(16) float sm[];
  const int GP = 3 * H + 4, UPC = H / CS, tiles = UPC * 2;
  float* sWt = sm;                     // [UPC][GP]     sWt[u][row] = W_hh[row][j0 + u]
  float* sG = sWt + UPC * GP;          // [2][BS][GP]   ping-pong copies of the slice's dgh of the current step
  float* sP = sG + 2 * BS * GP;        // [KS][tiles][4]
  const int rank = (int)cluster.block_rank();
  const int cid = blockIdx.x / CS;
  const int d = cid / nbs, bbase = (cid % nbs) * BS;
  const int j0 = rank * UPC;
  const int tid = threadIdx.x;
  const float* wd = whh + (size_t)d * 3 * H * H;
  for (int i = tid; i < 3 * H * UPC; i += NT) {
    const int row = i / UPC, u = i - row * UPC;
    sWt[u * GP + row] = wd[(size_t)row * H + j0 + u];
  }
  for (int i = tid; i < 2 * BS * GP; i += NT) sG[i] = 0.f;
  const int tile = tid % tiles, ks = tid / tiles;
  const bool mv = ks < KS;
  const int mu = tile >> 1, bg = tile & 1;
  const int k4lo = ks * per, k4hi = min(k4lo + per, (3 * H) >> 2);
  const bool fin = tid < UPC * BS;
  const int fu = tid % UPC, fb = tid / UPC;
  const int j = j0 + fu, b = bbase + fb;
  const bool live = fin && b < B;
  float* rG[CS];
#pragma unroll
  for (int r = 0; r < CS; ++r) rG[r] = cluster.map_shared_rank(sG, r);
  // saved values of the current step (prefetched one step ahead): r, z, n, hn, h_prev, dout
  float c_r = 0.f, c_z = 0.f, c_n = 0.f, c_hn = 0.f, c_hp = 0.f, c_do = 0.f;
  auto fetch = [&](int s, float& r_, float& z_, float& n_, float& hn_, float& hp_, float& do_) {
    const int t = d == 0 ? T - 1 - s : s;           // reverse of the forward order
    const int tp = d == 0 ? t - 1 : t + 1;
    const float* gs = gates + (((size_t)d * T + t) * B + b) * 4 * H;
    r_ = gs[j]; z_ = gs[H + j]; n_ = gs[2 * H + j]; hn_ = gs[3 * H + j];
    hp_ = (tp >= 0 && tp < T) ? out[((size_t)tp * B + b) * (ndir * H) + d * H + j] : 0.f;
    do_ = dout[((size_t)t * B + b) * (ndir * H) + d * H + j];
  };
  if (live) fetch(0, c_r, c_z, c_n, c_hn, c_hp, c_do);
  float dh_carry = 0.f;
  cluster.sync();
  for (int s = 0; s < T; ++s) {
    const int t = d == 0 ? T - 1 - s : s;
    float dhz = 0.f;
    if (fin) {
      float drp = 0.f, dzp = 0.f, dnr = 0.f;
      if (live) {
        const float dh = c_do + dh_carry;
        const float dnp = dh * (1.f - c_z) * (1.f - c_n * c_n);
        drp = dnp * c_hn * c_r * (1.f - c_r);
        dzp = dh * (c_hp - c_n) * c_z * (1.f - c_z);
        dnr = dnp * c_r;
        dhz = dh * c_z;
        const size_t go = (((size_t)d * T + t) * B + b) * 3 * H;
        dgi[go + j] = drp; dgi[go + H + j] = dzp; dgi[go + 2 * H + j] = dnp;
        dgh[go + j] = drp; dgh[go + H + j] = dzp; dgh[go + 2 * H + j] = dnr;
      }
      const int off = (s & 1) * BS * GP + fb * GP + j;
#pragma unroll
      for (int r = 0; r < CS; ++r) { float* p = rG[r] + off; p[0] = drp; p[H] = dzp; p[2 * H] = dnr; }
    }
    float n_r = 0.f, n_z = 0.f, n_n = 0.f, n_hn = 0.f, n_hp = 0.f, n_do = 0.f;
    if (live && s + 1 < T) fetch(s + 1, n_r, n_z, n_n, n_hn, n_hp, n_do);
    cluster.sync();   // this step's dgh is complete in every CTA
    // dh_{t-1}[b][k] = dh*z + sum_row dgh[b][row] * W_hh[row][k] for my units k
    if (mv) {
      const float* gc = sG + (s & 1) * BS * GP;
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
      const float4* w4 = reinterpret_cast<const float4*>(sWt + mu * GP);
      const float4* g0 = reinterpret_cast<const float4*>(gc + (0 + bg) * GP);
      const float4* g1 = reinterpret_cast<const float4*>(gc + (2 + bg) * GP);
      const float4* g2 = reinterpret_cast<const float4*>(gc + (4 + bg) * GP);
      const float4* g3 = reinterpret_cast<const float4*>(gc + (6 + bg) * GP);
#pragma unroll 4
      for (int k = k4lo; k < k4hi; ++k) {
        const float4 a = w4[k];
        const float4 x0 = g0[k], x1 = g1[k], x2 = g2[k], x3 = g3[k];
        FMA4(a0, a, x0) FMA4(a1, a, x1) FMA4(a2, a, x2) FMA4(a3, a, x3)
      }
      *reinterpret_cast<float4*>(sP + ((size_t)ks * tiles + tile) * 4) = make_float4(a0, a1, a2, a3);
    }
    __syncthreads();
    if (fin) {
      float acc = 0.f;
      const int ftile = fu * 2 + (fb & 1), fi = fb >> 1;
      for (int q = 0; q < KS; ++q) acc += sP[((size_t)q * tiles + ftile) * 4 + fi];
      dh_carry = acc + dhz;
    }
    c_r = n_r; c_z = n_z; c_n = n_n; c_hn = n_hn; c_hp = n_hp; c_do = n_do;
  }
  cluster.sync();   // no CTA may exit while a peer could still write into its shared memory
}

struct ClusterPlan { int nbs, KS, per_f, per_b; size_t smem_f, smem_b; };

bool plan_cluster(int B, int H, ClusterPlan& p) {
  if (!g_gru_cluster || H % CS != 0 || H % 4 != 0) return false;   // whole units per CTA, float4 rows
  const int UPC = H / CS, tiles = UPC * 2;
  if (tiles > NT || UPC * BS > NT) return false;
  p.nbs = (B + BS - 1) / BS;
  p.KS = NT / tiles; if (p.KS > 8) p.KS = 8;
  p.per_f = ((H >> 2) + p.KS - 1) / p.KS;
  p.per_b = (((3 * H) >> 2) + p.KS - 1) / p.KS;
  p.smem_f = sizeof(float) * ((size_t)3 * UPC * (H + 4) + 2 * BS * (H + 4) + (size_t)p.KS * tiles * 12);
  p.smem_b = sizeof(float) * ((size_t)UPC * (3 * H + 4) + 2 * BS * (3 * H + 4) + (size_t)p.KS * tiles * 4);
  return p.smem_f <= 220 * 1024 && p.smem_b <= 220 * 1024;
}

template <class K, class... Args>
int launch_cluster(K kernel, int grid, size_t smem, cudaStream_t s, const char* what, Args... args) {
  if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(NT); cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CS; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  int nclusters = 0;
  if (cudaOccupancyMaxActiveClusters(&nclusters, kernel, &cfg) != cudaSuccess || nclusters < 1) { cudaGetLastError(); return 0; }
  if (cudaLaunchKernelEx(&cfg, kernel, args...) != cudaSuccess) {
    vca_set_error("%s: cluster launch failed: %s", what, cudaGetErrorString(cudaGetLastError()));
    return VCA_ERR_CUDA;
  }
  return 1;
}

}  // namespace

// 1 = launched, 0 = shape / device not applicable (caller falls through to the cooperative-grid kernels), < 0 = error
int gru_cluster_fwd_try(const float* gi, const float* whh, const float* bhh, float* out, float* gates, int ndir, int T, int B,
                        int H, cudaStream_t s) {
  ClusterPlan p;
  if (!plan_cluster(B, H, p)) return 0;
  return launch_cluster(gru_cluster_fwd_kernel, ndir * p.nbs * CS, p.smem_f, s, "vca_gru_seq_fwd", gi, whh, bhh, out, gates,
                        ndir, p.nbs, T, B, H, p.KS, p.per_f);
}
int gru_cluster_bwd_try(const float* dout, const float* whh, const float* gates, const float* out, float* dgi, float* dgh,
                        int ndir, int T, int B, int H, cudaStream_t s) {
  ClusterPlan p;
  if (!plan_cluster(B, H, p)) return 0;
  return launch_cluster(gru_cluster_bwd_kernel, ndir * p.nbs * CS, p.smem_b, s, "vca_gru_seq_bwd", dout, whh, gates, out, dgi,
                        dgh, ndir, p.nbs, T, B, H, p.KS, p.per_b);
}
