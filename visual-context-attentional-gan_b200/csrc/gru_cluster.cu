// GRU recurrence (nn.GRU of visual_front.py:20,33-34) on thread-block clusters with distributed shared memory, fp32.
//
// The T time steps are strictly sequential and each is a tiny GEMM (B x H by H x 3H), so the step time is all
// latency: where does h_{t-1} live and what does it cost to hand h_t to whoever needs it next.  Here one cluster of
// CS = 8 or 16 CTAs owns one (direction, 8-row batch slice): batch rows never interact, so clusters are completely
// independent and there is NO grid-wide barrier.  Inside a cluster every CTA owns H/CS hidden units and keeps the 3
// rows of W_hh of each of them (fp32, up to 198 KB) in shared memory for the whole sequence.
//   forward : every CTA holds a full copy of the slice's h_{t-1}; a step is  matvec from shared memory (register tile
//             3 gates x 4 batch rows, K split over warps) -> gate math -> the new h values are PUSHED into all CS
//             CTAs' shared memory (st.shared::cluster) -> hardware cluster barrier.
//   backward: the same W rows give this CTA's PARTIAL  sum_{my rows} dgh[b][row] * W[row][k]  for all H columns k;
//             the partials are pushed to the CTA that owns column k (a reduce-scatter through distributed shared
//             memory, summed in rank order, so the result is deterministic) -> cluster barrier -> dh_{t-1}.
// The "everyone is done reading" barrier is split (arrive early, wait late) so that only one barrier latency per
// step is exposed.  Nothing on the recurrence path touches L2 or HBM: the input projections / saved gates are
// prefetched one step ahead and the outputs are fire-and-forget stores.
#include "common.cuh"
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

int g_gru_cluster = 1;   // "gru_cluster" in vca_set_option: 0 = use the cooperative-grid kernels of gru_persistent.cu
int g_gru_bs = 0;        // "gru_bs": force the batch rows per cluster (8 | 12 | 16), 0 = choose by the wave model

namespace {

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + __expf(-x)); }
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_arrive_relaxed() { asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }

#define FMA4(ACC, W, X) ACC = fmaf(W.x, X.x, ACC); ACC = fmaf(W.y, X.y, ACC); ACC = fmaf(W.z, X.z, ACC); ACC = fmaf(W.w, X.w, ACC);

// sW[(g * UPC + u) * HP + k] = W_hh[d][g * H + j0 + u][k]
__device__ __forceinline__ void load_w_rows(float* sW, const float* __restrict__ wd, int H, int HP, int UPC, int j0) {
  for (int i = threadIdx.x; i < 3 * UPC * (H >> 2); i += blockDim.x) {
    const int r = i / (H >> 2), c4 = i - r * (H >> 2);
    const int g = r / UPC, u = r - g * UPC;
    *reinterpret_cast<float4*>(sW + r * HP + c4 * 4) =
        *reinterpret_cast<const float4*>(wd + (size_t)(g * H + j0 + u) * H + c4 * 4);
  }
}

// gi [ndir][T][B][3H] (incl. b_ih); whh [ndir][3H][H]; bhh [ndir][3H]; out [T][B][ndir*H];
// gates [ndir][T][B][4H] = r, z, n, hn.  Grid = ndir * nbs clusters of CS CTAs of 32*BS threads; a cluster owns BS
// batch rows (BS = 8 | 12 | 16, chosen so that all clusters are co-resident).
template <int CS, int BS>
__global__ void __launch_bounds__(32 * BS, 1)
gru_cluster_fwd_kernel(const float* __restrict__ gi, const float* __restrict__ whh, const float* __restrict__ bhh,
                       float* __restrict__ out, float* __restrict__ gates, int ndir, int nbs, int T, int B, int H, int KS,
                       int per) {
  constexpr int NBG = BS / 4;          // groups of 4 batch rows: a matvec thread owns rows {bg, NBG+bg, 2NBG+bg, 3NBG+bg}
  cg::cluster_group cluster = cg::this_cluster();
  extern __shared__ __align__(16) float sm[];
  const int HP = H + 4, UPC = H / CS, tiles = UPC * NBG, half = KS >= 2 ? KS / 2 : 1;
  float* sW = sm;                      // [3*UPC][HP]   row = gate * UPC + unit
  float* sH = sW + 3 * UPC * HP;       // [BS][HP]      this CTA's copy of the slice's h_{t-1}
  float* sP = sH + BS * HP;            // [half][tiles][12] K-split partial sums (reduced in two stages)
  const int rank = (int)cluster.block_rank();
  const int cid = blockIdx.x / CS;
  const int d = cid / nbs, bbase = (cid % nbs) * BS;
  const int j0 = rank * UPC;
  const int tid = threadIdx.x;
  load_w_rows(sW, whh + (size_t)d * 3 * H * H, H, HP, UPC, j0);
  for (int i = tid; i < BS * HP; i += blockDim.x) sH[i] = 0.f;   // h_0 = 0
  // matvec role: tile = (unit mu, batch group bg), K range [k4lo, k4hi) in float4 units
  const int tile = tid % tiles, ks = tid / tiles;
  const bool mv = ks < KS;
  const int mu = tile / NBG, bg = tile - mu * NBG;
  const int k4lo = ks * per, k4hi = min(k4lo + per, H >> 2);
  // gate role: one thread per (unit fu, batch row fb)
  const bool fin = tid < UPC * BS;
  const int fu = tid % UPC, fb = tid / UPC;
  const int j = j0 + fu, b = bbase + fb;
  const bool live = fin && b < B;
  float br = 0.f, bz = 0.f, bn = 0.f;
  if (fin) { br = bhh[d * 3 * H + j]; bz = bhh[d * 3 * H + H + j]; bn = bhh[d * 3 * H + 2 * H + j]; }
  float g_r = 0.f, g_z = 0.f, g_n = 0.f;   // input projections of the current step (prefetched)
  if (live) {
    const float* gp = gi + (((size_t)d * T + (d == 0 ? 0 : T - 1)) * B + b) * 3 * H;
    g_r = gp[j]; g_z = gp[H + j]; g_n = gp[2 * H + j];
  }
  cluster_arrive(); cluster_wait();   // every CTA of the cluster is running and has initialised its shared memory
  for (int s = 0; s < T; ++s) {
    const int t = d == 0 ? s : T - 1 - s;
    float n_r = 0.f, n_z = 0.f, n_n = 0.f;
    if (live && s + 1 < T) {
      const float* gp = gi + (((size_t)d * T + (d == 0 ? s + 1 : T - 2 - s)) * B + b) * 3 * H;
      n_r = gp[j]; n_z = gp[H + j]; n_n = gp[2 * H + j];
    }
    const float hprev = fin ? sH[fb * HP + j] : 0.f;
    float acc[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) acc[i] = 0.f;
    if (mv) {
      const float4* w0 = reinterpret_cast<const float4*>(sW + (0 * UPC + mu) * HP);
      const float4* w1 = reinterpret_cast<const float4*>(sW + (1 * UPC + mu) * HP);
      const float4* w2 = reinterpret_cast<const float4*>(sW + (2 * UPC + mu) * HP);
      const float4* h0 = reinterpret_cast<const float4*>(sH + (0 * NBG + bg) * HP);
      const float4* h1 = reinterpret_cast<const float4*>(sH + (1 * NBG + bg) * HP);
      const float4* h2 = reinterpret_cast<const float4*>(sH + (2 * NBG + bg) * HP);
      const float4* h3 = reinterpret_cast<const float4*>(sH + (3 * NBG + bg) * HP);
#pragma unroll 2
      for (int k = k4lo; k < k4hi; ++k) {
        const float4 a = w0[k], bq = w1[k], c = w2[k];
        const float4 x0 = h0[k], x1 = h1[k], x2 = h2[k], x3 = h3[k];
        FMA4(acc[0], a, x0) FMA4(acc[1], a, x1) FMA4(acc[2], a, x2) FMA4(acc[3], a, x3)
        FMA4(acc[4], bq, x0) FMA4(acc[5], bq, x1) FMA4(acc[6], bq, x2) FMA4(acc[7], bq, x3)
        FMA4(acc[8], c, x0) FMA4(acc[9], c, x1) FMA4(acc[10], c, x2) FMA4(acc[11], c, x3)
      }
    }
    // K-split reduction, stage 1: the upper half of the K slices hands its partials to the lower half
    if (KS >= 2) {
      if (mv && ks >= half) {
        float4* pp = reinterpret_cast<float4*>(sP + ((size_t)(ks - half) * tiles + tile) * 12);
        pp[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
        pp[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
        pp[2] = make_float4(acc[8], acc[9], acc[10], acc[11]);
      }
      __syncthreads();
    }
    if (mv && ks < half) {
      float4* pp = reinterpret_cast<float4*>(sP + ((size_t)ks * tiles + tile) * 12);
      if (KS >= 2) {
        const float4 p0 = pp[0], p1 = pp[1], p2 = pp[2];
        acc[0] += p0.x; acc[1] += p0.y; acc[2] += p0.z; acc[3] += p0.w;
        acc[4] += p1.x; acc[5] += p1.y; acc[6] += p1.z; acc[7] += p1.w;
        acc[8] += p2.x; acc[9] += p2.y; acc[10] += p2.z; acc[11] += p2.w;
      }
      pp[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
      pp[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
      pp[2] = make_float4(acc[8], acc[9], acc[10], acc[11]);
    }
    __syncthreads();
    cluster_arrive_relaxed();   // (A) this thread is done reading h_{t-1} (no data published: execution ordering only)
    float h = 0.f, rr = 0.f, zz = 0.f, nn = 0.f, hn = 0.f;
    if (fin) {
      float ar = 0.f, az = 0.f, an = 0.f;
      const int ftile = fu * NBG + fb % NBG, fi = fb / NBG;
      for (int q = 0; q < half; ++q) {
        const float* pp = sP + ((size_t)q * tiles + ftile) * 12;
        ar += pp[fi]; az += pp[4 + fi]; an += pp[8 + fi];
      }
      if (live) {
        rr = sigmoidf_(g_r + ar + br);
        zz = sigmoidf_(g_z + az + bz);
        hn = an + bn;
        nn = tanhf(g_n + rr * hn);
        h = (1.f - zz) * nn + zz * hprev;
      }
    }
    cluster_wait();     // (A) every CTA is done reading its copy: safe to overwrite
    if (fin) {
      float* dst = sH + fb * HP + j;
#pragma unroll
      for (int r = 0; r < CS; ++r) *cluster.map_shared_rank(dst, r) = h;
    }
    cluster_arrive();   // (B) release: my pushes are visible to whoever completes the wait
    // The global stores go AFTER the release-arrive so it does not have to wait for them to reach L2; they have
    // the whole next step to drain before the next release.
    if (live) {
      out[((size_t)t * B + b) * (ndir * H) + d * H + j] = h;
      float* gs = gates + (((size_t)d * T + t) * B + b) * 4 * H;
      gs[j] = rr; gs[H + j] = zz; gs[2 * H + j] = nn; gs[3 * H + j] = hn;
    }
    g_r = n_r; g_z = n_z; g_n = n_n;
    cluster_wait();     // (B) h_t is complete in every CTA
  }
}

// Backward through time.  dout [T][B][ndir*H]; dgi/dgh [ndir][T][B][3H].
template <int CS, int BS>
__global__ void __launch_bounds__(32 * BS, 1)
gru_cluster_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ whh, const float* __restrict__ gates,
                       const float* __restrict__ out, float* __restrict__ dgi, float* __restrict__ dgh, int ndir, int nbs,
                       int T, int B, int H) {
  constexpr int NT = 32 * BS;
  cg::cluster_group cluster = cg::this_cluster();
  extern __shared__ __align__(16) float sm[];
  const int HP = H + 4, UPC = H / CS, rows = 3 * UPC;
  float* sW = sm;                      // [3*UPC][HP]     row = gate * UPC + unit (same slice as the forward)
  float* sR = sW + rows * HP;          // [CS][BS][UPC]   partial dh of my columns, one slab per source rank
  float* sD = sR + CS * BS * UPC;      // [3*UPC][BS]     this step's dgh of my rows
  const int rank = (int)cluster.block_rank();
  const int cid = blockIdx.x / CS;
  const int d = cid / nbs, bbase = (cid % nbs) * BS;
  const int j0 = rank * UPC;
  const int tid = threadIdx.x;
  load_w_rows(sW, whh + (size_t)d * 3 * H * H, H, HP, UPC, j0);
  // matvec role: columns c0 = tid and c1 = tid + NT of the H columns, all BS batch rows
  const int c0 = tid, c1 = tid + NT;
  const bool has0 = c0 < H, has1 = c1 < H;
  float* dst0 = has0 ? cluster.map_shared_rank(sR + (rank * BS) * UPC + c0 % UPC, c0 / UPC) : nullptr;
  float* dst1 = has1 ? cluster.map_shared_rank(sR + (rank * BS) * UPC + c1 % UPC, c1 / UPC) : nullptr;
  // gate role: one thread per (unit fu, batch row fb)
  const bool fin = tid < UPC * BS;
  const int fu = tid % UPC, fb = tid / UPC;
  const int j = j0 + fu, b = bbase + fb;
  const bool live = fin && b < B;
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
  cluster_arrive(); cluster_wait();   // every CTA of the cluster is running
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
      sD[(0 * UPC + fu) * BS + fb] = drp; sD[(1 * UPC + fu) * BS + fb] = dzp; sD[(2 * UPC + fu) * BS + fb] = dnr;
    }
    float n_r = 0.f, n_z = 0.f, n_n = 0.f, n_hn = 0.f, n_hp = 0.f, n_do = 0.f;
    if (live && s + 1 < T) fetch(s + 1, n_r, n_z, n_n, n_hn, n_hp, n_do);
    __syncthreads();
    // partial dh_{t-1}[b][c] = sum_{my rows} dgh[b][row] * W_hh[row][c]
    float a0[BS], a1[BS];
#pragma unroll
    for (int i = 0; i < BS; ++i) a0[i] = a1[i] = 0.f;
    if (has0) {
#pragma unroll 2
      for (int r = 0; r < rows; ++r) {
        const float w0 = sW[r * HP + c0];
        const float w1 = has1 ? sW[r * HP + c1] : 0.f;
#pragma unroll
        for (int q = 0; q < BS / 4; ++q) {
          const float4 dv = *reinterpret_cast<const float4*>(sD + r * BS + 4 * q);
          a0[4 * q + 0] = fmaf(w0, dv.x, a0[4 * q + 0]); a0[4 * q + 1] = fmaf(w0, dv.y, a0[4 * q + 1]);
          a0[4 * q + 2] = fmaf(w0, dv.z, a0[4 * q + 2]); a0[4 * q + 3] = fmaf(w0, dv.w, a0[4 * q + 3]);
          a1[4 * q + 0] = fmaf(w1, dv.x, a1[4 * q + 0]); a1[4 * q + 1] = fmaf(w1, dv.y, a1[4 * q + 1]);
          a1[4 * q + 2] = fmaf(w1, dv.z, a1[4 * q + 2]); a1[4 * q + 3] = fmaf(w1, dv.w, a1[4 * q + 3]);
        }
      }
    }
    if (s > 0) cluster_wait();          // (A) every CTA is done reading last step's partials
    if (has0) {
#pragma unroll
      for (int i = 0; i < BS; ++i) dst0[i * UPC] = a0[i];
    }
    if (has1) {
#pragma unroll
      for (int i = 0; i < BS; ++i) dst1[i * UPC] = a1[i];
    }
    cluster_arrive(); cluster_wait();   // (B) all partials of this step have landed
    if (fin) {
      float acc = 0.f;
#pragma unroll
      for (int r = 0; r < CS; ++r) acc += sR[(r * BS + fb) * UPC + fu];
      dh_carry = acc + dhz;
    }
    cluster_arrive_relaxed();           // (A) done reading the partials
    c_r = n_r; c_z = n_z; c_n = n_n; c_hn = n_hn; c_hp = n_hp; c_do = n_do;
  }
  cluster_wait();   // pairs with the last arrive; no CTA exits while a peer could still write into its shared memory
}

struct ClusterPlan { int cs, bs, nbs, KS, per, max_f, max_b; size_t smem_f, smem_b; };

typedef void (*FwdK)(const float*, const float*, const float*, float*, float*, int, int, int, int, int, int, int);
typedef void (*BwdK)(const float*, const float*, const float*, const float*, float*, float*, int, int, int, int, int);
FwdK fwd_kernel_of(int cs, int bs) {
  if (cs == 8) return bs == 8 ? gru_cluster_fwd_kernel<8, 8> : bs == 12 ? gru_cluster_fwd_kernel<8, 12> : gru_cluster_fwd_kernel<8, 16>;
  return bs == 8 ? gru_cluster_fwd_kernel<16, 8> : bs == 12 ? gru_cluster_fwd_kernel<16, 12> : gru_cluster_fwd_kernel<16, 16>;
}
BwdK bwd_kernel_of(int cs, int bs) {
  if (cs == 8) return bs == 8 ? gru_cluster_bwd_kernel<8, 8> : bs == 12 ? gru_cluster_bwd_kernel<8, 12> : gru_cluster_bwd_kernel<8, 16>;
  return bs == 8 ? gru_cluster_bwd_kernel<16, 8> : bs == 12 ? gru_cluster_bwd_kernel<16, 12> : gru_cluster_bwd_kernel<16, 16>;
}

void cluster_config(cudaLaunchConfig_t& cfg, cudaLaunchAttribute* attr, int cs, int grid, int nt, size_t smem, cudaStream_t s) {
  cfg = cudaLaunchConfig_t{};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(nt); cfg.dynamicSmemBytes = smem; cfg.stream = s;
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cs; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
}

// clusters of cs CTAs x nt threads x smem bytes the device keeps co-resident (0 = cannot launch)
int max_active_clusters(const void* fn, int cs, int nt, size_t smem) {
  if (cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess ||
      (cs > 8 && cudaFuncSetAttribute(fn, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess)) {
    cudaGetLastError();
    return 0;
  }
  cudaLaunchConfig_t cfg; cudaLaunchAttribute attr[1];
  cluster_config(cfg, attr, cs, cs, nt, smem, 0);
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, fn, &cfg) != cudaSuccess) { cudaGetLastError(); n = 0; }
  return n;
}

// Smallest cluster whose W slice fits in shared memory, then the batch-slice height that needs the fewest waves
// (cost model: waves x (rows of compute + ~10 rows' worth of barrier/push latency)).  Cached per (ndir, B, H).
bool plan_cluster(int ndir, int B, int H, ClusterPlan& best) {
  if (!g_gru_cluster || H % 4 != 0) return false;
  static int ck[4] = {0, 0, 0, -1};
  static ClusterPlan cached;
  static bool cached_ok = false;
  if (ck[0] == ndir && ck[1] == B && ck[2] == H && ck[3] == g_gru_bs) { best = cached; return cached_ok; }
  bool found = false;
  long best_cost = 0;
  for (int cs = 8; cs <= 16 && !found; cs *= 2) {
    if (H % cs) continue;
    const int UPC = H / cs;
    for (int bs = 8; bs <= 16; bs += 4) {
      const int nt = 32 * bs, tiles = UPC * (bs / 4);
      if (g_gru_bs && bs != g_gru_bs) continue;
      if (tiles > nt || UPC * bs > nt || H > 2 * nt) continue;
      ClusterPlan p;
      p.cs = cs; p.bs = bs; p.nbs = (B + bs - 1) / bs;
      p.KS = 8; while (p.KS > 1 && p.KS * tiles > nt) p.KS >>= 1;
      p.per = ((H >> 2) + p.KS - 1) / p.KS;
      const int half = p.KS >= 2 ? p.KS / 2 : 1;
      p.smem_f = sizeof(float) * ((size_t)3 * UPC * (H + 4) + bs * (H + 4) + (size_t)half * tiles * 12);
      p.smem_b = sizeof(float) * ((size_t)3 * UPC * (H + 4) + (size_t)cs * bs * UPC + (size_t)3 * UPC * bs);
      if (p.smem_f > 227 * 1024 || p.smem_b > 227 * 1024) continue;
      p.max_f = max_active_clusters((const void*)fwd_kernel_of(cs, bs), cs, nt, p.smem_f);
      p.max_b = max_active_clusters((const void*)bwd_kernel_of(cs, bs), cs, nt, p.smem_b);
      const int m = p.max_f < p.max_b ? p.max_f : p.max_b;
      if (m < 1) continue;
      const int need = ndir * p.nbs;
      const long cost = (long)((need + m - 1) / m) * (bs + 10);
      if (!found || cost < best_cost) { best = p; best_cost = cost; found = true; }
      if (bs >= B && !g_gru_bs) break;   // taller slices only add idle rows
    }
  }
  ck[0] = ndir; ck[1] = B; ck[2] = H; ck[3] = g_gru_bs; cached = best; cached_ok = found;
  return found;
}

template <class K, class... Args>
int launch_cluster(K kernel, int cs, int nt, int grid, size_t smem, cudaStream_t s, const char* what, Args... args) {
  cudaLaunchConfig_t cfg; cudaLaunchAttribute attr[1];
  cluster_config(cfg, attr, cs, grid, nt, smem, s);
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
  if (!plan_cluster(ndir, B, H, p)) return 0;
  return launch_cluster(fwd_kernel_of(p.cs, p.bs), p.cs, 32 * p.bs, ndir * p.nbs * p.cs, p.smem_f, s, "vca_gru_seq_fwd", gi, whh,
                        bhh, out, gates, ndir, p.nbs, T, B, H, p.KS, p.per);
}
int gru_cluster_bwd_try(const float* dout, const float* whh, const float* gates, const float* out, float* dgi, float* dgh,
                        int ndir, int T, int B, int H, cudaStream_t s) {
  ClusterPlan p;
  if (!plan_cluster(ndir, B, H, p)) return 0;
  return launch_cluster(bwd_kernel_of(p.cs, p.bs), p.cs, 32 * p.bs, ndir * p.nbs * p.cs, p.smem_b, s, "vca_gru_seq_bwd", dout, whh,
                        gates, out, dgi, dgh, ndir, p.nbs, T, B, H);
}

extern "C" {
// Introspection for tools / DESIGN.md: info[6] = {CTAs per cluster (0 = cluster path not applicable), batch rows per
// cluster, clusters a bidirectional layer needs, clusters the device keeps co-resident for the forward kernel, same
// for the backward kernel, K slices of the forward matvec}.
int vca_gru_cluster_query(int B, int H, int* info) {
  VCA_CHECK_ARG(info && B > 0 && H > 0);
  for (int i = 0; i < 6; ++i) info[i] = 0;
  ClusterPlan p;
  if (!plan_cluster(2, B, H, p)) return VCA_OK;
  info[0] = p.cs; info[1] = p.bs; info[2] = 2 * p.nbs; info[3] = p.max_f; info[4] = p.max_b; info[5] = p.KS;
  return VCA_OK;
}
}
