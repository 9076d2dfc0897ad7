// Persistent GRU recurrence (nn.GRU of visual_front.py:20,33-34), fp32, one cooperative launch per layer and pass.
//
// The T time steps are strictly sequential, so instead of 2 launches per step the whole sequence runs inside ONE
// kernel whose CTAs are all co-resident (cudaLaunchCooperativeKernel).  Each CTA owns 8 hidden units of one
// direction and keeps their 24 rows of W_hh (forward) / 8 columns of W_hh (backward) in shared memory for the whole
// sequence; per step it pulls the 32-row slab of h_{t-1} (written by its peers in the previous step) through L2,
// does its 24 x B dot products, applies the gates and publishes h_t; a release/acquire counter in global memory is
// the grid barrier between steps (bounded spin: a lost arrival traps instead of hanging the GPU).
#include "common.cuh"

// gru_cluster.cu: cluster / distributed-shared-memory kernels (1 = launched, 0 = not applicable, < 0 = error)
int gru_cluster_fwd_try(const float* gi, const float* whh, const float* bhh, float* out, float* gates, int ndir, int T, int B,
                        int H, cudaStream_t s);
int gru_cluster_bwd_try(const float* dout, const float* whh, const float* gates, const float* out, float* dgi, float* dgh,
                        int ndir, int T, int B, int H, cudaStream_t s);

namespace {

constexpr int UNITS = 8;       // hidden units (forward) / columns (backward) per CTA = warps per CTA
constexpr int BT = 32;         // batch rows per slab = lanes
constexpr unsigned SPIN_LIMIT = 1u << 26;

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + __expf(-x)); }

// monotonic counter barrier: barrier #k is complete when *ctr >= k * nblocks
__device__ __forceinline__ void grid_barrier(unsigned* ctr, unsigned target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(ctr, 1u);
    unsigned spins = 0;
    while (true) {
      unsigned v;
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
      if (v >= target) break;
      if (++spins > SPIN_LIMIT) __trap();
    }
    __threadfence();
  }
  __syncthreads();
}

// gi [ndir][T][B][3H] (incl. b_ih); whh [ndir][3H][H]; bhh [ndir][3H]; hbuf [2][ndir][B][H] (hbuf[0] = h_0 = 0);
// out [T][B][ndir*H]; gates [ndir][T][B][4H] = r, z, n, hn.
__global__ void __launch_bounds__(UNITS * BT, 1)
gru_seq_fwd_kernel(const float* __restrict__ gi, const float* __restrict__ whh, const float* __restrict__ bhh,
                   float* hbuf, float* __restrict__ out, float* __restrict__ gates, unsigned* bar, int ndir, int T, int B, int H) {
  extern __shared__ float sm[];
  float* sW = sm;                              // [3*UNITS][H]
  float* sH = sm + 3 * UNITS * H;              // [BT][H+4]
  const int ctas_per_dir = H / UNITS;
  const int d = blockIdx.x / ctas_per_dir, j0 = (blockIdx.x % ctas_per_dir) * UNITS;
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int j = j0 + w;
  const float* wd = whh + (long long)d * 3 * H * H;
  for (int i = threadIdx.x; i < 3 * UNITS * (H >> 2); i += blockDim.x) {
    const int r = i / (H >> 2), c4 = i - r * (H >> 2);
    const int g = r / UNITS, u = r - g * UNITS;                     // smem row r = gate g of unit u
    *reinterpret_cast<float4*>(sW + (u * 3 + g) * H + c4 * 4) =
        *reinterpret_cast<const float4*>(wd + (long long)(g * H + j0 + u) * H + c4 * 4);
  }
  const float br = bhh[d * 3 * H + j], bz = bhh[d * 3 * H + H + j], bn = bhh[d * 3 * H + 2 * H + j];
  const long long hsz = (long long)ndir * B * H;
  for (int s = 0; s < T; ++s) {
    const int t = d == 0 ? s : T - 1 - s;
    const float* hp = hbuf + (s & 1) * hsz + (long long)d * B * H;
    float* hn_ = hbuf + ((s + 1) & 1) * hsz + (long long)d * B * H;
    for (int b0 = 0; b0 < B; b0 += BT) {
      __syncthreads();
      for (int i = threadIdx.x; i < BT * (H >> 2); i += blockDim.x) {
        const int r = i / (H >> 2), c4 = i - r * (H >> 2);
        float4 v = make_float4(0, 0, 0, 0);
        if (b0 + r < B) v = __ldcg(reinterpret_cast<const float4*>(hp + (long long)(b0 + r) * H + c4 * 4));
        *reinterpret_cast<float4*>(sH + r * (H + 4) + c4 * 4) = v;
      }
      __syncthreads();
      float ar = 0.f, az = 0.f, an = 0.f;
      const float4* h4 = reinterpret_cast<const float4*>(sH + lane * (H + 4));
      const float4* r4 = reinterpret_cast<const float4*>(sW + (w * 3 + 0) * H);
      const float4* z4 = reinterpret_cast<const float4*>(sW + (w * 3 + 1) * H);
      const float4* n4 = reinterpret_cast<const float4*>(sW + (w * 3 + 2) * H);
#pragma unroll 4
      for (int k = 0; k < (H >> 2); ++k) {
        const float4 h = h4[k], a = r4[k], bq = z4[k], c = n4[k];
        ar = fmaf(a.x, h.x, ar); ar = fmaf(a.y, h.y, ar); ar = fmaf(a.z, h.z, ar); ar = fmaf(a.w, h.w, ar);
        az = fmaf(bq.x, h.x, az); az = fmaf(bq.y, h.y, az); az = fmaf(bq.z, h.z, az); az = fmaf(bq.w, h.w, az);
        an = fmaf(c.x, h.x, an); an = fmaf(c.y, h.y, an); an = fmaf(c.z, h.z, an); an = fmaf(c.w, h.w, an);
      }
      const int b = b0 + lane;
      if (b < B) {
        const float* gp = gi + (((long long)d * T + t) * B + b) * 3 * H;
        const float rr = sigmoidf_(gp[j] + ar + br);
        const float zz = sigmoidf_(gp[H + j] + az + bz);
        const float hn = an + bn;
        const float nn = tanhf(gp[2 * H + j] + rr * hn);
        const float hprev = sH[lane * (H + 4) + j];
        const float h = (1.f - zz) * nn + zz * hprev;
        hn_[(long long)b * H + j] = h;
        out[((long long)t * B + b) * (ndir * H) + d * H + j] = h;
        float* gs = gates + (((long long)d * T + t) * B + b) * 4 * H;
        gs[j] = rr; gs[H + j] = zz; gs[2 * H + j] = nn; gs[3 * H + j] = hn;
      }
    }
    grid_barrier(bar, (unsigned)(s + 1) * gridDim.x);
  }
}

// Backward through time.  dout [T][B][ndir*H]; dgi/dgh [ndir][T][B][3H]; scratch: dhc [2][ndir][B][H] (dhc[0] = 0),
// dghc [ndir][B][3H] (this step's dgh), dhz [ndir][B][H] (dh_total * z).
__global__ void __launch_bounds__(UNITS * BT, 1)
gru_seq_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ whh, const float* __restrict__ gates,
                   const float* __restrict__ out, float* __restrict__ dgi, float* __restrict__ dgh, float* dhc, float* dghc,
                   float* dhz, unsigned* bar, int ndir, int T, int B, int H) {
  extern __shared__ float sm[];
  float* sWt = sm;                             // [UNITS][3H]: sWt[kk][row] = W_hh[row][k0+kk]
  float* sG = sm + UNITS * 3 * H;              // [BT][H+4] slab of dgh
  const int ctas_per_dir = H / UNITS;
  const int d = blockIdx.x / ctas_per_dir, j0 = (blockIdx.x % ctas_per_dir) * UNITS;
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int j = j0 + w;
  const float* wd = whh + (long long)d * 3 * H * H;
  for (int i = threadIdx.x; i < 3 * H * UNITS; i += blockDim.x) {
    const int row = i / UNITS, kk = i - row * UNITS;
    sWt[kk * 3 * H + row] = wd[(long long)row * H + j0 + kk];
  }
  const long long hsz = (long long)ndir * B * H;
  unsigned nbar = 0;
  for (int s = 0; s < T; ++s) {
    const int t = d == 0 ? T - 1 - s : s;           // reverse of the forward order
    const int tp = d == 0 ? t - 1 : t + 1;
    const float* dc = dhc + (s & 1) * hsz + (long long)d * B * H;
    float* dn_ = dhc + ((s + 1) & 1) * hsz + (long long)d * B * H;
    // ---- phase A: gate gradients of my 8 units for every batch row
    for (int b = lane; b < B; b += BT) {
      const float hp = (tp >= 0 && tp < T) ? out[((long long)tp * B + b) * (ndir * H) + d * H + j] : 0.f;
      const float* gs = gates + (((long long)d * T + t) * B + b) * 4 * H;
      const float rr = gs[j], zz = gs[H + j], nn = gs[2 * H + j], hn = gs[3 * H + j];
      const float dh = dout[((long long)t * B + b) * (ndir * H) + d * H + j] + __ldcg(dc + (long long)b * H + j);
      const float dnp = dh * (1.f - zz) * (1.f - nn * nn);
      const float drp = dnp * hn * rr * (1.f - rr);
      const float dzp = dh * (hp - nn) * zz * (1.f - zz);
      const long long go = (((long long)d * T + t) * B + b) * 3 * H;
      dgi[go + j] = drp; dgi[go + H + j] = dzp; dgi[go + 2 * H + j] = dnp;
      dgh[go + j] = drp; dgh[go + H + j] = dzp; dgh[go + 2 * H + j] = dnp * rr;
      float* gc = dghc + ((long long)d * B + b) * 3 * H;
      gc[j] = drp; gc[H + j] = dzp; gc[2 * H + j] = dnp * rr;
      dhz[((long long)d * B + b) * H + j] = dh * zz;
    }
    grid_barrier(bar, (++nbar) * gridDim.x);
    // ---- phase B: dh_{t-1}[b][k] = dh*z + sum_row dgh[b][row] * W_hh[row][k] for my 8 columns k
    for (int b0 = 0; b0 < B; b0 += BT) {
      float acc = 0.f;
      for (int g = 0; g < 3; ++g) {
        __syncthreads();
        for (int i = threadIdx.x; i < BT * (H >> 2); i += blockDim.x) {
          const int r = i / (H >> 2), c4 = i - r * (H >> 2);
          float4 v = make_float4(0, 0, 0, 0);
          if (b0 + r < B) v = __ldcg(reinterpret_cast<const float4*>(dghc + ((long long)d * B + b0 + r) * 3 * H + g * H + c4 * 4));
          *reinterpret_cast<float4*>(sG + r * (H + 4) + c4 * 4) = v;
        }
        __syncthreads();
        const float4* g4 = reinterpret_cast<const float4*>(sG + lane * (H + 4));
        const float4* w4 = reinterpret_cast<const float4*>(sWt + w * 3 * H + g * H);
#pragma unroll 4
        for (int k = 0; k < (H >> 2); ++k) {
          const float4 a = w4[k], x = g4[k];
          acc = fmaf(a.x, x.x, acc); acc = fmaf(a.y, x.y, acc); acc = fmaf(a.z, x.z, acc); acc = fmaf(a.w, x.w, acc);
        }
      }
      const int b = b0 + lane;
      if (b < B) dn_[(long long)b * H + j] = acc + __ldcg(dhz + ((long long)d * B + b) * H + j);
    }
    grid_barrier(bar, (++nbar) * gridDim.x);
  }
}

int coop_ok(const void* fn, int grid, int block, size_t smem) {
  int dev = 0, coop = 0, per_sm = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
  if (!coop) return 0;
  if (cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, block, smem) != cudaSuccess) return 0;
  return per_sm * vca_num_sms() >= grid;
}

}  // namespace

extern "C" {

// Whole-sequence GRU recurrence, forward.  bar: device uint32 (zeroed here).  Returns VCA_ERR_UNSUPPORTED when the
// shape cannot run as one co-resident grid (the caller then uses the per-step kernels).
int vca_gru_seq_fwd(const float* gi, const float* whh, const float* bhh, float* hbuf, float* out, float* gates, unsigned* bar,
                    int ndir, int T, int B, int H, cudaStream_t s) {
  VCA_CHECK_ARG(gi && whh && bhh && hbuf && out && gates && bar && ndir > 0 && T > 0 && B > 0 && H > 0);
  if (const int r = gru_cluster_fwd_try(gi, whh, bhh, out, gates, ndir, T, B, H, s)) return r < 0 ? r : VCA_OK;
  if (H % UNITS || H % 4) { vca_set_error("vca_gru_seq_fwd: H must be a multiple of %d", UNITS); return VCA_ERR_UNSUPPORTED; }
  const int grid = ndir * (H / UNITS);
  const size_t smem = (size_t)(3 * UNITS * H + BT * (H + 4)) * sizeof(float);
  if (smem > 220 * 1024 || !coop_ok((const void*)gru_seq_fwd_kernel, grid, UNITS * BT, smem)) {
    vca_set_error("vca_gru_seq_fwd: grid of %d CTAs x %zu B smem cannot be co-resident", grid, smem);
    return VCA_ERR_UNSUPPORTED;
  }
  cudaMemsetAsync(bar, 0, sizeof(unsigned), s);
  cudaMemsetAsync(hbuf, 0, sizeof(float) * (size_t)ndir * B * H, s);   // h_0 = 0 (first half of the ping-pong)
  void* args[] = {&gi, &whh, &bhh, &hbuf, &out, &gates, &bar, &ndir, &T, &B, &H};
  if (cudaLaunchCooperativeKernel((const void*)gru_seq_fwd_kernel, dim3(grid), dim3(UNITS * BT), args, smem, s) != cudaSuccess) {
    vca_set_error("vca_gru_seq_fwd: cooperative launch failed: %s", cudaGetErrorString(cudaGetLastError()));
    return VCA_ERR_CUDA;
  }
  return VCA_OK;
}

int vca_gru_seq_bwd(const float* dout, const float* whh, const float* gates, const float* out, float* dgi, float* dgh,
                    float* dhc, float* dghc, float* dhz, unsigned* bar, int ndir, int T, int B, int H, cudaStream_t s) {
  VCA_CHECK_ARG(dout && whh && gates && out && dgi && dgh && dhc && dghc && dhz && bar && ndir > 0 && T > 0 && B > 0 && H > 0);
  if (const int r = gru_cluster_bwd_try(dout, whh, gates, out, dgi, dgh, ndir, T, B, H, s)) return r < 0 ? r : VCA_OK;
  if (H % UNITS || H % 4) { vca_set_error("vca_gru_seq_bwd: H must be a multiple of %d", UNITS); return VCA_ERR_UNSUPPORTED; }
  const int grid = ndir * (H / UNITS);
  const size_t smem = (size_t)(3 * UNITS * H + BT * (H + 4)) * sizeof(float);
  if (smem > 220 * 1024 || !coop_ok((const void*)gru_seq_bwd_kernel, grid, UNITS * BT, smem)) {
    vca_set_error("vca_gru_seq_bwd: grid of %d CTAs x %zu B smem cannot be co-resident", grid, smem);
    return VCA_ERR_UNSUPPORTED;
  }
  cudaMemsetAsync(bar, 0, sizeof(unsigned), s);
  cudaMemsetAsync(dhc, 0, sizeof(float) * (size_t)ndir * B * H, s);    // dh_T = 0
  void* args[] = {&dout, &whh, &gates, &out, &dgi, &dgh, &dhc, &dghc, &dhz, &bar, &ndir, &T, &B, &H};
  if (cudaLaunchCooperativeKernel((const void*)gru_seq_bwd_kernel, dim3(grid), dim3(UNITS * BT), args, smem, s) != cudaSuccess) {
    vca_set_error("vca_gru_seq_bwd: cooperative launch failed: %s", cudaGetErrorString(cudaGetLastError()));
    return VCA_ERR_CUDA;
  }
  return VCA_OK;
}

}  // extern "C"
