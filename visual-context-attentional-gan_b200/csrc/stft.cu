// Griffin-Lim mel/spectrogram -> waveform loop (src/data/audio_processing.py:51-68 over src/data/stft.py:70-129),
// n_fft = win = 640, hop = 160, periodic Hann, reflect-padded "center" frames, batched over clips.
//
// The reference runs every STFT / ISTFT as a dense 642x640 fp32 DFT convolution (29.8 GFLOP per clip for 60
// iterations) and recomputes the window envelope on the host 61 times.  Here one iteration is two kernels:
//   gl_frames_kernel : one warp per frame -- gather the 640 reflect-padded samples, window, 640-point real FFT, keep only
//                      the unit phasor, multiply by the target magnitude, inverse real FFT, window again, write the
//                      640-sample frame.  The phase never leaves the chip.  (mode 0: phases come from a given angle
//                      tensor = the reference's random initial phase.)
//   gl_ola_kernel    : overlap-add of the <= 4 frames covering each output sample, divided by the window
//                      sum-of-squares (audio_processing.py:7-48), i.e. the reference's inverse() tail.
// Algebra: inverse_basis = pinv(4 F)^T * w (stft.py:45-68) is exactly irfft-weights * w / 4 (rows of Im at DC and
// Nyquist are zero, so those imaginary parts are ignored), and the trailing * n_fft/hop = 4 cancels the 1/4.
//
// The FFT lives in REGISTERS.  The 640-point real transform is a 320-point complex one on z[n] = x[2n] + i x[2n+1];
// 320 = 10 x 32: lane n2 holds z[32 n1 + n2], n1 = 0..9.  Forward: a 10-point DFT per lane (2 x radix 5, constants
// only), the twiddle W_320^(n2 k1) (nine per-lane constants kept in registers for the whole kernel), then a 32-point
// decimation-in-frequency FFT ACROSS the lanes with shfl_xor (5 stages, per-lane stage twiddles in registers).  Lane l
// ends up with Z[k1 + 10 bitrev5(l)].  The real-FFT split needs Z[k] next to conj Z[320 - k]: that is slot 10 - k1 of
// lane l ^ 31 (one more shuffle); only slot 0 needs a general lane permutation.  The inverse mirrors it (decimation in
// time from the bit-reversed order back to natural order), so no reordering pass exists anywhere and shared memory
// only holds the read-only twiddle / window tables.  (The earlier Stockham version in shared memory was bound by
// the shared-memory pipe: 850 wavefronts per frame, a quarter of them bank conflicts.)
#include "common.cuh"

namespace {

constexpr int NFFT = 640, HOP = 160, NH = 320, NBIN = 321, WARPS = 8;
constexpr float PI2 = 6.283185307179586f;

struct cpx { float x, y; };
__device__ __forceinline__ cpx cmul(cpx a, cpx b) { return {a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x}; }
__device__ __forceinline__ cpx cadd(cpx a, cpx b) { return {a.x + b.x, a.y + b.y}; }
__device__ __forceinline__ cpx csub(cpx a, cpx b) { return {a.x - b.x, a.y - b.y}; }
__device__ __forceinline__ cpx cconj(cpx a) { return {a.x, -a.y}; }
template <bool INV> __device__ __forceinline__ cpx rot(cpx a) { return INV ? cpx{-a.y, a.x} : cpx{a.y, -a.x}; }  // * -+i
__device__ __forceinline__ cpx shfl_xor(cpx a, int m) {
  return {__shfl_xor_sync(0xffffffffu, a.x, m), __shfl_xor_sync(0xffffffffu, a.y, m)};
}
__device__ __forceinline__ cpx shfl_idx(cpx a, int src) {
  return {__shfl_sync(0xffffffffu, a.x, src), __shfl_sync(0xffffffffu, a.y, src)};
}

// 5-point DFT of (v0..v4) in place; INV conjugates the roots.
template <bool INV>
__device__ __forceinline__ void dft5(cpx& v0, cpx& v1, cpx& v2, cpx& v3, cpx& v4) {
  const float c1 = 0.30901699437494745f, c2 = -0.8090169943749475f, s1 = 0.9510565162951535f, s2 = 0.5877852522924731f;
  const cpx b1 = cadd(v1, v4), b2 = cadd(v2, v3), d1 = csub(v1, v4), d2 = csub(v2, v3);
  const cpx m1 = {v0.x + c1 * b1.x + c2 * b2.x, v0.y + c1 * b1.y + c2 * b2.y};
  const cpx m2 = {v0.x + c2 * b1.x + c1 * b2.x, v0.y + c2 * b1.y + c1 * b2.y};
  const cpx n1 = {s1 * d1.x + s2 * d2.x, s1 * d1.y + s2 * d2.y};
  const cpx n2 = {s2 * d1.x - s1 * d2.x, s2 * d1.y - s1 * d2.y};
  const cpx i1 = rot<INV>(n1), i2 = rot<INV>(n2);
  v0 = {v0.x + b1.x + b2.x, v0.y + b1.y + b2.y};
  v1 = cadd(m1, i1); v4 = csub(m1, i1);
  v2 = cadd(m2, i2); v3 = csub(m2, i2);
}

// 10-point DFT over the register index: r[k] <- sum_n r[n] W_10^(+-nk)  (even / odd halves, two radix-5 DFTs).
template <bool INV>
__device__ __forceinline__ void dft10(cpx (&r)[10]) {
  dft5<INV>(r[0], r[2], r[4], r[6], r[8]);          // E[k] now sits in r[2k]
  dft5<INV>(r[1], r[3], r[5], r[7], r[9]);          // O[k] in r[2k+1]
  const float wc[5] = {1.f, 0.8090169943749475f, 0.30901699437494745f, -0.30901699437494745f, -0.8090169943749475f};
  const float ws[5] = {0.f, 0.5877852522924731f, 0.9510565162951535f, 0.9510565162951535f, 0.5877852522924731f};
  cpx out[10];
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    const cpx w = {wc[k], INV ? ws[k] : -ws[k]};
    const cpx o = k == 0 ? r[1] : cmul(r[2 * k + 1], w);
    out[k] = cadd(r[2 * k], o);
    out[k + 5] = csub(r[2 * k], o);
  }
#pragma unroll
  for (int k = 0; k < 10; ++k) r[k] = out[k];
}

// Per-lane constants of the transform, fetched once per warp from the tables.
struct LaneConst {
  const cpx* wl;   // shared table: wl[k1 * 32] = W_320^(lane k1) for this lane (pointer already offset by the lane)
  const cpx* ws;   // shared table: ws[st * 32] = stage twiddle for span 16 >> st: W_(2h)^(lane & (h-1)) if bit h set, else 1
  int src0;        // lane holding Z[320 - k] of this lane's slot-0 bin k = 10 bitrev5(lane)
  int k2;          // bitrev5(lane)
};

// 32-point FFT across the lanes for all 10 register slots.  Forward: decimation in frequency, natural -> bit-reversed
// lane order.  Inverse: decimation in time, bit-reversed -> natural, conjugate twiddles.
template <int H, bool INV>
__device__ __forceinline__ void lane_stage(cpx (&r)[10], const cpx w, int lane) {
  const float sg = (lane & H) ? -1.f : 1.f;
  const cpx wv = INV ? cconj(w) : w;
#pragma unroll
  for (int s = 0; s < 10; ++s) {
    if (!INV) {
      const cpx p = shfl_xor(r[s], H);
      const cpx t = {fmaf(sg, r[s].x, p.x), fmaf(sg, r[s].y, p.y)};     // bit clear: mine + partner; set: partner - mine
      r[s] = H == 1 ? t : cmul(t, wv);
    } else {
      const cpx wm = H == 1 ? r[s] : cmul(r[s], wv);
      const cpx p = shfl_xor(wm, H);
      r[s] = {fmaf(sg, wm.x, p.x), fmaf(sg, wm.y, p.y)};
    }
  }
}

template <bool INV>
__device__ __forceinline__ void fft320(cpx (&r)[10], const LaneConst& c, int lane) {
  if (!INV) {
    dft10<false>(r);
#pragma unroll
    for (int k = 1; k < 10; ++k) r[k] = cmul(r[k], c.wl[k * 32]);
    lane_stage<16, false>(r, c.ws[0 * 32], lane);
    lane_stage<8, false>(r, c.ws[1 * 32], lane);
    lane_stage<4, false>(r, c.ws[2 * 32], lane);
    lane_stage<2, false>(r, c.ws[3 * 32], lane);
    lane_stage<1, false>(r, cpx{1.f, 0.f}, lane);
  } else {
    lane_stage<1, true>(r, cpx{1.f, 0.f}, lane);
    lane_stage<2, true>(r, c.ws[3 * 32], lane);
    lane_stage<4, true>(r, c.ws[2 * 32], lane);
    lane_stage<8, true>(r, c.ws[1 * 32], lane);
    lane_stage<16, true>(r, c.ws[0 * 32], lane);
#pragma unroll
    for (int k = 1; k < 10; ++k) r[k] = cmul(r[k], cconj(c.wl[k * 32]));
    dft10<true>(r);
  }
}

// Per-device tables built once by gl_tables_kernel, all laid out so that a warp reads them with unit stride:
//   wl  [10][32] cpx : W_320^(lane k1)
//   tw2p[10][32] cpx : W_640^k for this lane's bin k = k1 + 10 bitrev5(lane)
//   ws  [4][32]  cpx : stage twiddles for spans 16, 8, 4, 2 (1 on the lanes whose bit is clear)
//   win [640]        : periodic Hann
constexpr int TAB_WL = 0, TAB_TW2 = 2 * NH, TAB_WS = 4 * NH, TAB_WIN = 4 * NH + 256, TAB_FLOATS = TAB_WIN + NFFT;
__device__ __align__(16) float g_tab[TAB_FLOATS];

__device__ __forceinline__ int bitrev5(int l) { return (int)(__brev((unsigned)l) >> 27); }

__global__ void gl_tables_kernel() {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < NFFT; i += gridDim.x * blockDim.x) {
    float s, c;
    if (i < NH) {
      const int k1 = i >> 5, lane = i & 31;
      sincosf(-PI2 * (float)(lane * k1) / (float)NH, &s, &c);
      g_tab[TAB_WL + 2 * i] = c; g_tab[TAB_WL + 2 * i + 1] = s;
      sincosf(-PI2 * (float)(k1 + 10 * bitrev5(lane)) / (float)NFFT, &s, &c);
      g_tab[TAB_TW2 + 2 * i] = c; g_tab[TAB_TW2 + 2 * i + 1] = s;
    }
    if (i < 128) {
      const int st = i >> 5, lane = i & 31, h = 16 >> st;
      c = 1.f; s = 0.f;
      if (lane & h) sincosf(-PI2 * (float)(lane & (h - 1)) / (float)(2 * h), &s, &c);
      g_tab[TAB_WS + 2 * i] = c; g_tab[TAB_WS + 2 * i + 1] = s;
    }
    g_tab[TAB_WIN + i] = 0.5f - 0.5f * cosf(PI2 * (float)i / (float)NFFT);
  }
}

// mag / angle tensors in "lane-major" bin order: out[row][k1 * 32 + lane] = in[row][k1 + 10 bitrev5(lane)], bin 320 stays
// last.  Done once per Griffin-Lim call so that the 60 iterations read the magnitudes with unit stride.
__global__ void gl_permute_bins_kernel(const float* __restrict__ in, float* __restrict__ out, long long rows) {
  const long long total = rows * NBIN;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / NBIN;
    const int j = (int)(i - row * NBIN);
    const int k = j == NH ? NH : (j >> 5) + 10 * bitrev5(j & 31);
    out[i] = in[row * NBIN + k];
  }
}

__device__ __forceinline__ int reflect_idx(int j, int L) {   // F.pad(mode='reflect')
  if (j < 0) j = -j;
  if (j >= L) j = 2 * (L - 1) - j;
  return j;
}

// window sum-of-squares envelope at padded position n = m + 320 (audio_processing.py:7-48 with T frames): the <= 4 frames covering it
__device__ __forceinline__ float gl_wss(int n, int T, const float* __restrict__ win) {
  const int t_hi = min(n / HOP, T - 1);
  const int t_lo = n < NFFT - HOP ? 0 : (n - (NFFT - HOP)) / HOP;     // ceil((n - 639) / 160)
  float wss = 0.f;
  for (int t = t_lo; t <= t_hi; ++t) { const float w = win[n - t * HOP]; wss = fmaf(w, w, wss); }
  return wss;
}

// v / envelope at signal position j -- the rare path (frames within three hops of either end): kept out of line so that it
// does not cost the common path registers
__device__ __noinline__ float gl_env_div(float v, int j, int T, const float* __restrict__ win) {
  const float e = gl_wss(j + NFFT / 2, T, win);
  return e > 1.1754944e-38f ? v / e : v;
}

// The 640 reflect-padded, windowed samples of frame t as z[n] = x[2n] + i x[2n+1], n = 32 n1 + lane.
// ENV: sb holds UN-normalised overlap-add sums (gl_iter_kernel's output): divide by the window envelope on the way in --
// 1.5 for every sample at least 480 away from both ends (periodic Hann, hop = N/4), the exact <= 4-term sum elsewhere.
template <bool ENV>
__device__ __forceinline__ void gl_load_frame(const float* __restrict__ sb, int t, int T, int L, const float2* __restrict__ win2,
                                              const float* __restrict__ win, int lane, cpx (&r)[10]) {
  const int base = t * HOP - NFFT / 2;
  const bool interior = base >= 0 && base + NFFT <= L;            // no reflection: aligned float2 loads
  const bool deep = base >= NFFT - HOP && base + NFFT <= L - (NFFT - HOP);
#pragma unroll
  for (int n1 = 0; n1 < 10; ++n1) {
    const int n = 32 * n1 + lane;
    const float2 w = win2[n];
    float2 x;
    if (interior) {
      x = *reinterpret_cast<const float2*>(sb + base + 2 * n);
      if (ENV) {
        if (deep) { x.x *= (2.f / 3.f); x.y *= (2.f / 3.f); }
        else { x.x = gl_env_div(x.x, base + 2 * n, T, win); x.y = gl_env_div(x.y, base + 2 * n + 1, T, win); }
      }
    } else {
      const int j0 = reflect_idx(base + 2 * n, L), j1 = reflect_idx(base + 2 * n + 1, L);
      x = make_float2(sb[j0], sb[j1]);
      if (ENV) { x.x = gl_env_div(x.x, j0, T, win); x.y = gl_env_div(x.y, j1, T, win); }
    }
    r[n1] = {x.x * w.x, x.y * w.y};
  }
}

// rFFT of the frame in r, then Y[k] = mag[k] * X[k] / |X[k]| (the phase of the signal, the target magnitude).
// spec_out (optional): X itself as (re, im) at spec_out[(fo * 321 + k) * 2].
__device__ __forceinline__ void gl_analyze(cpx (&r)[10], const LaneConst& c, int lane, const cpx* __restrict__ tw2p,
                                           const float* __restrict__ mg, bool lane_major, float* __restrict__ spec_out, long long fo,
                                           cpx (&y)[10], cpx& y320) {
  fft320<false>(r, c, lane);
  // X[k] = E[k] + W640^k O[k];  E = (Z[k] + conj Z[N-k])/2, O = (Z[k] - conj Z[N-k])/(2i).  Then Y[k] = mag * X/|X|.
#pragma unroll
  for (int k1 = 0; k1 < 10; ++k1) {
    const cpx zn = cconj(k1 == 0 ? shfl_idx(r[0], c.src0) : shfl_xor(r[10 - k1], 31));
    const cpx zk = r[k1];
    const cpx e = {0.5f * (zk.x + zn.x), 0.5f * (zk.y + zn.y)};
    const cpx d = csub(zk, zn);
    const cpx o = {0.5f * d.y, -0.5f * d.x};          // d / (2i)
    const int k = k1 + 10 * c.k2;
    const cpx X = cadd(e, cmul(tw2p[k1 * 32 + lane], o));
    if (spec_out) { spec_out[(fo * NBIN + k) * 2] = X.x; spec_out[(fo * NBIN + k) * 2 + 1] = X.y; }
    const float r2 = X.x * X.x + X.y * X.y;
    const float m = __ldg(mg + (lane_major ? k1 * 32 + lane : k));
    const float sc = m * rsqrtf(r2);                                  // mag / |X|
    y[k1] = r2 > 0.f ? cpx{sc * X.x, sc * X.y} : cpx{m, 0.f};         // atan2(0,0) = 0
    if (k1 == 0) {                                                    // Nyquist bin from the same pair (k = 0, lane 0)
      const cpx Xn = csub(e, o);
      if (spec_out && lane == 0) { spec_out[(fo * NBIN + NH) * 2] = Xn.x; spec_out[(fo * NBIN + NH) * 2 + 1] = Xn.y; }
      const float q2 = Xn.x * Xn.x + Xn.y * Xn.y;
      const float mn = __ldg(mg + NH);
      const float sn = mn * rsqrtf(q2);
      y320 = q2 > 0.f ? cpx{sn * Xn.x, sn * Xn.y} : cpx{mn, 0.f};
    }
  }
}

// inverse real FFT of the Hermitian spectrum Y[0..320] (imaginary parts of DC / Nyquist ignored, as in the reference's
// basis), windowed and scaled: z[n1] = the frame's samples (2n, 2n+1), n = 32 n1 + lane
//      Z[k] = E[k] + i O[k],  E = (Y[k] + conj Y[N-k])/2,  O = (Y[k] - conj Y[N-k])/2 * W640^{-k}
__device__ __forceinline__ void gl_synthesize(cpx (&y)[10], cpx y320, const LaneConst& c, int lane, const cpx* __restrict__ tw2p,
                                              const float2* __restrict__ win2, cpx (&z)[10]) {
  if (lane == 0) y[0].y = 0.f;
  y320.y = 0.f;
#pragma unroll
  for (int k1 = 0; k1 < 10; ++k1) {
    cpx yn = cconj(k1 == 0 ? shfl_idx(y[0], c.src0) : shfl_xor(y[10 - k1], 31));
    if (k1 == 0 && lane == 0) yn = y320;                 // k = 0 pairs with Y[320] (real)
    const cpx yk = y[k1];
    const cpx e = {0.5f * (yk.x + yn.x), 0.5f * (yk.y + yn.y)};
    const cpx d = {0.5f * (yk.x - yn.x), 0.5f * (yk.y - yn.y)};
    const cpx o = cmul(d, cconj(tw2p[k1 * 32 + lane]));
    z[k1] = {e.x - o.y, e.y + o.x};                      // e + i o
  }
  fft320<true>(z, c, lane);
  const float sc = 1.f / (float)NH;
#pragma unroll
  for (int n1 = 0; n1 < 10; ++n1) {
    const float2 w = win2[32 * n1 + lane];
    z[n1] = {z[n1].x * sc * w.x, z[n1].y * sc * w.y};
  }
}

// mode 0: phases from angles_t [B][T][321] (radians);  mode 1: phases from the STFT of sig [B][L].
// mag_t [B][T][321]; frames [B][T][640].  Optional spec_out [B][T][321][2] = STFT (re, im) of sig (mode 1 only).
__global__ void __launch_bounds__(WARPS * 32, 3) gl_frames_kernel(int mode, const float* __restrict__ sig,
                                                               const float* __restrict__ angles_t,
                                                               const float* __restrict__ mag_t, float* __restrict__ frames,
                                                               float* __restrict__ spec_out, int B, int T, int L, int fpw) {
  __shared__ __align__(16) float tab[TAB_FLOATS];
  for (int i = threadIdx.x; i < TAB_FLOATS / 4; i += blockDim.x)
    reinterpret_cast<float4*>(tab)[i] = reinterpret_cast<const float4*>(g_tab)[i];
  __syncthreads();
  const cpx* wl_t = reinterpret_cast<const cpx*>(tab + TAB_WL);
  const cpx* tw2p = reinterpret_cast<const cpx*>(tab + TAB_TW2);
  const cpx* ws_t = reinterpret_cast<const cpx*>(tab + TAB_WS);
  const float2* win2 = reinterpret_cast<const float2*>(tab + TAB_WIN);   // pairs (2n, 2n+1)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.y;
  const bool lane_major = (mode & 2) != 0;   // mag_t bins already in k1 * 32 + lane order
  mode &= 1;

  LaneConst c;
  c.wl = wl_t + lane;
  c.ws = ws_t + lane;
  c.k2 = bitrev5(lane);
  c.src0 = bitrev5((32 - c.k2) & 31);
  const float* sb = sig + (long long)b * L;

#pragma unroll 1
  for (int f = 0; f < fpw; ++f) {                                     // fpw frames per warp amortise the table load
    const int t = (blockIdx.x * fpw + f) * WARPS + warp;
    if (t >= T) return;                                               // warp-uniform
    const long long fo = ((long long)b * T + t);
    const float* mg = mag_t + fo * NBIN;
    cpx y[10];       // Hermitian half-spectrum Y[k], k = k1 + 10 k2 (k1 = register slot)
    cpx y320;        // Y[320]; meaningful on lane 0 only

    if (mode == 1) {
      cpx r[10];
      gl_load_frame<false>(sb, t, T, L, win2, tab + TAB_WIN, lane, r);
      gl_analyze(r, c, lane, tw2p, mg, lane_major, spec_out, fo, y, y320);
    } else {
      const float* an = angles_t + fo * NBIN;
#pragma unroll
      for (int k1 = 0; k1 < 10; ++k1) {
        const int k = k1 + 10 * c.k2;
        float s, co;
        sincosf(__ldg(an + k), &s, &co);
        const float m = __ldg(mg + k);
        y[k1] = {m * co, m * s};
      }
      float s, co;
      sincosf(__ldg(an + NH), &s, &co);
      const float mn = __ldg(mg + NH);
      y320 = {mn * co, mn * s};
    }
    cpx z[10];
    gl_synthesize(y, y320, c, lane, tw2p, win2, z);
    float* fr = frames + fo * NFFT;
#pragma unroll
    for (int n1 = 0; n1 < 10; ++n1) *reinterpret_cast<float2*>(fr + 2 * (32 * n1 + lane)) = make_float2(z[n1].x, z[n1].y);
  }
}

// One WHOLE Griffin-Lim iteration in one kernel: STFT of the current signal -> unit phasor x target magnitude -> ISTFT frame
// -> overlap-add INSIDE the CTA.  A CTA owns FPC = 8 * fpw consecutive frames of one clip; its warps add their frames into a
// shared accumulator of (FPC + 3) hops in four conflict-free phases (frames 4 apart do not overlap), then the CTA writes
// the hops only it touches with plain stores and adds the 3 + 3 boundary hops it shares with its neighbours with atomics
// (two addends: order-independent).  The 49 MB of frames per iteration never exist; the accumulator holds UN-normalised
// sums, the window-envelope division happens when the next iteration (or gl_normalize_kernel) reads them.
// in_norm: sig_in is already normalised (the first iteration, fed by gl_ola_kernel).  acc_out must be zero on entry.
template <bool ENV>
__global__ void __launch_bounds__(WARPS * 32, 2) gl_iter_kernel(const float* __restrict__ sig_in, const float* __restrict__ mag_p,
                                                              float* __restrict__ acc_out, int B, int T, int L, int fpw) {
  __shared__ __align__(16) float tab[TAB_FLOATS];
  extern __shared__ __align__(16) float s_acc[];                    // (8 * fpw + 3) * 160
  const int FPC = WARPS * fpw, NACC = (FPC + 3) * HOP;
  for (int i = threadIdx.x; i < TAB_FLOATS / 4; i += blockDim.x)
    reinterpret_cast<float4*>(tab)[i] = reinterpret_cast<const float4*>(g_tab)[i];
  for (int i = threadIdx.x; i < NACC; i += blockDim.x) s_acc[i] = 0.f;
  __syncthreads();
  const cpx* wl_t = reinterpret_cast<const cpx*>(tab + TAB_WL);
  const cpx* tw2p = reinterpret_cast<const cpx*>(tab + TAB_TW2);
  const cpx* ws_t = reinterpret_cast<const cpx*>(tab + TAB_WS);
  const float2* win2 = reinterpret_cast<const float2*>(tab + TAB_WIN);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.y;
  const int t0 = blockIdx.x * FPC;
  LaneConst c;
  c.wl = wl_t + lane;
  c.ws = ws_t + lane;
  c.k2 = bitrev5(lane);
  c.src0 = bitrev5((32 - c.k2) & 31);
  const float* sb = sig_in + (long long)b * L;

#pragma unroll 1
  for (int f = 0; f < fpw; ++f) {
    const int tl = f * WARPS + warp, t = t0 + tl;
    const bool live = t < T;                                          // warp-uniform
    cpx z[10];
    if (live) {
      cpx r[10], y[10], y320;
      gl_load_frame<ENV>(sb, t, T, L, win2, tab + TAB_WIN, lane, r);
      gl_analyze(r, c, lane, tw2p, mag_p + ((long long)b * T + t) * NBIN, true, nullptr, 0, y, y320);
      gl_synthesize(y, y320, c, lane, tw2p, win2, z);
    }
    // frames tl, tl + 4 of this pass do not overlap (4 hops apart): four phases, no conflicts, a fixed summation order
#pragma unroll 1
    for (int ph = 0; ph < 4; ++ph) {
      if (live && (warp & 3) == ph) {
        float2* dst = reinterpret_cast<float2*>(s_acc + tl * HOP);
#pragma unroll
        for (int n1 = 0; n1 < 10; ++n1) { float2 a = dst[32 * n1 + lane]; a.x += z[n1].x; a.y += z[n1].y; dst[32 * n1 + lane] = a; }
      }
      __syncthreads();
    }
  }
  // s_acc[i] = sum over this CTA's frames at padded position n = 160 t0 + i; output sample m = n - 320
  const int nf = min(FPC, T - t0);
  float* out = acc_out + (long long)b * L;
  for (int i = threadIdx.x; i < (nf + 3) * HOP; i += blockDim.x) {
    const int m = t0 * HOP + i - NFFT / 2;
    if (m < 0 || m >= L) continue;
    const bool shared_lo = i < NFFT - HOP && t0 > 0;                 // also covered by the previous CTA's last three frames
    const bool shared_hi = i >= nf * HOP && t0 + nf < T;             // ... by the next CTA's first three
    if (shared_lo || shared_hi) atomicAdd(out + m, s_acc[i]); else out[m] = s_acc[i];
  }
}

// sig[b][m] = acc[b][m] / wss[m + 320]: the normalisation gl_ola_kernel does, for the un-normalised sums of gl_iter_kernel
__global__ void __launch_bounds__(256) gl_normalize_kernel(const float* __restrict__ acc, float* __restrict__ sig_out, int T, int L) {
  const int m = blockIdx.x * 256 + threadIdx.x, b = blockIdx.y;
  if (m >= L) return;
  const float wss = gl_wss(m + NFFT / 2, T, g_tab + TAB_WIN);
  const float a = acc[(size_t)b * L + m];
  sig_out[(size_t)b * L + m] = wss > 1.1754944e-38f ? a / wss : a;
}

// sig_out[b][m] = (sum_t frames[b][t][m + 320 - 160 t]) / wss[m + 320]   for m in [0, L), L = 160 (T - 1)
__global__ void __launch_bounds__(256) gl_ola_kernel(const float* __restrict__ frames, float* __restrict__ sig_out, int T,
                                                     int L) {
  const int m = blockIdx.x * 256 + threadIdx.x, b = blockIdx.y;
  if (m >= L) return;
  const int n = m + NFFT / 2;
  const int t_hi = min(n / HOP, T - 1);
  const int t_lo = n < NFFT - HOP ? 0 : (n - (NFFT - HOP)) / HOP;     // ceil((n - 639) / 160)
  const float* win = g_tab + TAB_WIN;
  const float* fb = frames + (size_t)b * T * NFFT;
  float acc = 0.f, wss = 0.f;
  for (int t = t_lo; t <= t_hi; ++t) {
    const int r = n - t * HOP;                                        // 0 <= r < 640 by construction
    const float w = __ldg(win + r);
    wss = fmaf(w, w, wss);
    acc += fb[(size_t)t * NFFT + r];
  }
  sig_out[(size_t)b * L + m] = wss > 1.1754944e-38f ? acc / wss : acc;
}

}  // namespace

int g_gl_fpw = 2;   // "gl_fpw" in vca_set_option: frames per warp of gl_frames_kernel

// Build the per-device tables, stream-ordered in front of their first use on this device.
static int gl_ensure_tables(cudaStream_t s) {
  static bool ready[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64 || !ready[dev]) {
    gl_tables_kernel<<<3, 256, 0, s>>>();
    VCA_LAUNCH_CHECK();
    if (dev >= 0 && dev < 64) ready[dev] = true;
  }
  return VCA_OK;
}

extern "C" {

// One half-iteration of Griffin-Lim: frames[b][t][:] = window * irfft( mag[b][t][:] * unit_phase ), where the phase is
// exp(i*angles_t) (mode 0) or that of rfft(window * reflect-padded sig frame) (mode 1).  spec_out (optional,
// mode 1): the STFT itself as (re, im).  All tensors fp32; mag_t / angles_t are frame-major [B][T][321].
int vca_gl_frames(int mode, const float* sig, const float* angles_t, const float* mag_t, float* frames, float* spec_out, int B,
                  int T, int L, cudaStream_t s) {
  VCA_CHECK_ARG(mag_t && frames && B > 0 && T > 1 && L == HOP * (T - 1) && mode >= 0 && mode <= 3 && mode != 2 && ((mode & 1) == 0 ? angles_t != nullptr : sig != nullptr));
  VCA_CHECK_ARG(B <= 65535);
  if (int e = gl_ensure_tables(s)) return e;
  const int fpw = g_gl_fpw;
  dim3 grid((T + WARPS * fpw - 1) / (WARPS * fpw), B);
  gl_frames_kernel<<<grid, WARPS * 32, 0, s>>>(mode, sig, angles_t, mag_t, frames, spec_out, B, T, L, fpw);
  VCA_LAUNCH_CHECK();
  return VCA_OK;
}
// One whole Griffin-Lim iteration (STFT -> phase x magnitude -> ISTFT -> overlap-add) in one kernel.  sig_in [B][L]: the
// current signal, normalised (in_norm != 0: the output of vca_gl_ola / vca_gl_normalize) or the UN-normalised sums a previous
// call left in its acc_out; mag_p [B][T][321] in the lane-major bin order of vca_gl_permute_bins; acc_out [B][L]: receives
// the un-normalised overlap-add sums and must be ZERO on entry (boundary hops of neighbouring CTAs are added atomically).
int vca_gl_iter(const float* sig_in, int in_norm, const float* mag_p, float* acc_out, int B, int T, int L, cudaStream_t s) {
  VCA_CHECK_ARG(sig_in && mag_p && acc_out && sig_in != acc_out && B > 0 && T > 1 && L == HOP * (T - 1));
  VCA_CHECK_ARG(B <= 65535);
  if (int e = gl_ensure_tables(s)) return e;
  const int fpw = g_gl_fpw;
  const int FPC = WARPS * fpw;
  dim3 grid((T + FPC - 1) / FPC, B);
  const size_t smem = (size_t)(FPC + 3) * HOP * sizeof(float);
  if (in_norm) gl_iter_kernel<false><<<grid, WARPS * 32, smem, s>>>(sig_in, mag_p, acc_out, B, T, L, fpw);
  else gl_iter_kernel<true><<<grid, WARPS * 32, smem, s>>>(sig_in, mag_p, acc_out, B, T, L, fpw);
  VCA_LAUNCH_CHECK();
  return VCA_OK;
}
// sig_out = acc / window envelope (the tail of STFT.inverse, stft.py:110-127) for the sums vca_gl_iter leaves
int vca_gl_normalize(const float* acc, float* sig_out, int B, int T, int L, cudaStream_t s) {
  VCA_CHECK_ARG(acc && sig_out && B > 0 && T > 1 && L == HOP * (T - 1));
  VCA_CHECK_ARG(B <= 65535);
  if (int e = gl_ensure_tables(s)) return e;
  gl_normalize_kernel<<<dim3((L + 255) / 256, B), 256, 0, s>>>(acc, sig_out, T, L);
  VCA_LAUNCH_CHECK();
  return VCA_OK;
}
// Reorder the 321 bins of every row into the lane-major order gl_frames reads with unit stride (mode 3).
int vca_gl_permute_bins(const float* in, float* out, long long rows, cudaStream_t s) {
  VCA_CHECK_ARG(in && out && in != out && rows > 0);
  gl_permute_bins_kernel<<<vca_grid_1d(rows * NBIN, 256), 256, 0, s>>>(in, out, rows);
  VCA_LAUNCH_CHECK();
  return VCA_OK;
}
// Overlap-add + window-envelope normalisation + trim (the tail of STFT.inverse, stft.py:110-127).
int vca_gl_ola(const float* frames, float* sig_out, int B, int T, int L, cudaStream_t s) {
  VCA_CHECK_ARG(frames && sig_out && B > 0 && T > 1 && L == HOP * (T - 1));
  VCA_CHECK_ARG(B <= 65535);
  if (int e = gl_ensure_tables(s)) return e;
  gl_ola_kernel<<<dim3((L + 255) / 256, B), 256, 0, s>>>(frames, sig_out, T, L);
  VCA_LAUNCH_CHECK();
  return VCA_OK;
}

}  // extern "C"
