// Griffin-Lim mel/spectrogram -> waveform loop (src/data/audio_processing.py:51-68 over src/data/stft.py:70-129),
// n_fft = win = 640, hop = 160, periodic Hann, reflect-padded "center" frames, batched over clips.
//
// The reference runs every STFT / ISTFT as a dense 642x640 fp32 DFT convolution (29.8 GFLOP per clip for 60
// iterations) and recomputes the window envelope on the host 61 times.  Here one iteration is two HBM-bound kernels:
//   gl_frames_kernel : one warp per frame -- gather the 640 reflect-padded samples, window, 640-point real FFT
//                      (320-point complex Stockham FFT, radix 4*4*4*5, in shared memory) , keep only the unit phasor,
//                      multiply by the target magnitude, inverse real FFT, window again, write the 640-sample frame.
//                      The phase never leaves the chip.  (mode 0: phases come from a given angle tensor = the
//                      reference's random initial phase.)
//   gl_ola_kernel    : overlap-add of the <= 4 frames covering each output sample, divided by the window
//                      sum-of-squares (audio_processing.py:7-48), i.e. the reference's inverse() tail.
// Algebra: inverse_basis = pinv(4 F)^T * w (stft.py:45-68) is exactly irfft-weights * w / 4 (rows of Im at DC and
// Nyquist are zero, so those imaginary parts are ignored), and the trailing * n_fft/hop = 4 cancels the 1/4.
#include "common.cuh"

namespace {

constexpr int NFFT = 640, HOP = 160, NH = 320, NBIN = 321, WARPS = 8;
constexpr float PI2 = 6.283185307179586f;

struct cpx { float x, y; };
__device__ __forceinline__ cpx cmul(cpx a, cpx b) { return {a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x}; }
__device__ __forceinline__ cpx cadd(cpx a, cpx b) { return {a.x + b.x, a.y + b.y}; }
__device__ __forceinline__ cpx csub(cpx a, cpx b) { return {a.x - b.x, a.y - b.y}; }
__device__ __forceinline__ cpx cconj(cpx a) { return {a.x, -a.y}; }
// multiply by -i (forward) or +i (inverse)
__device__ __forceinline__ cpx rot(cpx a, bool inv) { return inv ? cpx{-a.y, a.x} : cpx{a.y, -a.x}; }

// One Stockham stage of the 320-point FFT done by one warp: radix R butterflies with input stride NH/R, output
// blocks of NS*R.  NS, R and the direction are compile-time, so the index arithmetic is shifts and constants.
// tw[m] = exp(-2 pi i m / 320); the twiddle of input t of butterfly column k is W_{NS*R}^{k t} = tw[k t NH/(NS R)]
// (k t < NS R, so the index never wraps).
template <int NS, int R, bool INV>
__device__ __forceinline__ void fft_stage(const cpx* __restrict__ in, cpx* __restrict__ out, const cpx* __restrict__ tw,
                                          int lane) {
  constexpr int nb = NH / R, tstep = NH / (NS * R);
#pragma unroll
  for (int j = lane; j < nb; j += 32) {
    const int k = j % NS;                        // NS is a power of two
    const int j0 = (j / NS) * (NS * R) + k;
    cpx v[R];
#pragma unroll
    for (int t = 0; t < R; ++t) {
      cpx x = in[j + t * nb];
      if (NS > 1 && t > 0) {                     // k == 0 multiplies by tw[0] = 1: cheaper than a divergent branch
        cpx w = tw[k * t * tstep];
        if (INV) w.y = -w.y;
        x = cmul(x, w);
      }
      v[t] = x;
    }
    if (R == 4) {
      const cpx t0 = cadd(v[0], v[2]), t1 = csub(v[0], v[2]), t2 = cadd(v[1], v[3]), t3 = rot(csub(v[1], v[3]), INV);
      const cpx o0 = cadd(t0, t2), o1 = cadd(t1, t3), o2 = csub(t0, t2), o3 = csub(t1, t3);
      if (NS == 1) {                             // 4 consecutive outputs: two 16-byte stores (buffers are 16-byte aligned)
        float4* o = reinterpret_cast<float4*>(out + j0);
        o[0] = make_float4(o0.x, o0.y, o1.x, o1.y);
        o[1] = make_float4(o2.x, o2.y, o3.x, o3.y);
      } else {
        out[j0] = o0; out[j0 + NS] = o1; out[j0 + 2 * NS] = o2; out[j0 + 3 * NS] = o3;
      }
    } else {
      const float c1 = 0.30901699437494745f, c2 = -0.8090169943749475f, s1 = 0.9510565162951535f, s2 = 0.5877852522924731f;
      const cpx b1 = cadd(v[1], v[R - 1]), b2 = cadd(v[2], v[R - 2]), d1 = csub(v[1], v[R - 1]), d2 = csub(v[2], v[R - 2]);
      const cpx m1 = {v[0].x + c1 * b1.x + c2 * b2.x, v[0].y + c1 * b1.y + c2 * b2.y};
      const cpx m2 = {v[0].x + c2 * b1.x + c1 * b2.x, v[0].y + c2 * b1.y + c1 * b2.y};
      const cpx n1 = {s1 * d1.x + s2 * d2.x, s1 * d1.y + s2 * d2.y};
      const cpx n2 = {s2 * d1.x - s1 * d2.x, s2 * d1.y - s1 * d2.y};
      const cpx in1 = rot(n1, INV), in2 = rot(n2, INV);    // (-i) n  for forward, (+i) n for inverse
      out[j0] = {v[0].x + b1.x + b2.x, v[0].y + b1.y + b2.y};
      out[j0 + NS] = cadd(m1, in1); out[j0 + 4 * NS] = csub(m1, in1);
      out[j0 + 2 * NS] = cadd(m2, in2); out[j0 + 3 * NS] = csub(m2, in2);
    }
  }
  __syncwarp();
}

// One warp: 320-point complex FFT (Stockham autosort, radices 4,4,4,5) between two shared buffers; the result sits
// where it started (a).  INV: conjugate twiddles (no 1/N scaling).
template <bool INV>
__device__ __forceinline__ void fft320(cpx* a, cpx* b, const cpx* __restrict__ tw, int lane) {
  fft_stage<1, 4, INV>(a, b, tw, lane);
  fft_stage<4, 4, INV>(b, a, tw, lane);
  fft_stage<16, 4, INV>(a, b, tw, lane);
  fft_stage<64, 5, INV>(b, a, tw, lane);
}

// Twiddle / window tables, built once per device by gl_tables_kernel (the frames kernel used to spend a fifth of its
// instructions on sincosf for them in every CTA): [tw 320 cpx][tw2 322 cpx (321 used)][win 640 floats].
constexpr int TAB_FLOATS = 2 * NH + 2 * (NH + 2) + NFFT;
__device__ __align__(16) float g_tab[TAB_FLOATS];

__global__ void gl_tables_kernel() {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < NFFT; i += gridDim.x * blockDim.x) {
    float s, c;
    if (i < NH) {
      sincosf(-PI2 * (float)i / (float)NH, &s, &c);
      g_tab[2 * i] = c; g_tab[2 * i + 1] = s;
    }
    if (i < NH + 2) {
      sincosf(-PI2 * (float)i / (float)NFFT, &s, &c);
      g_tab[2 * NH + 2 * i] = c; g_tab[2 * NH + 2 * i + 1] = s;
    }
    g_tab[2 * NH + 2 * (NH + 2) + i] = 0.5f - 0.5f * cosf(PI2 * (float)i / (float)NFFT);
  }
}

__device__ __forceinline__ int reflect_idx(int j, int L) {   // F.pad(mode='reflect')
  if (j < 0) j = -j;
  if (j >= L) j = 2 * (L - 1) - j;
  return j;
}

// mode 0: phases from angles_t [B][T][321] (radians);  mode 1: phases from the STFT of sig [B][L].
// mag_t [B][T][321]; frames [B][T][640].  Optional spec_out [B][T][321][2] = STFT (re, im) of sig (mode 1 only).
__global__ void __launch_bounds__(WARPS * 32) gl_frames_kernel(int mode, const float* __restrict__ sig,
                                                               const float* __restrict__ angles_t,
                                                               const float* __restrict__ mag_t, float* __restrict__ frames,
                                                               float* __restrict__ spec_out, int B, int T, int L) {
  __shared__ __align__(16) float tab[TAB_FLOATS];
  __shared__ __align__(16) cpx buf[WARPS][2][NH + 2];   // stride 322 * 8 B keeps every buffer 16-byte aligned
  for (int i = threadIdx.x; i < TAB_FLOATS / 4; i += blockDim.x)
    reinterpret_cast<float4*>(tab)[i] = reinterpret_cast<const float4*>(g_tab)[i];
  __syncthreads();
  const cpx* tw = reinterpret_cast<const cpx*>(tab);                  // exp(-2 pi i m / 320), m < 320
  const cpx* tw2 = reinterpret_cast<const cpx*>(tab + 2 * NH);        // exp(-2 pi i k / 640), k <= 320
  const float* win = tab + 2 * NH + 2 * (NH + 2);                     // periodic Hann, 640
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.y;
  const int t = blockIdx.x * WARPS + warp;
  if (t >= T) return;
  cpx* A = buf[warp][0]; cpx* Bf = buf[warp][1];
  const long long fo = ((long long)b * T + t);
  const float* mg = mag_t + fo * NBIN;

  if (mode == 1) {
    // ---- forward: windowed frame -> packed complex z[n] = x[2n] + i x[2n+1] -> FFT320 -> split to X[0..320]
    const float* sb = sig + (long long)b * L;
    for (int n = lane; n < NH; n += 32) {
      const int i0 = t * HOP + 2 * n - NFFT / 2;
      A[n] = {sb[reflect_idx(i0, L)] * win[2 * n], sb[reflect_idx(i0 + 1, L)] * win[2 * n + 1]};
    }
    __syncwarp();
    fft320<false>(A, Bf, tw, lane);
    // X[k] = E[k] + W640^k O[k];  E = (Z[k] + conj Z[N-k])/2, O = (Z[k] - conj Z[N-k])/(2i).  Then Y[k] = mag * X/|X|.
    for (int k = lane; k <= NH; k += 32) {
      const cpx zk = A[k % NH], zn = cconj(A[(NH - k) % NH]);
      const cpx e = {0.5f * (zk.x + zn.x), 0.5f * (zk.y + zn.y)};
      const cpx d = csub(zk, zn);
      const cpx o = {0.5f * d.y, -0.5f * d.x};          // d / (2i)
      cpx X = (k == NH) ? csub(e, o) : cadd(e, cmul(tw2[k], o));
      if (spec_out) { spec_out[(fo * NBIN + k) * 2] = X.x; spec_out[(fo * NBIN + k) * 2 + 1] = X.y; }
      const float r2 = X.x * X.x + X.y * X.y;
      const float m = mg[k];
      const float sc = m * rsqrtf(r2);                                  // mag / |X|
      cpx Y = r2 > 0.f ? cpx{sc * X.x, sc * X.y} : cpx{m, 0.f};         // atan2(0,0) = 0
      Bf[k] = Y;
    }
  } else {
    const float* an = angles_t + fo * NBIN;
    for (int k = lane; k <= NH; k += 32) {
      float s, c;
      sincosf(an[k], &s, &c);
      const float m = mg[k];
      Bf[k] = {m * c, m * s};
    }
  }
  __syncwarp();
  // ---- inverse: Hermitian spectrum Y[0..320] (imaginary parts of DC / Nyquist ignored, as in the reference's basis)
  //      Z[k] = E[k] + i O[k],  E = (Y[k] + conj Y[N-k])/2,  O = (Y[k] - conj Y[N-k])/2 * W640^{-k}
  if (lane == 0) { Bf[0].y = 0.f; Bf[NH].y = 0.f; }
  __syncwarp();
  for (int k = lane; k < NH; k += 32) {
    const cpx yk = Bf[k], yn = cconj(Bf[NH - k]);
    const cpx e = {0.5f * (yk.x + yn.x), 0.5f * (yk.y + yn.y)};
    const cpx d = {0.5f * (yk.x - yn.x), 0.5f * (yk.y - yn.y)};
    const cpx o = cmul(d, cconj(tw2[k]));
    A[k] = {e.x - o.y, e.y + o.x};                      // e + i o
  }
  __syncwarp();
  fft320<true>(A, Bf, tw, lane);
  float* fr = frames + fo * NFFT;
  const float sc = 1.f / (float)NH;
  for (int n = lane; n < NH; n += 32) {
    const cpx z = A[n];
    float2 o = make_float2(z.x * sc * win[2 * n], z.y * sc * win[2 * n + 1]);
    *reinterpret_cast<float2*>(fr + 2 * n) = o;
  }
}

// sig_out[b][m] = (sum_t frames[b][t][m + 320 - 160 t]) / wss[m + 320]   for m in [0, L), L = 160 (T - 1)
__global__ void gl_ola_kernel(const float* __restrict__ frames, float* __restrict__ sig_out, int B, int T, int L) {
  const long long total = (long long)B * L;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / L), m = (int)(i % L);
    const int n = m + NFFT / 2;
    int t_hi = n / HOP; if (t_hi > T - 1) t_hi = T - 1;
    int t_lo = (n - NFFT + HOP) / HOP; if (n - NFFT + 1 <= 0) t_lo = 0;   // ceil((n-639)/160)
    if (t_lo < 0) t_lo = 0;
    float acc = 0.f, wss = 0.f;
    for (int t = t_lo; t <= t_hi; ++t) {
      const int r = n - t * HOP;
      if (r < 0 || r >= NFFT) continue;
      const float w = 0.5f - 0.5f * cosf(PI2 * (float)r / (float)NFFT);
      wss += w * w;
      acc += frames[((long long)b * T + t) * NFFT + r];
    }
    sig_out[i] = wss > 1.1754944e-38f ? acc / wss : acc;
  }
}

}  // namespace

extern "C" {

// One half-iteration of Griffin-Lim: frames[b][t][:] = window * irfft( mag[b][t][:] * unit_phase ), where the phase is
// exp(i*angles_t) (mode 0) or that of rfft(window * reflect-padded sig frame) (mode 1).  spec_out (optional,
// mode 1): the STFT itself as (re, im).  All tensors fp32; mag_t / angles_t are frame-major [B][T][321].
int vca_gl_frames(int mode, const float* sig, const float* angles_t, const float* mag_t, float* frames, float* spec_out, int B,
                  int T, int L, cudaStream_t s) {
  VCA_CHECK_ARG(mag_t && frames && B > 0 && T > 1 && L == HOP * (T - 1) && (mode == 0 ? angles_t != nullptr : sig != nullptr));
  VCA_CHECK_ARG(B <= 65535);
  static bool tables_ready[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64 || !tables_ready[dev]) {   // stream-ordered in front of the first use on this device
    gl_tables_kernel<<<3, 256, 0, s>>>();
    VCA_LAUNCH_CHECK();
    if (dev >= 0 && dev < 64) tables_ready[dev] = true;
  }
  dim3 grid((T + WARPS - 1) / WARPS, B);
  gl_frames_kernel<<<grid, WARPS * 32, 0, s>>>(mode, sig, angles_t, mag_t, frames, spec_out, B, T, L);
  VCA_LAUNCH_CHECK();
  return VCA_OK;
}
// Overlap-add + window-envelope normalisation + trim (the tail of STFT.inverse, stft.py:110-127).
int vca_gl_ola(const float* frames, float* sig_out, int B, int T, int L, cudaStream_t s) {
  VCA_CHECK_ARG(frames && sig_out && B > 0 && T > 1 && L == HOP * (T - 1));
  gl_ola_kernel<<<vca_grid_1d((long long)B * L, 256), 256, 0, s>>>(frames, sig_out, B, T, L);
  VCA_LAUNCH_CHECK();
  return VCA_OK;
}

}  // extern "C"
