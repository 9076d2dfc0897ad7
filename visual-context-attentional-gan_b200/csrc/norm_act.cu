// Stand-alone element-wise kernels (activations, axpby, mul, cast) on flat tensors: one coalesced pass, 128-bit
// accesses, 8 elements per thread per iteration.  (BatchNorm and the column sums live in bn.cu.)
// Reference sites: generator.py:105-126,179,209-225,325-329 (LeakyReLU / tanh / residual adds).
#include "vec.cuh"

namespace {

// ---- flat element-wise kernels (8 elements per thread per iteration) ------------------------------------------
struct OpLRelu { float s; __device__ float operator()(float a, float) const { return a > 0.f ? a : a * s; } };
struct OpLReluBwd { float s; __device__ float operator()(float g, float x) const { return x > 0.f ? g : g * s; } };
struct OpTanh { __device__ float operator()(float a, float) const { return tanhf(a); } };
struct OpTanhBwd { __device__ float operator()(float g, float y) const { return g * (1.f - y * y); } };
struct OpAxpby { float al, be; __device__ float operator()(float a, float b) const { return fmaf(be, b, al * a); } };
struct OpMul { __device__ float operator()(float a, float b) const { return a * b; } };

template <class TI, class TO, class Op, bool TWO>
__global__ void __launch_bounds__(256) ew_kernel(const TI* __restrict__ a, const TI* __restrict__ b, TO* __restrict__ out,
                                                 long long n, Op op, bool vec_ok) {
  constexpr int E = 8;
  const long long nv = vec_ok ? n / E : 0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nv; i += (long long)gridDim.x * blockDim.x) {
    float x[E], y[E];
    constexpr int VI = Vec<TI>::N, VO = Vec<TO>::N;
#pragma unroll
    for (int k = 0; k < E / VI; ++k) Vec<TI>::load(a + i * E + k * VI, x + k * VI);
    if (TWO) {
#pragma unroll
      for (int k = 0; k < E / VI; ++k) Vec<TI>::load(b + i * E + k * VI, y + k * VI);
    }
#pragma unroll
    for (int k = 0; k < E; ++k) x[k] = op(x[k], TWO ? y[k] : 0.f);
#pragma unroll
    for (int k = 0; k < E / VO; ++k) Vec<TO>::store(out + i * E + k * VO, x + k * VO);
  }
  for (long long i = nv * E + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = from_f<TO>(op(to_f(a[i]), TWO ? to_f(b[i]) : 0.f));
}

template <class TI, class TO, class Op>
int ew_launch(const void* a, const void* b, void* out, long long n, Op op, cudaStream_t s) {
  if (n <= 0) return VCA_OK;
  const bool vec_ok = vca_aligned16(a) && vca_aligned16(out) && (!b || vca_aligned16(b));
  unsigned grid = vca_grid_1d(n, 256, 8);
  if (b) ew_kernel<TI, TO, Op, true><<<grid, 256, 0, s>>>((const TI*)a, (const TI*)b, (TO*)out, n, op, vec_ok);
  else ew_kernel<TI, TO, Op, false><<<grid, 256, 0, s>>>((const TI*)a, nullptr, (TO*)out, n, op, vec_ok);
  VCA_LAUNCH_CHECK();
  return VCA_OK;
}
template <class Op>
int ew_dispatch(int dtype, const void* a, const void* b, void* out, long long n, Op op, cudaStream_t s) {
  return dtype == VCA_F32 ? ew_launch<float, float, Op>(a, b, out, n, op, s) : ew_launch<bf16, bf16, Op>(a, b, out, n, op, s);
}

}  // namespace

extern "C" {

int vca_lrelu_fwd(int dtype, const void* x, void* y, long long n, float slope, cudaStream_t s) {
  VCA_CHECK_ARG(x && y && n >= 0);
  return ew_dispatch(dtype, x, nullptr, y, n, OpLRelu{slope}, s);
}
int vca_lrelu_bwd(int dtype, const void* dy, const void* x, void* dx, long long n, float slope, cudaStream_t s) {
  VCA_CHECK_ARG(dy && x && dx && n >= 0);
  return ew_dispatch(dtype, dy, x, dx, n, OpLReluBwd{slope}, s);
}
int vca_tanh_fwd(int dtype, const void* x, void* y, long long n, cudaStream_t s) {
  VCA_CHECK_ARG(x && y && n >= 0);
  return ew_dispatch(dtype, x, nullptr, y, n, OpTanh{}, s);
}
int vca_tanh_bwd(int dtype, const void* dy, const void* y, void* dx, long long n, cudaStream_t s) {
  VCA_CHECK_ARG(dy && y && dx && n >= 0);
  return ew_dispatch(dtype, dy, y, dx, n, OpTanhBwd{}, s);
}
int vca_axpby(int dtype, const void* a, const void* b, void* out, long long n, float alpha, float beta, cudaStream_t s) {
  VCA_CHECK_ARG(a && out && n >= 0);
  return ew_dispatch(dtype, a, b, out, n, OpAxpby{alpha, b ? beta : 0.f}, s);
}
int vca_mul(int dtype, const void* x, const void* m, void* y, long long n, cudaStream_t s) {
  VCA_CHECK_ARG(x && m && y && n > 0);
  return ew_dispatch(dtype, x, m, y, n, OpMul{}, s);
}
int vca_cast(int dt_in, int dt_out, const void* x, void* y, long long n, cudaStream_t s) {
  VCA_CHECK_ARG(x && y && n >= 0);
  OpAxpby id{1.f, 0.f};
  if (dt_in == VCA_F32 && dt_out == VCA_BF16) return ew_launch<float, bf16, OpAxpby>(x, nullptr, y, n, id, s);
  if (dt_in == VCA_BF16 && dt_out == VCA_F32) return ew_launch<bf16, float, OpAxpby>(x, nullptr, y, n, id, s);
  if (dt_in == VCA_F32) return ew_launch<float, float, OpAxpby>(x, nullptr, y, n, id, s);
  return ew_launch<bf16, bf16, OpAxpby>(x, nullptr, y, n, id, s);
}
}  // extern "C"
