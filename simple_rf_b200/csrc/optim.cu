// "Next" row f4: the optimiser tail of a training iteration.  One flat fused Adam step replaces what
// torch.optim.Adam (created by src/optimizers/OptimizerFactory02.py:9-22 with lr / betas only, stepped at
// src/Trainer10.py:109-110) runs as ~10 multi-tensor passes over ~100 small parameter tensors (56 launches, 13 % of a
// Simple-NeRF iteration): parameters, gradients and both moments are flat fp32 arrays, one thread handles 4 elements.
//
// Arithmetic follows torch/optim/adam.py::_single_tensor_adam (non-capturable, amsgrad off, maximize off):
//   m += (g - m) (1 - b1);   v = v b2 + (1 - b2) g g;   p -= (lr / (1 - b1^t)) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
// with optional L2 weight decay folded into g first.  HBM-bound: 16 B read + 12 B written per parameter.
#include "common.cuh"

namespace srf {

struct AdamParams {
  float* p; const float* g; float* m; float* v;
  long long n;
  float one_minus_b1, b2, one_minus_b2, step_size, inv_bc2_sqrt, eps, weight_decay;
};

__device__ __forceinline__ void adam_one(float& p, float g, float& m, float& v, const AdamParams& a) {
  if (a.weight_decay != 0.f) g = __fadd_rn(g, __fmul_rn(a.weight_decay, p));
  m = __fadd_rn(m, __fmul_rn(__fsub_rn(g, m), a.one_minus_b1));                 // lerp_(g, 1 - b1), weight < 0.5 form
  v = __fadd_rn(__fmul_rn(v, a.b2), __fmul_rn(__fmul_rn(a.one_minus_b2, g), g));   // mul_(b2).addcmul_(g, g, 1 - b2)
  const float denom = __fadd_rn(__fmul_rn(sqrtf(v), a.inv_bc2_sqrt), a.eps);
  p = __fsub_rn(p, __fmul_rn(a.step_size, __fdiv_rn(m, denom)));               // addcdiv_(m, denom, -step_size)
}

__global__ void __launch_bounds__(256) adam_kernel(const AdamParams a) {
  const long long n4 = a.n >> 2;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 p = reinterpret_cast<float4*>(a.p)[i], m = reinterpret_cast<float4*>(a.m)[i], v = reinterpret_cast<float4*>(a.v)[i];
    const float4 g = __ldg(reinterpret_cast<const float4*>(a.g) + i);
    adam_one(p.x, g.x, m.x, v.x, a); adam_one(p.y, g.y, m.y, v.y, a);
    adam_one(p.z, g.z, m.z, v.z, a); adam_one(p.w, g.w, m.w, v.w, a);
    reinterpret_cast<float4*>(a.p)[i] = p; reinterpret_cast<float4*>(a.m)[i] = m; reinterpret_cast<float4*>(a.v)[i] = v;
  }
  // tail (n % 4 elements)
  const long long t = (n4 << 2) + (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < a.n) adam_one(a.p[t], a.g[t], a.m[t], a.v[t], a);
}

}  // namespace srf

using namespace srf;

SRF_API int srf_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, float lr, float beta1,
                          float beta2, float eps, float weight_decay, int64_t step, void* stream) {
  if (n == 0) return 0;
  SRF_REQUIRE(params && grads && exp_avg && exp_avg_sq, "srf_adam_step", "null pointer");
  SRF_REQUIRE(step >= 1, "srf_adam_step", "step counts from 1");
  SRF_REQUIRE(((uintptr_t)params | (uintptr_t)grads | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq) % 16 == 0, "srf_adam_step",
              "buffers must be 16-byte aligned");
  AdamParams a{};
  a.p = params; a.g = grads; a.m = exp_avg; a.v = exp_avg_sq; a.n = n;
  // scalar part in double, as the Python floats of torch/optim/adam.py
  const double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
  a.one_minus_b1 = (float)(1.0 - (double)beta1);
  a.b2 = beta2;
  a.one_minus_b2 = (float)(1.0 - (double)beta2);
  a.step_size = (float)((double)lr / bc1);
  a.inv_bc2_sqrt = (float)(1.0 / sqrt(bc2));
  a.eps = eps;
  a.weight_decay = weight_decay;
  const long long n4 = (n + 3) / 4;
  long long blocks = (n4 + 255) / 256;
  const long long cap = (long long)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  adam_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(a);
  return check_launch("srf_adam_step");
}

// ---- capturable form: the step count and the learning rate live in device memory, so a CUDA graph that contains the step
// replays correctly (torch.optim.Adam(capturable=True) does the same).  `srf_adam_advance` increments the counter (one
// thread); the step kernel derives the bias corrections from it in double, once per block.
namespace srf {

__global__ void adam_advance_kernel(long long* step) { *step += 1; }

__global__ void __launch_bounds__(256) adam_dev_kernel(AdamParams a, const long long* __restrict__ step, const float* __restrict__ lr,
                                                       float beta1, float beta2) {
  __shared__ float s_step_size, s_inv_bc2_sqrt;
  if (threadIdx.x == 0) {
    const double t = (double)*step;
    const double bc1 = 1.0 - pow((double)beta1, t), bc2 = 1.0 - pow((double)beta2, t);
    s_step_size = (float)((double)*lr / bc1);
    s_inv_bc2_sqrt = (float)(1.0 / sqrt(bc2));
  }
  __syncthreads();
  a.step_size = s_step_size;
  a.inv_bc2_sqrt = s_inv_bc2_sqrt;
  const long long n4 = a.n >> 2;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 p = reinterpret_cast<float4*>(a.p)[i], m = reinterpret_cast<float4*>(a.m)[i], v = reinterpret_cast<float4*>(a.v)[i];
    const float4 g = __ldg(reinterpret_cast<const float4*>(a.g) + i);
    adam_one(p.x, g.x, m.x, v.x, a); adam_one(p.y, g.y, m.y, v.y, a);
    adam_one(p.z, g.z, m.z, v.z, a); adam_one(p.w, g.w, m.w, v.w, a);
    reinterpret_cast<float4*>(a.p)[i] = p; reinterpret_cast<float4*>(a.m)[i] = m; reinterpret_cast<float4*>(a.v)[i] = v;
  }
  const long long t = (n4 << 2) + (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < a.n) adam_one(a.p[t], a.g[t], a.m[t], a.v[t], a);
}

}  // namespace srf

SRF_API int srf_adam_advance(int64_t* step, void* stream) {
  SRF_REQUIRE(step, "srf_adam_advance", "null pointer");
  srf::adam_advance_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(reinterpret_cast<long long*>(step));
  return srf::check_launch("srf_adam_advance");
}

SRF_API int srf_adam_step_capturable(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, const float* lr,
                                     float beta1, float beta2, float eps, float weight_decay, const int64_t* step, void* stream) {
  using namespace srf;
  if (n == 0) return 0;
  SRF_REQUIRE(params && grads && exp_avg && exp_avg_sq && lr && step, "srf_adam_step_capturable", "null pointer");
  SRF_REQUIRE(((uintptr_t)params | (uintptr_t)grads | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq) % 16 == 0, "srf_adam_step_capturable",
              "buffers must be 16-byte aligned");
  AdamParams a{};
  a.p = params; a.g = grads; a.m = exp_avg; a.v = exp_avg_sq; a.n = n;
  a.one_minus_b1 = (float)(1.0 - (double)beta1);
  a.b2 = beta2;
  a.one_minus_b2 = (float)(1.0 - (double)beta2);
  a.eps = eps;
  a.weight_decay = weight_decay;
  const long long n4 = (n + 3) / 4;
  long long blocks = (n4 + 255) / 256;
  const long long cap = (long long)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  adam_dev_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(a, reinterpret_cast<const long long*>(step), lr, beta1, beta2);
  return check_launch("srf_adam_step_capturable");
}
