// Volume-rendering compositing, forward and hand-written backward.  One warp per ray.
//
// Replaces src/models/SimpleNeRF17.py:486-539 (volume_rendering) and
// src/models/SimpleTensoRF09.py:767-819 (get_volume_rendering_weights + volume_render),
// NDC->world depth conversion of src/utils/CommonUtils04.py:208-224 fused in.
//
// HBM-bound: forward reads sigma, rgb, z (20 B/sample) and writes weights (+alpha, visibility when
// the caller keeps per-sample outputs); backward reads sigma, rgb, z, visibility (24 B/sample) and
// writes g_sigma, g_rgb (16 B/sample).  The exclusive transmittance product is a warp-shuffle scan
// with a running carry; the backward is one reverse pass with a suffix-sum scan (closed form in
// SURVEY.md §8a row VIII / oracle/composite.py).
#include "common.cuh"

namespace srf {

constexpr int CMP_WARPS = 8;

struct CompositeFwd {
  const float* sigma;      // [R,S]
  const float* rgb;        // [R,S,3] or nullptr
  const float* z;          // [R,S]
  const float* rays_o;     // [R,3] (ndc only)
  const float* rays_d;     // [R,3]
  const float* rays_d_ndc; // [R,3] (ndc only)
  float* alpha;            // [R,S] or nullptr
  float* visibility;       // [R,S] or nullptr
  float* weights;          // [R,S] (required: second pass re-reads it)
  float* rgb_map;          // [R,3] or nullptr
  float* acc;              // [R]
  float* depth;            // [R]
  float* depth_var;        // [R]
  float* depth_ndc;        // [R] (ndc only)
  float* depth_var_ndc;    // [R] (ndc only)
  long long R;
  int S;
  int ndc, white_bkgd;
  float distance_scale;
};

// world depth of an NDC depth (CommonUtils04.py:217-223): A * (1/(1 - z + [z==1]*1e-3) - 1) + tn
__device__ __forceinline__ float ndc_to_world(float zn, float A, float tn) {
  const float eps = zn == 1.f ? 1e-3f : 0.f;
  return __fadd_rn(__fmul_rn(A, __fadd_rn(__fdiv_rn(1.f, __fadd_rn(__fadd_rn(1.f, -zn), eps)), -1.f)), tn);
}

__device__ __forceinline__ float warp_incl_prod(float v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const float t = __shfl_up_sync(FULL, v, o);
    if (lane >= o) v *= t;
  }
  return v;
}

__global__ void __launch_bounds__(CMP_WARPS * 32) composite_fwd_kernel(CompositeFwd p) {
  __shared__ float s_rgb[CMP_WARPS][96];
  const int warp = threadIdx.x >> 5, lane = lane_id();
  const long long r = (long long)blockIdx.x * CMP_WARPS + warp;
  if (r >= p.R) return;
  const int S = p.S;
  const float* sig = p.sigma + r * S;
  const float* zz = p.z + r * S;
  float* wout = p.weights + r * S;

  const float* dsrc = p.ndc ? p.rays_d_ndc : p.rays_d;
  const float dx = dsrc[r * 3 + 0], dy = dsrc[r * 3 + 1], dz = dsrc[r * 3 + 2];
  const float nrm = sqrtf(dx * dx + dy * dy + dz * dz);
  const float far_z = p.ndc ? 1.f : 1e10f;
  float A = 0.f, tn = 0.f;
  if (p.ndc) {
    const float oz = p.rays_o[r * 3 + 2], wz = p.rays_d[r * 3 + 2];
    tn = __fdiv_rn(-__fadd_rn(1.f, oz), wz);
    A = __fdiv_rn(__fadd_rn(oz, __fmul_rn(tn, wz)), wz);
  }

  float carry = 1.f;
  float acc = 0.f, nz = 0.f, nzw = 0.f, c0 = 0.f, c1 = 0.f, c2 = 0.f;
  const int chunks = (S + 31) >> 5;
  for (int c = 0; c < chunks; ++c) {
    const int i = (c << 5) + lane;
    const bool ok = i < S;
    float zi = 0.f, zn = 0.f, sg = 0.f;
    if (ok) {
      zi = ldg_stream(zz + i);
      zn = (i + 1 < S) ? __ldg(zz + i + 1) : far_z;
      sg = ldg_stream(sig + i);
    }
    if (p.rgb != nullptr) {
      const long long base = (r * S + (c << 5)) * 3;
      const int lim = min(96, (S - (c << 5)) * 3);
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const int e = k * 32 + lane;
        if (e < lim) s_rgb[warp][e] = ldg_stream(p.rgb + base + e);
      }
    }
    const float delta = (zn - zi) * nrm;
    const float al = ok ? 1.f - expf(-sg * delta * p.distance_scale) : 0.f;
    const float q = ok ? (1.f - al + 1e-10f) : 1.f;
    const float incl = warp_incl_prod(q, lane);
    float excl = __shfl_up_sync(FULL, incl, 1);
    if (lane == 0) excl = 1.f;
    const float T = carry * excl;
    carry *= __shfl_sync(FULL, incl, 31);
    const float w = al * T;
    __syncwarp();
    if (ok) {
      stg_stream(wout + i, w);
      if (p.alpha) stg_stream(p.alpha + r * S + i, al);
      if (p.visibility) stg_stream(p.visibility + r * S + i, T);
      acc += w;
      nz += w * zi;
      if (p.ndc) nzw += w * ndc_to_world(zi, A, tn);
      if (p.rgb != nullptr) {
        c0 += w * s_rgb[warp][lane * 3 + 0];
        c1 += w * s_rgb[warp][lane * 3 + 1];
        c2 += w * s_rgb[warp][lane * 3 + 2];
      }
    }
    __syncwarp();
  }
  acc = warp_sum(acc);
  nz = warp_sum(nz);
  const float inv = 1.f / (acc + 1e-6f);
  const float d_main = nz * inv;          // NDC depth when ndc, world depth otherwise
  float d_world = d_main;
  if (p.ndc) d_world = warp_sum(nzw) * inv;
  // second pass: variances (each lane re-reads exactly the weights it wrote)
  float v_main = 0.f, v_world = 0.f;
  for (int i = lane; i < S; i += 32) {
    const float w = wout[i];
    const float zi = __ldg(zz + i);
    const float a = zi - d_main;
    v_main += w * a * a;
    if (p.ndc) {
      const float b = ndc_to_world(zi, A, tn) - d_world;
      v_world += w * b * b;
    }
  }
  v_main = warp_sum(v_main);
  if (p.ndc) v_world = warp_sum(v_world);
  if (p.rgb_map != nullptr) {
    c0 = warp_sum(c0); c1 = warp_sum(c1); c2 = warp_sum(c2);
    if (p.white_bkgd) { const float bg = 1.f - acc; c0 += bg; c1 += bg; c2 += bg; }
  }
  if (lane == 0) {
    p.acc[r] = acc;
    if (p.ndc) {
      p.depth_ndc[r] = d_main;
      p.depth_var_ndc[r] = v_main;
      p.depth[r] = d_world;
      p.depth_var[r] = v_world;
    } else {
      p.depth[r] = d_main;
      p.depth_var[r] = v_main;
    }
    if (p.rgb_map != nullptr) { p.rgb_map[r * 3 + 0] = c0; p.rgb_map[r * 3 + 1] = c1; p.rgb_map[r * 3 + 2] = c2; }
  }
}

// Rays of at most 32*K samples (the Simple-NeRF shapes: 64 coarse, 192 fine): every load of the ray is issued
// before any arithmetic (K*(2+3) independent loads in flight per lane), weights and depths stay in registers for
// the variance pass, and the successor depth comes from a shuffle instead of a second load.
template <int K>
__global__ void __launch_bounds__(CMP_WARPS * 32) composite_fwd_small_kernel(CompositeFwd p) {
  __shared__ float s_rgb[CMP_WARPS][K * 96];
  const int warp = threadIdx.x >> 5, lane = lane_id();
  const long long r = (long long)blockIdx.x * CMP_WARPS + warp;
  if (r >= p.R) return;
  const int S = p.S;
  const float* sig = p.sigma + r * S;
  const float* zz = p.z + r * S;
  const float far_z = p.ndc ? 1.f : 1e10f;
  float sg[K], zi[K];
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const int i = (k << 5) + lane;
    sg[k] = i < S ? ldg_stream(sig + i) : 0.f;
    zi[k] = i < S ? ldg_stream(zz + i) : far_z;
  }
  if (p.rgb != nullptr) {
    const float* src = p.rgb + r * S * 3;
#pragma unroll
    for (int k = 0; k < 3 * K; ++k) {
      const int e = (k << 5) + lane;
      if (e < S * 3) s_rgb[warp][e] = ldg_stream(src + e);
    }
  }
  const float* dsrc = p.ndc ? p.rays_d_ndc : p.rays_d;
  const float dx = dsrc[r * 3 + 0], dy = dsrc[r * 3 + 1], dz = dsrc[r * 3 + 2];
  const float nrm = sqrtf(dx * dx + dy * dy + dz * dz);
  float A = 0.f, tn = 0.f;
  if (p.ndc) {
    const float oz = p.rays_o[r * 3 + 2], wz = p.rays_d[r * 3 + 2];
    tn = __fdiv_rn(-__fadd_rn(1.f, oz), wz);
    A = __fdiv_rn(__fadd_rn(oz, __fmul_rn(tn, wz)), wz);
  }
  __syncwarp();
  float carry = 1.f, acc = 0.f, nz = 0.f, nzw = 0.f, c0 = 0.f, c1 = 0.f, c2 = 0.f;
  float w[K], zw[K];
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const int i = (k << 5) + lane;
    const bool ok = i < S;
    float zn = __shfl_down_sync(FULL, zi[k], 1);
    const float head_next = (k + 1 < K) ? __shfl_sync(FULL, zi[(k + 1 < K) ? k + 1 : k], 0) : far_z;
    if (lane == 31) zn = head_next;
    if (i >= S - 1) zn = far_z;
    const float delta = (zn - zi[k]) * nrm;
    const float al = ok ? 1.f - expf(-sg[k] * delta * p.distance_scale) : 0.f;
    const float q = ok ? (1.f - al + 1e-10f) : 1.f;
    const float incl = warp_incl_prod(q, lane);
    float excl = __shfl_up_sync(FULL, incl, 1);
    if (lane == 0) excl = 1.f;
    const float T = carry * excl;
    carry *= __shfl_sync(FULL, incl, 31);
    w[k] = ok ? al * T : 0.f;
    zw[k] = p.ndc ? ndc_to_world(zi[k], A, tn) : zi[k];
    if (ok) {
      stg_stream(p.weights + r * S + i, w[k]);
      if (p.alpha) stg_stream(p.alpha + r * S + i, al);
      if (p.visibility) stg_stream(p.visibility + r * S + i, T);
      acc += w[k];
      nz += w[k] * zi[k];
      nzw += w[k] * zw[k];
      if (p.rgb != nullptr) {
        c0 += w[k] * s_rgb[warp][i * 3 + 0];
        c1 += w[k] * s_rgb[warp][i * 3 + 1];
        c2 += w[k] * s_rgb[warp][i * 3 + 2];
      }
    }
  }
  acc = warp_sum(acc);
  nz = warp_sum(nz);
  const float inv = 1.f / (acc + 1e-6f);
  const float d_main = nz * inv;
  float d_world = d_main;
  if (p.ndc) d_world = warp_sum(nzw) * inv;
  float v_main = 0.f, v_world = 0.f;
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const float a = zi[k] - d_main;
    v_main += w[k] * a * a;
    const float b = zw[k] - d_world;
    v_world += w[k] * b * b;
  }
  v_main = warp_sum(v_main);
  if (p.ndc) v_world = warp_sum(v_world);
  if (p.rgb_map != nullptr) {
    c0 = warp_sum(c0); c1 = warp_sum(c1); c2 = warp_sum(c2);
    if (p.white_bkgd) { const float bg = 1.f - acc; c0 += bg; c1 += bg; c2 += bg; }
  }
  if (lane == 0) {
    p.acc[r] = acc;
    if (p.ndc) {
      p.depth_ndc[r] = d_main; p.depth_var_ndc[r] = v_main; p.depth[r] = d_world; p.depth_var[r] = v_world;
    } else {
      p.depth[r] = d_main; p.depth_var[r] = v_main;
    }
    if (p.rgb_map != nullptr) { p.rgb_map[r * 3 + 0] = c0; p.rgb_map[r * 3 + 1] = c1; p.rgb_map[r * 3 + 2] = c2; }
  }
}

struct CompositeBwd {
  const float* sigma; const float* rgb; const float* z; const float* visibility;
  const float* rays_o; const float* rays_d; const float* rays_d_ndc;
  const float* acc; const float* depth; const float* depth_ndc;      // saved per-ray forward outputs
  const float* g_rgb;        // [R,3] or nullptr
  const float* g_acc;        // [R] or nullptr
  const float* g_depth;      // [R] or nullptr
  const float* g_depth_ndc;  // [R] or nullptr
  const float* g_depth_var;     // [R] or nullptr
  const float* g_depth_var_ndc; // [R] or nullptr
  const float* g_weights;    // [R,S] or nullptr
  float* g_sigma;            // [R,S]
  float* g_rgb_s;            // [R,S,3] or nullptr
  long long R;
  int S;
  int ndc, white_bkgd;
  float distance_scale;
};

__global__ void __launch_bounds__(CMP_WARPS * 32) composite_bwd_kernel(CompositeBwd p) {
  __shared__ float s_rgb[CMP_WARPS][96];
  const int warp = threadIdx.x >> 5, lane = lane_id();
  const long long r = (long long)blockIdx.x * CMP_WARPS + warp;
  if (r >= p.R) return;
  const int S = p.S;
  const float* sig = p.sigma + r * S;
  const float* zz = p.z + r * S;
  const float* vis = p.visibility + r * S;

  const float* dsrc = p.ndc ? p.rays_d_ndc : p.rays_d;
  const float dx = dsrc[r * 3 + 0], dy = dsrc[r * 3 + 1], dz = dsrc[r * 3 + 2];
  const float nrm = sqrtf(dx * dx + dy * dy + dz * dz);
  const float far_z = p.ndc ? 1.f : 1e10f;
  float A = 0.f, tn = 0.f;
  if (p.ndc) {
    const float oz = p.rays_o[r * 3 + 2], wz = p.rays_d[r * 3 + 2];
    tn = __fdiv_rn(-__fadd_rn(1.f, oz), wz);
    A = __fdiv_rn(__fadd_rn(oz, __fmul_rn(tn, wz)), wz);
  }
  const float acc = p.acc[r];
  const float invA = 1.f / (acc + 1e-6f);
  const float d_world = p.depth[r];
  const float d_ndc = p.ndc ? p.depth_ndc[r] : 0.f;
  float gr0 = 0.f, gr1 = 0.f, gr2 = 0.f;
  if (p.g_rgb) { gr0 = p.g_rgb[r * 3 + 0]; gr1 = p.g_rgb[r * 3 + 1]; gr2 = p.g_rgb[r * 3 + 2]; }
  float g_acc = p.g_acc ? p.g_acc[r] : 0.f;
  if (p.white_bkgd) g_acc -= (gr0 + gr1 + gr2);
  // depth gradients: "main" is the depth on the z the ray was sampled in (NDC z when ndc)
  const float gd_world = p.g_depth ? p.g_depth[r] : 0.f;
  const float gd_ndc = (p.ndc && p.g_depth_ndc) ? p.g_depth_ndc[r] : 0.f;
  const float gv_world = p.g_depth_var ? p.g_depth_var[r] : 0.f;
  const float gv_ndc = (p.ndc && p.g_depth_var_ndc) ? p.g_depth_var_ndc[r] : 0.f;
  // N - depth*acc = depth*(acc+1e-6) - depth*acc = depth*1e-6 (exactly the forward's N)
  const float kv_world = 2.f * (d_world * 1e-6f) * invA;
  const float kv_ndc = 2.f * (d_ndc * 1e-6f) * invA;

  float carry = 0.f;    // sum_{k > current chunk} g_w[k] w_k
  const int chunks = (S + 31) >> 5;
  const bool has_rgb = p.rgb != nullptr && p.g_rgb != nullptr;
  for (int c = chunks - 1; c >= 0; --c) {
    const int i = (c << 5) + lane;
    const bool ok = i < S;
    float zi = 0.f, zn = 0.f, sg = 0.f, T = 0.f;
    if (ok) {
      zi = ldg_stream(zz + i);
      zn = (i + 1 < S) ? __ldg(zz + i + 1) : far_z;
      sg = ldg_stream(sig + i);
      T = ldg_stream(vis + i);
    }
    const long long base = (r * S + (c << 5)) * 3;
    const int lim = min(96, (S - (c << 5)) * 3);
    if (has_rgb) {
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const int e = k * 32 + lane;
        if (e < lim) s_rgb[warp][e] = ldg_stream(p.rgb + base + e);
      }
    }
    __syncwarp();
    const float dscale = (zn - zi) * nrm * p.distance_scale;
    const float e = ok ? expf(-sg * dscale) : 1.f;
    const float al = 1.f - e;
    const float q = al == 1.f ? 1e-10f : (1.f - al + 1e-10f);
    const float w = al * T;
    float gw = g_acc;
    if (p.g_weights && ok) gw += ldg_stream(p.g_weights + r * S + i);
    if (has_rgb && ok)
      gw += gr0 * s_rgb[warp][lane * 3 + 0] + gr1 * s_rgb[warp][lane * 3 + 1] + gr2 * s_rgb[warp][lane * 3 + 2];
    if (p.ndc) {
      const float a = zi - d_ndc;
      gw += gd_ndc * a * invA + gv_ndc * (a * a - kv_ndc * a);
      const float b = ndc_to_world(zi, A, tn) - d_world;
      gw += gd_world * b * invA + gv_world * (b * b - kv_world * b);
    } else {
      const float b = zi - d_world;
      gw += gd_world * b * invA + gv_world * (b * b - kv_world * b);
    }
    const float gww = ok ? gw * w : 0.f;
    // inclusive suffix sum within the chunk (towards higher lanes)
    float suf = gww;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const float t = __shfl_down_sync(FULL, suf, o);
      if (lane + o < 32) suf += t;
    }
    const float later = carry + (suf - gww);     // sum over k > i
    carry += __shfl_sync(FULL, suf, 0);
    const float g_alpha = gw * T - later / q;
    __syncwarp();
    if (ok) stg_stream(p.g_sigma + r * S + i, g_alpha * (dscale * e));
    if (p.g_rgb_s != nullptr) {
      if (ok) {
        s_rgb[warp][lane * 3 + 0] = w * gr0;
        s_rgb[warp][lane * 3 + 1] = w * gr1;
        s_rgb[warp][lane * 3 + 2] = w * gr2;
      }
      __syncwarp();
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const int e2 = k * 32 + lane;
        if (e2 < lim) stg_stream(p.g_rgb_s + base + e2, s_rgb[warp][e2]);
      }
    }
    __syncwarp();
  }
}

}  // namespace srf

using namespace srf;

SRF_API int srf_composite_fwd(const float* sigma, const float* rgb, const float* z, const float* rays_o,
                              const float* rays_d, const float* rays_d_ndc, int64_t num_rays, int num_samples,
                              int ndc, int white_bkgd, float distance_scale, float* alpha, float* visibility,
                              float* weights, float* rgb_map, float* acc, float* depth, float* depth_var,
                              float* depth_ndc, float* depth_var_ndc, void* stream) {
  if (num_rays == 0) return 0;
  SRF_REQUIRE(sigma && z && rays_d && weights && acc && depth && depth_var, "srf_composite_fwd", "null pointer");
  SRF_REQUIRE(!ndc || (rays_o && rays_d_ndc && depth_ndc && depth_var_ndc), "srf_composite_fwd",
              "ndc needs rays_o, rays_d_ndc, depth_ndc, depth_var_ndc");
  SRF_REQUIRE((rgb == nullptr) == (rgb_map == nullptr), "srf_composite_fwd", "rgb and rgb_map go together");
  SRF_REQUIRE(num_samples > 0 && num_rays >= 0, "srf_composite_fwd", "bad sizes");
  CompositeFwd p{sigma, rgb, z, rays_o, rays_d, rays_d_ndc, alpha, visibility, weights, rgb_map, acc, depth,
                 depth_var, depth_ndc, depth_var_ndc, num_rays, num_samples, ndc, white_bkgd, distance_scale};
  const unsigned blocks = (unsigned)((num_rays + CMP_WARPS - 1) / CMP_WARPS);
  if (num_samples <= 64) composite_fwd_small_kernel<2><<<blocks, CMP_WARPS * 32, 0, (cudaStream_t)stream>>>(p);
  else if (num_samples <= 192) composite_fwd_small_kernel<6><<<blocks, CMP_WARPS * 32, 0, (cudaStream_t)stream>>>(p);
  else composite_fwd_kernel<<<blocks, CMP_WARPS * 32, 0, (cudaStream_t)stream>>>(p);
  return check_launch("srf_composite_fwd");
}

SRF_API int srf_composite_bwd(const float* sigma, const float* rgb, const float* z, const float* visibility,
                              const float* rays_o, const float* rays_d, const float* rays_d_ndc, const float* acc,
                              const float* depth, const float* depth_ndc, const float* g_rgb, const float* g_acc,
                              const float* g_depth, const float* g_depth_ndc, const float* g_depth_var,
                              const float* g_depth_var_ndc, const float* g_weights, int64_t num_rays,
                              int num_samples, int ndc, int white_bkgd, float distance_scale, float* g_sigma,
                              float* g_rgb_samples, void* stream) {
  if (num_rays == 0) return 0;
  SRF_REQUIRE(sigma && z && visibility && rays_d && acc && depth && g_sigma, "srf_composite_bwd", "null pointer");
  SRF_REQUIRE(!ndc || (rays_o && rays_d_ndc && depth_ndc), "srf_composite_bwd", "ndc needs rays_o, rays_d_ndc, depth_ndc");
  SRF_REQUIRE(g_rgb_samples == nullptr || (rgb && g_rgb), "srf_composite_bwd", "g_rgb_samples needs rgb and g_rgb");
  SRF_REQUIRE(num_samples > 0 && num_rays >= 0, "srf_composite_bwd", "bad sizes");
  CompositeBwd p{sigma, rgb, z, visibility, rays_o, rays_d, rays_d_ndc, acc, depth, depth_ndc, g_rgb, g_acc, g_depth,
                 g_depth_ndc, g_depth_var, g_depth_var_ndc, g_weights, g_sigma, g_rgb_samples, num_rays, num_samples,
                 ndc, white_bkgd, distance_scale};
  const unsigned blocks = (unsigned)((num_rays + CMP_WARPS - 1) / CMP_WARPS);
  composite_bwd_kernel<<<blocks, CMP_WARPS * 32, 0, (cudaStream_t)stream>>>(p);
  return check_launch("srf_composite_bwd");
}
