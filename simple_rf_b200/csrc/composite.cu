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
  float* weights;          // [R,S]
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
  int vec;                 // every per-sample pointer is 16-byte aligned: 128-bit loads / stores
};

// world depth of an NDC depth (CommonUtils04.py:217-223): A * (1/(1 - z + [z==1]*1e-3) - 1) + tn.  The per-ray constants
// keep IEEE divisions; the per-sample reciprocal is the 1-ulp MUFU approximation (the kernels are issue-bound on exactly
// this per-sample arithmetic, and the tolerance of this fp32 path is 1e-3).
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float sqrt_approx(float x) {
  float y;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float ndc_to_world(float zn, float A, float tn) {
  const float eps = zn == 1.f ? 1e-3f : 0.f;
  return fmaf(A, rcp_approx(__fadd_rn(__fadd_rn(1.f, -zn), eps)) - 1.f, tn);
}

__device__ __forceinline__ float warp_incl_prod(float v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const float t = __shfl_up_sync(FULL, v, o);
    if (lane >= o) v *= t;
  }
  return v;
}

__device__ __forceinline__ void stg_stream4(float* p, float a, float b, float c, float d) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void stg_stream2(float* p, float a, float b) {
  asm volatile("st.global.L1::no_allocate.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(a), "f"(b) : "memory");
}

// N consecutive floats at `src` (N = L or 3 L).  `fast`: the whole run belongs to the ray and the address is aligned
// to the vector piece (128-bit pieces when N % 4 == 0, else 64-bit; L1-allocating loads, because a lane's pieces of
// the 3 L-wide rgb run share sectors with its neighbours').  Otherwise element k is read iff its local index
// i0 + k / PER lies in [0, S).
template <int N, int PER>
__device__ __forceinline__ void load_run(const float* src, bool fast, int i0, int S, float (&v)[N]) {
  if (fast) {
    if (N % 4 == 0) {
#pragma unroll
      for (int k = 0; k < N / 4; ++k) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(src) + k);
        v[4 * k] = t.x; v[4 * k + 1] = t.y; v[4 * k + 2] = t.z; v[4 * k + 3] = t.w;
      }
    } else {
#pragma unroll
      for (int k = 0; k < N / 2; ++k) {
        const float2 t = __ldg(reinterpret_cast<const float2*>(src) + k);
        v[2 * k] = t.x; v[2 * k + 1] = t.y;
      }
    }
  } else {
#pragma unroll
    for (int k = 0; k < N; ++k) v[k] = (unsigned)(i0 + k / PER) < (unsigned)S ? __ldg(src + k) : 0.f;
  }
}

template <int N, int PER>
__device__ __forceinline__ void store_run(float* dst, bool fast, int i0, int S, const float (&v)[N]) {
  if (fast) {
    if (N % 4 == 0) {
#pragma unroll
      for (int k = 0; k < N / 4; ++k) stg_stream4(dst + 4 * k, v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
    } else {
#pragma unroll
      for (int k = 0; k < N / 2; ++k) stg_stream2(dst + 2 * k, v[2 * k], v[2 * k + 1]);
    }
  } else {
#pragma unroll
    for (int k = 0; k < N; ++k)
      if ((unsigned)(i0 + k / PER) < (unsigned)S) stg_stream(dst + k, v[k]);
  }
}

// Sums of 8 per-lane values over the warp in 4 + 2 + 1 + 2 = 9 shuffles (instead of 8 x 5): each exchange step halves
// the number of values a lane carries.  On return lane 4k (k = 0..7) holds the total of value k.
__device__ __forceinline__ float warp_reduce8(const float (&v)[8], int lane) {
  float a[4], b[2], c;
  const bool u16 = lane & 16, u8 = lane & 8, u4 = lane & 4;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float send = u16 ? v[k] : v[k + 4];
    const float keep = u16 ? v[k + 4] : v[k];
    a[k] = keep + __shfl_xor_sync(FULL, send, 16);
  }
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const float send = u8 ? a[k] : a[k + 2];
    const float keep = u8 ? a[k + 2] : a[k];
    b[k] = keep + __shfl_xor_sync(FULL, send, 8);
  }
  {
    const float send = u4 ? b[0] : b[1];
    const float keep = u4 ? b[1] : b[0];
    c = keep + __shfl_xor_sync(FULL, send, 4);
  }
  c += __shfl_xor_sync(FULL, c, 2);
  c += __shfl_xor_sync(FULL, c, 1);
  return c;
}
// two values in 1 + 4 shuffles: lanes 0..15 end with the total of x, lanes 16..31 with the total of y
__device__ __forceinline__ float warp_reduce2(float x, float y, int lane) {
  const bool u16 = lane & 16;
  float c = (u16 ? y : x) + __shfl_xor_sync(FULL, u16 ? x : y, 16);
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) c += __shfl_xor_sync(FULL, c, o);
  return c;
}

struct RayConsts { float nrm, far_z, A, tn; };

template <typename P>
__device__ __forceinline__ RayConsts ray_consts(const P& p, long long r) {
  RayConsts k;
  const float* dsrc = p.ndc ? p.rays_d_ndc : p.rays_d;
  const float dx = dsrc[r * 3 + 0], dy = dsrc[r * 3 + 1], dz = dsrc[r * 3 + 2];
  k.nrm = sqrt_approx(dx * dx + dy * dy + dz * dz);
  k.far_z = p.ndc ? 1.f : 1e10f;
  k.A = 0.f; k.tn = 0.f;
  if (p.ndc) {
    const float oz = p.rays_o[r * 3 + 2], wz = p.rays_d[r * 3 + 2];
    const float iw = rcp_approx(wz);
    k.tn = -(1.f + oz) * iw;
    k.A = (oz + k.tn * wz) * iw;
  }
  return k;
}

// Forward.  One warp per ray; a lane owns RUNS of L consecutive samples (32 L samples per warp step): vector loads per
// input array and run ([R,S,3] rgb needs no shared-memory transpose this way), the transmittance product is sequential
// inside the run with ONE 5-step shuffle scan per step across lanes, and the per-ray sums are reduced once at the end
// with the multi-value butterfly above.  Runs are aligned in the FLAT [R*S] arrays, so any S (e.g. TensoRF's 1083)
// vectorises: elements of the neighbouring rays that fall into a ray's first / last run are loaded but masked.
// L is chosen by the host so that the steps cover S with few idle lanes (64 -> 2, 128 -> 4, 192 -> 6, 256 -> 8).
// NC == 1: a single step, weights and depths stay in registers for the variance pass; NC == 0: any length, the variance
// pass re-reads what the lane wrote (L2 hits).  Keeping 2-4 steps in registers was measured slower (512 samples: 47 % ->
// 60 % of HBM peak without it): the registers cost more occupancy than the re-read costs bandwidth.
template <int L, int NC>
__device__ __forceinline__ void composite_fwd_ray(const CompositeFwd& p, long long r, int lane) {
  constexpr int G = L % 4 == 0 ? 4 : 2;          // alignment granule of a run (floats)
  const int S = p.S;
  const long long first = r * S;
  const int lead = (int)(first & (G - 1));       // elements of the previous ray in this ray's first run
  const int nchunks = (lead + S + 32 * L - 1) / (32 * L);
  const bool vec = p.vec != 0;
  const bool has_rgb = p.rgb != nullptr;
  const float far_z = p.ndc ? 1.f : 1e10f;

  float carry = 1.f;
  float red[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};     // acc, sum w z, sum w z_world, r, g, b
  constexpr int KEEP = NC > 0 ? NC : 1;
  float wk[KEEP][L], zk[KEEP][L], zwk[KEEP][L];
  float sgk[KEEP][L], colk[KEEP][3 * L], znk[KEEP];

  // Single-step rays (NC > 0): the sample loads are issued BEFORE anything that waits for the per-ray constants (direction norm, NDC
  // terms).  In program order they used to follow ray_consts(), and the fast / slow branches around every run kept the scheduler from
  // hoisting them over its square root: two dependent memory round trips per ray instead of one (128-sample rays: 0.73 -> 0.81-0.83
  // of the copy bandwidth; the same reordering in the backward spilled under its register cap and lost 10-20 %: not done there).
  auto load = [&](int c, float (&sg)[L], float (&zi)[L], float (&col)[3 * L], float& zn_raw) {
    const int i0 = c * 32 * L + lane * L - lead;              // local index of the run's first element
    const long long e0 = first + i0;
    const bool fast = vec && i0 >= 0 && i0 + L <= S;
    load_run<L, 1>(p.sigma + e0, fast, i0, S, sg);
    load_run<L, 1>(p.z + e0, fast, i0, S, zi);
    if (has_rgb) load_run<3 * L, 3>(p.rgb + e0 * 3, fast, i0, S, col);
    zn_raw = (lane == 31 && i0 + L < S) ? __ldg(p.z + e0 + L) : far_z;      // the depth behind the warp's last run
  };
  auto compute = [&](int c, const RayConsts& k, const float (&sg)[L], const float (&zi)[L], const float (&col)[3 * L], float zn_raw,
                     float (&w)[L], float (&zw)[L]) {
    const int i0 = c * 32 * L + lane * L - lead;
    const long long e0 = first + i0;
    const bool fast = vec && i0 >= 0 && i0 + L <= S;
    const float ns = k.nrm * p.distance_scale;
    float znext = __shfl_down_sync(FULL, zi[0], 1);
    if (lane == 31) znext = zn_raw;
    float q[L], al[L];
    float P = 1.f;
#pragma unroll
    for (int j = 0; j < L; ++j) {
      const int i = i0 + j;
      const bool ok = (unsigned)i < (unsigned)S;
      float zn = j < L - 1 ? zi[j < L - 1 ? j + 1 : j] : znext;
      if (i + 1 >= S) zn = far_z;
      const float ex = __expf(-sg[j] * ((zn - zi[j]) * ns));
      al[j] = ok ? 1.f - ex : 0.f;
      q[j] = ok ? (1.f - al[j] + 1e-10f) : 1.f;
      P *= q[j];
    }
    const float incl = warp_incl_prod(P, lane);
    float excl = __shfl_up_sync(FULL, incl, 1);
    if (lane == 0) excl = 1.f;
    float T = carry * excl;
    carry *= __shfl_sync(FULL, incl, 31);
    float vis[L];
#pragma unroll
    for (int j = 0; j < L; ++j) {
      vis[j] = T;
      w[j] = al[j] * T;                 // 0 for masked elements (al = 0)
      T *= q[j];
      red[0] += w[j];
      red[1] += w[j] * zi[j];
      zw[j] = 0.f;
      if (p.ndc) {
        zw[j] = ndc_to_world(zi[j], k.A, k.tn);
        red[2] += w[j] * zw[j];
      }
      if (has_rgb) {
        red[3] += w[j] * col[3 * j + 0];
        red[4] += w[j] * col[3 * j + 1];
        red[5] += w[j] * col[3 * j + 2];
      }
    }
    store_run<L, 1>(p.weights + e0, fast, i0, S, w);
    if (p.alpha) store_run<L, 1>(p.alpha + e0, fast, i0, S, al);
    if (p.visibility) store_run<L, 1>(p.visibility + e0, fast, i0, S, vis);
  };

  RayConsts k;
  if (NC > 0) {               // unconditional: a step beyond the ray is fully masked (no loads, no stores)
#pragma unroll
    for (int c = 0; c < KEEP; ++c) load(c, sgk[c], zk[c], colk[c], znk[c]);
    k = ray_consts(p, r);
#pragma unroll
    for (int c = 0; c < KEEP; ++c) compute(c, k, sgk[c], zk[c], colk[c], znk[c], wk[c], zwk[c]);
  } else {
    // several steps per ray: the constants first (keeping a step's samples live across them cost 5-14 registers and 2-3 % here)
    k = ray_consts(p, r);
    for (int c = 0; c < nchunks; ++c) {
      load(c, sgk[0], zk[0], colk[0], znk[0]);
      compute(c, k, sgk[0], zk[0], colk[0], znk[0], wk[0], zwk[0]);
    }
  }

  const float mine = warp_reduce8(red, lane);
  const float acc = __shfl_sync(FULL, mine, 0), nz = __shfl_sync(FULL, mine, 4), nzw = __shfl_sync(FULL, mine, 8);
  float c0 = __shfl_sync(FULL, mine, 12), c1 = __shfl_sync(FULL, mine, 16), c2 = __shfl_sync(FULL, mine, 20);
  const float inv = rcp_approx(acc + 1e-6f);
  const float d_main = nz * inv;          // NDC depth when ndc, world depth otherwise
  const float d_world = p.ndc ? nzw * inv : d_main;
  // variances around the means (each lane handles exactly the runs it produced)
  float v_main = 0.f, v_world = 0.f;
  if (NC > 0) {
#pragma unroll
    for (int c = 0; c < KEEP; ++c)
#pragma unroll
      for (int j = 0; j < L; ++j) {
        const float a = zk[c][j] - d_main;
        v_main += wk[c][j] * a * a;
        if (p.ndc) {
          const float b = zwk[c][j] - d_world;
          v_world += wk[c][j] * b * b;
        }
      }
  } else {
    for (int c = 0; c < nchunks; ++c) {
      const int i0 = c * 32 * L + lane * L - lead;
#pragma unroll
      for (int j = 0; j < L; ++j) {            // plain loads: the weights were written by this thread in this kernel
        const int i = i0 + j;
        const bool ok = (unsigned)i < (unsigned)S;
        const float w = ok ? p.weights[first + i] : 0.f;
        const float zi = ok ? __ldg(p.z + first + i) : 0.f;
        const float a = zi - d_main;
        v_main += w * a * a;
        if (p.ndc) {
          const float b = ndc_to_world(zi, k.A, k.tn) - d_world;
          v_world += w * b * b;
        }
      }
    }
  }
  const float vv = warp_reduce2(v_main, v_world, lane);
  v_main = __shfl_sync(FULL, vv, 0);
  v_world = __shfl_sync(FULL, vv, 16);
  if (lane == 0) {
    p.acc[r] = acc;
    if (p.ndc) {
      p.depth_ndc[r] = d_main; p.depth_var_ndc[r] = v_main; p.depth[r] = d_world; p.depth_var[r] = v_world;
    } else {
      p.depth[r] = d_main; p.depth_var[r] = v_main;
    }
    if (p.rgb_map != nullptr) {
      if (p.white_bkgd) { const float bg = 1.f - acc; c0 += bg; c1 += bg; c2 += bg; }
      p.rgb_map[r * 3 + 0] = c0; p.rgb_map[r * 3 + 1] = c1; p.rgb_map[r * 3 + 2] = c2;
    }
  }
}

// LOOP: a warp walks rays r, r + (warps of the grid), ... under a grid capped by the host at CMP_FWD_WAVES CTAs per SM, so that a launch over
// millions of short rays is not paced by the block scheduler's CTA turnover (64-sample rays: 0.58 -> 0.67-0.69 of the copy bandwidth, two-step
// 256-sample rays 0.67 -> 0.71).  The single-step 4-sample-run kernel (128-sample rays) keeps one warp per ray: inside the loop it lost
// the early issue of its loads (0.82 -> 0.73).
template <int L, int NC, bool LOOP>
__global__ void __launch_bounds__(CMP_WARPS * 32) composite_fwd_kernel(CompositeFwd p) {
  const int warp = threadIdx.x >> 5, lane = lane_id();
  const long long r0 = (long long)blockIdx.x * CMP_WARPS + warp;
  if constexpr (LOOP) {
    for (long long r = r0; r < p.R; r += (long long)gridDim.x * CMP_WARPS) composite_fwd_ray<L, NC>(p, r, lane);
  } else {
    if (r0 < p.R) composite_fwd_ray<L, NC>(p, r0, lane);
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
  int vec;
};

// Backward: one reverse pass with the same run layout; sum_{k>i} g_w[k] w_k is a sequential suffix sum inside the run
// plus one shuffle suffix scan per step (closed form in oracle/composite.py).
#ifndef SRF_CMP_BWD_MINB2
#define SRF_CMP_BWD_MINB2 5
#endif
#ifndef SRF_CMP_BWD_MINB4
#define SRF_CMP_BWD_MINB4 4
#endif
template <int L>
__device__ __forceinline__ void composite_bwd_ray(const CompositeBwd& p, long long r, int lane) {
  constexpr int G = L % 4 == 0 ? 4 : 2;
  const int S = p.S;
  const long long first = r * S;
  const int lead = (int)(first & (G - 1));
  const int nchunks = (lead + S + 32 * L - 1) / (32 * L);
  const bool vec = p.vec != 0;
  const bool has_rgb = p.rgb != nullptr && p.g_rgb != nullptr;
  float sg[L], zi[L], T[L], col[3 * L], gwt[L], zn_raw;
  auto load = [&](int c) {
    const int i0 = c * 32 * L + lane * L - lead;
    const long long e0 = first + i0;
    const bool fast = vec && i0 >= 0 && i0 + L <= S;
    load_run<L, 1>(p.sigma + e0, fast, i0, S, sg);
    load_run<L, 1>(p.z + e0, fast, i0, S, zi);
    load_run<L, 1>(p.visibility + e0, fast, i0, S, T);
    if (has_rgb) load_run<3 * L, 3>(p.rgb + e0 * 3, fast, i0, S, col);
    if (p.g_weights) load_run<L, 1>(p.g_weights + e0, fast, i0, S, gwt);
    // the depth behind the warp's last run: requested together with the runs, not after the shuffle that waits for them
    zn_raw = (lane == 31 && i0 + L < S) ? __ldg(p.z + e0 + L) : (p.ndc ? 1.f : 1e10f);
  };
  // (issuing the first step's loads here, ahead of the per-ray scalars below - what the forward does - measured 4-12 % slower with every
  // register cap tried: the backward keeps 12 per-ray scalars and five arrays per step live)
  const RayConsts k = ray_consts(p, r);
  const float ns = k.nrm * p.distance_scale;

  const float acc = p.acc[r];
  const float invA = rcp_approx(acc + 1e-6f);
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
  const bool depth_main = gd_ndc != 0.f || gv_ndc != 0.f;                 // warp-uniform: skip unused depth terms
  const bool depth_world = gd_world != 0.f || gv_world != 0.f;

  float carry = 0.f;    // sum over the samples of later steps of g_w[k] w_k
  for (int c = nchunks - 1; c >= 0; --c) {
    const int i0 = c * 32 * L + lane * L - lead;
    const long long e0 = first + i0;
    const bool fast = vec && i0 >= 0 && i0 + L <= S;
    load(c);
    float znext = __shfl_down_sync(FULL, zi[0], 1);
    if (lane == 31) znext = zn_raw;
    float gw[L], gww[L], w[L], q[L], de[L];
#pragma unroll
    for (int j = 0; j < L; ++j) {
      const int i = i0 + j;
      const bool ok = (unsigned)i < (unsigned)S;
      float zn = j < L - 1 ? zi[j < L - 1 ? j + 1 : j] : znext;
      if (i + 1 >= S) zn = k.far_z;
      const float dscale = (zn - zi[j]) * ns;
      const float ex = ok ? __expf(-sg[j] * dscale) : 1.f;
      const float al = 1.f - ex;
      q[j] = al == 1.f ? 1e-10f : (1.f - al + 1e-10f);
      w[j] = ok ? al * T[j] : 0.f;
      de[j] = dscale * ex;
      float g = g_acc;
      if (p.g_weights) g += gwt[j];
      if (has_rgb) g += gr0 * col[3 * j + 0] + gr1 * col[3 * j + 1] + gr2 * col[3 * j + 2];
      if (p.ndc) {
        if (depth_main) {
          const float a = zi[j] - d_ndc;
          g += gd_ndc * a * invA + gv_ndc * (a * a - kv_ndc * a);
        }
        if (depth_world) {
          const float b = ndc_to_world(zi[j], k.A, k.tn) - d_world;
          g += gd_world * b * invA + gv_world * (b * b - kv_world * b);
        }
      } else if (depth_world) {
        const float b = zi[j] - d_world;
        g += gd_world * b * invA + gv_world * (b * b - kv_world * b);
      }
      gw[j] = g;
      gww[j] = ok ? g * w[j] : 0.f;
    }
    // suffix sums: after[j] = sum of gww over the later elements of this run
    float after[L];
    after[L - 1] = 0.f;
#pragma unroll
    for (int j = L - 2; j >= 0; --j) after[j] = after[j + 1] + gww[j + 1];
    const float tot = after[0] + gww[0];
    float suf = tot;      // inclusive suffix sum over lanes (towards higher lanes)
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const float t = __shfl_down_sync(FULL, suf, o);
      if (lane + o < 32) suf += t;
    }
    const float base_later = carry + (suf - tot);
    carry += __shfl_sync(FULL, suf, 0);
    float gs[L];
#pragma unroll
    for (int j = 0; j < L; ++j) {
      const float later = base_later + after[j];
      const float g_alpha = gw[j] * T[j] - __fdividef(later, q[j]);
      gs[j] = g_alpha * de[j];
    }
    store_run<L, 1>(p.g_sigma + e0, fast, i0, S, gs);
    if (p.g_rgb_s != nullptr) {
      float gc[3 * L];
#pragma unroll
      for (int j = 0; j < L; ++j) { gc[3 * j] = w[j] * gr0; gc[3 * j + 1] = w[j] * gr1; gc[3 * j + 2] = w[j] * gr2; }
      store_run<3 * L, 3>(p.g_rgb_s + e0 * 3, fast, i0, S, gc);
    }
  }
}

// (one warp per ray: walking several rays per warp under a capped grid, which helps the forward, measured 3-12 % slower here)
template <int L>
__global__ void __launch_bounds__(CMP_WARPS * 32, L == 2 ? SRF_CMP_BWD_MINB2 : SRF_CMP_BWD_MINB4) composite_bwd_kernel(CompositeBwd p) {
  const int warp = threadIdx.x >> 5, lane = lane_id();
  const long long r = (long long)blockIdx.x * CMP_WARPS + warp;
  if (r < p.R) composite_bwd_ray<L>(p, r, lane);
}

}  // namespace srf

using namespace srf;

// samples per lane and step: 2 (64 per warp step, 64-bit pieces) or 4 (128 per step, 128-bit pieces) - whichever
// covers the ray with fewer idle lane slots; longer runs win unless they idle more than 15 % extra (fewer scans).
// Every piece a lane loads or stores is contiguous with its neighbours', so all accesses are fully coalesced.
#ifndef SRF_CMP_L4_PCT
#define SRF_CMP_L4_PCT 115
#endif
constexpr unsigned CMP_FWD_WAVES = 32;     // CTAs per SM in the grid of the looping forward kernels (8 are resident at a time)
static int run_length(int S) {
  const long long s2 = (long long)((S + (S % 2) + 63) / 64) * 64;
  const long long s4 = (long long)((S + (S % 4 ? 3 : 0) + 127) / 128) * 128;
  return s4 * 100 <= s2 * SRF_CMP_L4_PCT ? 4 : 2;
}

template <typename... Ts>
static int aligned16(const Ts*... ptrs) {
  uintptr_t bits = 0;
  for (const void* q : {static_cast<const void*>(ptrs)...}) bits |= reinterpret_cast<uintptr_t>(q);     // nullptr contributes nothing
  return (bits & 15) == 0 ? 1 : 0;
}

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
  const int vec = aligned16(sigma, rgb, z, alpha, visibility, weights);
  CompositeFwd p{sigma, rgb, z, rays_o, rays_d, rays_d_ndc, alpha, visibility, weights, rgb_map, acc, depth,
                 depth_var, depth_ndc, depth_var_ndc, num_rays, num_samples, ndc, white_bkgd, distance_scale, vec};
  const unsigned blocks = (unsigned)((num_rays + CMP_WARPS - 1) / CMP_WARPS);
  const unsigned cap = (unsigned)sm_count() * CMP_FWD_WAVES;
  const unsigned looped = blocks < cap ? blocks : cap;
  const dim3 blk(CMP_WARPS * 32);
  const cudaStream_t st = (cudaStream_t)stream;
  const int L = run_length(num_samples);
  const int G = L == 4 ? 4 : 2;
  const int span = num_samples + (num_samples % G != 0 ? G - 1 : 0);       // worst case with the leading partial run
  const int steps = (span + 32 * L - 1) / (32 * L);
  if (L == 2) {
    if (steps <= 1) composite_fwd_kernel<2, 1, true><<<looped, blk, 0, st>>>(p);
    else composite_fwd_kernel<2, 0, true><<<looped, blk, 0, st>>>(p);
  } else {
    if (steps <= 1) composite_fwd_kernel<4, 1, false><<<blocks, blk, 0, st>>>(p);
    else composite_fwd_kernel<4, 0, true><<<looped, blk, 0, st>>>(p);
  }
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
                 ndc, white_bkgd, distance_scale, aligned16(sigma, rgb, z, visibility, g_weights, g_sigma, g_rgb_samples)};
  const unsigned blocks = (unsigned)((num_rays + CMP_WARPS - 1) / CMP_WARPS);
  if (run_length(num_samples) == 2) composite_bwd_kernel<2><<<blocks, CMP_WARPS * 32, 0, (cudaStream_t)stream>>>(p);
  else composite_bwd_kernel<4><<<blocks, CMP_WARPS * 32, 0, (cudaStream_t)stream>>>(p);
  return check_launch("srf_composite_bwd");
}
