// Simple-TensoRF vector-matrix (VM) tensor: occupancy test, sample compaction, density / appearance gathers.
//
// Replaces (reference file:line, relative to the upstream checkout):
//   srf_pack_alpha_bits        derived cache of AlphaGridMask.alpha_volume (src/models/SimpleTensoRF09.py:1323-1345)
//   srf_tensorf_mask           :263 (pts = o + d z), :705 (box test), :707-710 + :1342-1349 (alphaMask test)
//   srf_threshold_mask         :726 (weights > ray_marching_weight_threshold)
//   srf_compact                the boolean-mask indexing of :1221 / :1248 (stable row-major order) without the host sync
//   srf_vm_density_fwd / _bwd  :763-765 (normalise), :1214-1239 (get_volume_density) and its autograd
//   srf_vm_color_features_fwd / _bwd   :1241-1263 (plane x line products, basis_matrix_color) and its autograd
//
// Planes and lines are read from channels-last derived caches ([H][W][C] / [L][C]) so that one texel is one or
// a few 16-byte vectors; the fp32 `nn.Parameter`s keep the reference's [1,C,H,W] layout.  The occupancy volume
// is 1 bit per voxel (x fastest), i.e. L2/shared-memory sized, and the test reproduces ATen's grid_sampler_3d
// coordinate arithmetic exactly, so mask and compaction order are bit-identical to the reference.
#include <cuda_bf16.h>
#include <limits.h>
#include <stdlib.h>

#include "tensorf_common.cuh"

namespace srf {

#ifndef SRF_VM_CFWD_RUNS_DEFAULT
#define SRF_VM_CFWD_RUNS_DEFAULT 8
#endif
constexpr int CMP_BLOCK = 1024;     // elements per compaction block (256 threads x 4)

// ------------------------------------------------------------------------------------------ occupancy bits
__global__ void pack_alpha_kernel(const float* __restrict__ vol, long long n, uint32_t* __restrict__ bits) {
  const long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long base = w * 32;
  if (base >= n) return;
  uint32_t word = 0;
  for (int b = 0; b < 32; ++b) {
    const long long i = base + b;
    if (i < n && vol[i] > 0.f) word |= 1u << b;
  }
  bits[w] = word;
}

// thread = 4 consecutive samples (normally of one ray): one 16-byte depth load, one 4-byte mask store, ray origin and
// direction fetched once; the only 64-bit division is one per thread
__global__ void __launch_bounds__(256) tensorf_mask_kernel(MaskParams p) {
  const long long i0 = (long long)blockIdx.x * CMP_BLOCK + threadIdx.x * 4;
  int local = 0;
  if (i0 < p.total) {
    const int nvalid = (int)min(4LL, p.total - i0);
    float zz[4];
    if (nvalid == 4 && (reinterpret_cast<uintptr_t>(p.z + i0) & 15) == 0) {
      const float4 z4 = __ldg(reinterpret_cast<const float4*>(p.z + i0));
      zz[0] = z4.x; zz[1] = z4.y; zz[2] = z4.z; zz[3] = z4.w;
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k) zz[k] = k < nvalid ? p.z[i0 + k] : 0.f;
    }
    long long r = i0 / p.S;
    int rem = (int)(i0 - r * p.S);
    long long loaded = -1;
    float o[3] = {0.f, 0.f, 0.f}, d[3] = {0.f, 0.f, 0.f};
    uint32_t packed = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (k < nvalid) {
        while (rem >= p.S) { rem -= p.S; ++r; }
        if (r != loaded) {
#pragma unroll
          for (int a = 0; a < 3; ++a) { o[a] = __ldg(p.rays_o + r * 3 + a); d[a] = __ldg(p.rays_d + r * 3 + a); }
          loaded = r;
        }
        float pt[3];
        bool ok = true;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
          pt[a] = __fadd_rn(o[a], __fmul_rn(d[a], zz[k]));
          ok = ok && (p.bb0[a] <= pt[a]) && (pt[a] <= p.bb1[a]);
        }
        if (ok && p.alpha_bits != nullptr) ok = alpha_hit(p, pt);
        packed |= (ok ? 1u : 0u) << (8 * k);
        local += ok ? 1 : 0;
        ++rem;
      }
    }
    if (nvalid == 4 && (reinterpret_cast<uintptr_t>(p.mask + i0) & 3) == 0) {
      *reinterpret_cast<uint32_t*>(p.mask + i0) = packed;
    } else {
      for (int k = 0; k < nvalid; ++k) p.mask[i0 + k] = (uint8_t)((packed >> (8 * k)) & 1u);
    }
  }
  __shared__ int s_cnt[8];
  int w = local;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) w += __shfl_xor_sync(FULL, w, o);
  if ((threadIdx.x & 31) == 0) s_cnt[threadIdx.x >> 5] = w;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int k = 0; k < 8; ++k) t += s_cnt[k];
    p.block_counts[blockIdx.x] = t;
  }
}

__global__ void __launch_bounds__(256) threshold_mask_kernel(const float* __restrict__ v, float thr, long long total,
                                                             uint8_t* __restrict__ mask, int* __restrict__ block_counts) {
  const long long base = (long long)blockIdx.x * CMP_BLOCK;
  int local = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const long long i = base + k * 256 + threadIdx.x;
    bool ok = false;
    if (i < total) { ok = v[i] > thr; mask[i] = ok ? 1 : 0; }
    local += ok ? 1 : 0;
  }
  __shared__ int s_cnt[8];
  int w = local;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) w += __shfl_xor_sync(FULL, w, o);
  if ((threadIdx.x & 31) == 0) s_cnt[threadIdx.x >> 5] = w;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int k = 0; k < 8; ++k) t += s_cnt[k];
    block_counts[blockIdx.x] = t;
  }
}

// exclusive scan of the per-block counts (one CTA; nb is a few thousand at most), total -> count[0]
__global__ void __launch_bounds__(1024) scan_blocks_kernel(const int* __restrict__ counts, int nb, int* __restrict__ offsets,
                                                           int* __restrict__ count) {
  __shared__ int s_warp[32];
  __shared__ int s_carry;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  for (int base = 0; base < nb; base += 1024) {
    const int i = base + threadIdx.x;
    const int v = i < nb ? counts[i] : 0;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(FULL, incl, o);
      if ((threadIdx.x & 31) >= o) incl += t;
    }
    if ((threadIdx.x & 31) == 31) s_warp[threadIdx.x >> 5] = incl;
    __syncthreads();
    if (threadIdx.x < 32) {
      int w = s_warp[threadIdx.x];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(FULL, w, o);
        if (threadIdx.x >= o) w += t;
      }
      s_warp[threadIdx.x] = w;
    }
    __syncthreads();
    const int warp_off = (threadIdx.x >> 5) == 0 ? 0 : s_warp[(threadIdx.x >> 5) - 1];
    const int carry = s_carry;
    if (i < nb) offsets[i] = carry + warp_off + incl - v;
    __syncthreads();
    if (threadIdx.x == 1023) s_carry = carry + warp_off + incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) count[0] = s_carry;
}

// stable scatter: thread t of a block owns the 4 consecutive elements base + 4t .. 4t + 3 (one 32-bit mask load), the
// block-wide exclusive scan of the per-thread counts is one shuffle scan per warp plus an 8-entry exchange; blocks whose
// count is zero (most of them behind a sparse alpha mask) exit after one load
__global__ void __launch_bounds__(256) compact_scatter_kernel(const uint8_t* __restrict__ mask, long long total,
                                                              const int* __restrict__ counts, const int* __restrict__ offsets,
                                                              int* __restrict__ idx) {
  if (counts[blockIdx.x] == 0) return;
  __shared__ int s_warp[8];
  const long long i0 = (long long)blockIdx.x * CMP_BLOCK + threadIdx.x * 4;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t m = 0;
  if (i0 + 4 <= total && (reinterpret_cast<uintptr_t>(mask + i0) & 3) == 0) {
    m = *reinterpret_cast<const uint32_t*>(mask + i0);
  } else {
    for (int k = 0; k < 4; ++k)
      if (i0 + k < total) m |= (uint32_t)mask[i0 + k] << (8 * k);
  }
  const int f0 = (m & 0xffu) != 0, f1 = (m & 0xff00u) != 0, f2 = (m & 0xff0000u) != 0, f3 = (m & 0xff000000u) != 0;
  const int mine = f0 + f1 + f2 + f3;
  int incl = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(FULL, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  int before = offsets[blockIdx.x] + incl - mine;
#pragma unroll
  for (int w = 0; w < 8; ++w)
    if (w < warp) before += s_warp[w];
  if (f0) idx[before++] = (int)i0;
  if (f1) idx[before++] = (int)(i0 + 1);
  if (f2) idx[before++] = (int)(i0 + 2);
  if (f3) idx[before] = (int)(i0 + 3);
}

// density: sigma = act(sum_i sum_c plane_i,c * line_i,c); one thread per compacted sample
__global__ void __launch_bounds__(256) vm_density_fwd_kernel(VmGeom g, VmGrid t, int softplus, float offset,
                                                             float* __restrict__ sigma, float* __restrict__ feat_out) {
  const int n = g.count[0];
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
    const int flat = g.idx[j];
    float pn[3];
    normalized_point(g, flat, pn);
    float feat = 0.f;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const Bilerp b = plane_coords(pn, t.res, i);
      int l0, L; float w0, w1;
      line_coords(pn, t.res, i, l0, L, w0, w1);
      float part = 0.f;
      for (int c = 0; c < t.C[i]; c += 4) {
        const float4 pv = plane_fetch4(t.plane[i], b, t.C[i], c);
        const float4 lv = line_fetch4(t.line[i], l0, L, w0, w1, t.C[i], c);
        part += pv.x * lv.x + pv.y * lv.y + pv.z * lv.z + pv.w * lv.w;
      }
      feat += part;
    }
    if (feat_out) feat_out[j] = feat;
    float s;
    if (softplus) { const float x = feat + offset; s = x > 20.f ? x : log1pf(expf(x)); }
    else s = fmaxf(feat, 0.f);
    sigma[flat] = s;
  }
}

// backward: d feat -> bilinear/linear scatter-add into the channels-last gradient buffers
__global__ void __launch_bounds__(256) vm_density_bwd_kernel(VmGeom g, VmGrid t, int softplus, float offset,
                                                             const float* __restrict__ g_sigma, const float* __restrict__ feat_in,
                                                             float* gp0, float* gp1, float* gp2, float* gl0, float* gl1, float* gl2) {
  const int n = g.count[0];
  float* gplane[3] = {gp0, gp1, gp2};
  float* gline[3] = {gl0, gl1, gl2};
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
    const int flat = g.idx[j];
    const float feat = feat_in[j];
    float gf = g_sigma[flat];
    if (softplus) { const float x = feat + offset; gf *= 1.f / (1.f + expf(-x)); }
    else gf = feat > 0.f ? gf : 0.f;
    if (gf == 0.f) continue;
    float pn[3];
    normalized_point(g, flat, pn);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const Bilerp b = plane_coords(pn, t.res, i);
      int l0, L; float w0, w1;
      line_coords(pn, t.res, i, l0, L, w0, w1);
      const int C = t.C[i];
      for (int c = 0; c < C; c += 4) {
        const float4 pv = plane_fetch4(t.plane[i], b, C, c);
        const float4 lv = line_fetch4(t.line[i], l0, L, w0, w1, C, c);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int x = b.x0 + (k & 1), y = b.y0 + (k >> 1);
          if (x < 0 || y < 0 || x >= b.W || y >= b.H) continue;
          const float w = gf * ((k & 1) ? b.wx1 : b.wx0) * ((k >> 1) ? b.wy1 : b.wy0);
          red_add4(gplane[i] + ((size_t)y * b.W + x) * C + c, w * lv.x, w * lv.y, w * lv.z, w * lv.w);
        }
        if (l0 >= 0 && l0 < L) red_add4(gline[i] + (size_t)l0 * C + c, gf * w0 * pv.x, gf * w0 * pv.y, gf * w0 * pv.z, gf * w0 * pv.w);
        if (l0 + 1 >= 0 && l0 + 1 < L)
          red_add4(gline[i] + (size_t)(l0 + 1) * C + c, gf * w1 * pv.x, gf * w1 * pv.y, gf * w1 * pv.z, gf * w1 * pv.w);
      }
    }
  }
}

// ---- run-merged backward.  The per-SM issue rate of global reductions is the bound of the scatter (REDG: ~1.3 cycles per
// lane and 32-bit value, i.e. ~165 cycles per warp-wide red.v4 — the six float4 loads and the arithmetic of a sample are noise
// next to its six red.v4), so the lever is the NUMBER of reductions.  The compacted list is ray-major and consecutive samples of a
// ray are half a voxel apart, so consecutive samples mostly share their bilinear footprint: in NDC the rays run along z, the
// (x, y) plane — 16 of the 24 density channels — keeps the same 2x2 texels for 5-20 samples in a row.  A thread therefore owns
// a RUN of K consecutive compacted samples and, per (plane, channel group), accumulates the four corner gradients (and the two
// line gradients) in registers for as long as the footprint stays the same; it emits one red.v4 per corner per footprint
// instead of one per corner per sample, and re-uses the loaded texels across the streak as well.
struct Footprint {
  int x0, y0, l0;
  float wx0, wx1, wy0, wy1, wl0, wl1;       // the forward's weights, bit for bit: w1 = i - floor(i), w0 = (floor(i) + 1) - i
};

__device__ __forceinline__ Footprint footprint_of(const float (&pn)[3], const int (&res)[3], int i) {
  Footprint f;
  const int W = res[axis0(i)], H = res[axis1(i)], L = res[axisv(i)];
  const float ix = unnorm(pn[axis0(i)], W), iy = unnorm(pn[axis1(i)], H), il = unnorm(pn[axisv(i)], L);
  const float fx = floorf(ix), fy = floorf(iy), fl = floorf(il);
  f.x0 = (int)fx; f.y0 = (int)fy; f.l0 = (int)fl;
  f.wx1 = ix - fx; f.wy1 = iy - fy; f.wl1 = il - fl;
  f.wx0 = (fx + 1.f) - ix; f.wy0 = (fy + 1.f) - iy; f.wl0 = (fl + 1.f) - il;
  return f;
}

__device__ __forceinline__ float4 f4_zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ void f4_fma(float4& a, float s, const float4& b) { a.x += s * b.x; a.y += s * b.y; a.z += s * b.z; a.w += s * b.w; }
__device__ __forceinline__ bool f4_any(const float4& a) { return a.x != 0.f || a.y != 0.f || a.z != 0.f || a.w != 0.f; }

template <int K>
__global__ void __launch_bounds__(128) vm_density_bwd_runs_kernel(VmGeom g, VmGrid t, int softplus, float offset,
                                                                  const float* __restrict__ g_sigma, const float* __restrict__ feat_in,
                                                                  float* gp0, float* gp1, float* gp2, float* gl0, float* gl1, float* gl2) {
  const int n = g.count[0];
  const int nruns = (n + K - 1) / K;
  for (int run = blockIdx.x * blockDim.x + threadIdx.x; run < nruns; run += gridDim.x * blockDim.x) {
    const int j0 = run * K;
    float pn[K][3], gf[K];
    bool any = false;
#pragma unroll
    for (int k = 0; k < K; ++k) {
      gf[k] = 0.f;
      pn[k][0] = pn[k][1] = pn[k][2] = 0.f;
      if (j0 + k < n) {
        const int flat = g.idx[j0 + k];
        const float feat = feat_in[j0 + k];
        float v = g_sigma[flat];
        if (softplus) { const float x = feat + offset; v *= 1.f / (1.f + expf(-x)); }
        else v = feat > 0.f ? v : 0.f;
        gf[k] = v;
        if (v != 0.f) { normalized_point(g, flat, pn[k]); any = true; }
      }
    }
    if (!any) continue;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const int C = t.C[i];
      const int W = t.res[axis0(i)], H = t.res[axis1(i)], L = t.res[axisv(i)];
      const float* plane = i == 0 ? t.plane[0] : (i == 1 ? t.plane[1] : t.plane[2]);
      const float* line = i == 0 ? t.line[0] : (i == 1 ? t.line[1] : t.line[2]);
      float* gplane = i == 0 ? gp0 : (i == 1 ? gp1 : gp2);
      float* gline = i == 0 ? gl0 : (i == 1 ? gl1 : gl2);
      for (int c = 0; c < C; c += 4) {
        int kx = INT_MIN, ky = INT_MIN, kl = INT_MIN;             // footprint the accumulators belong to
        float4 tex[4], acc[4], ltex[2], lacc[2];
#pragma unroll
        for (int q = 0; q < 4; ++q) { tex[q] = f4_zero(); acc[q] = f4_zero(); }
        ltex[0] = ltex[1] = lacc[0] = lacc[1] = f4_zero();
#pragma unroll
        for (int k = 0; k <= K; ++k) {
          const bool live = k < K && gf[k < K ? k : 0] != 0.f;
          Footprint f;
          if (live) f = footprint_of(pn[k < K ? k : 0], t.res, i);
          const bool last = k == K;
          if (!live && !last) continue;
          if (last || f.x0 != kx || f.y0 != ky) {                // plane footprint changes: flush, then load the new texels
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const int x = kx + (q & 1), y = ky + (q >> 1);
              if (kx != INT_MIN && x >= 0 && y >= 0 && x < W && y < H && f4_any(acc[q]))
                red_add4(gplane + ((size_t)y * W + x) * C + c, acc[q].x, acc[q].y, acc[q].z, acc[q].w);
              acc[q] = f4_zero();
            }
            if (!last) {
              kx = f.x0; ky = f.y0;
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const int x = kx + (q & 1), y = ky + (q >> 1);
                tex[q] = (x >= 0 && y >= 0 && x < W && y < H) ? ldg4(plane + ((size_t)y * W + x) * C + c) : f4_zero();
              }
            }
          }
          if (last || f.l0 != kl) {
#pragma unroll
            for (int q = 0; q < 2; ++q) {
              const int l = kl + q;
              if (kl != INT_MIN && l >= 0 && l < L && f4_any(lacc[q])) red_add4(gline + (size_t)l * C + c, lacc[q].x, lacc[q].y, lacc[q].z, lacc[q].w);
              lacc[q] = f4_zero();
            }
            if (!last) {
              kl = f.l0;
#pragma unroll
              for (int q = 0; q < 2; ++q) {
                const int l = kl + q;
                ltex[q] = (l >= 0 && l < L) ? ldg4(line + (size_t)l * C + c) : f4_zero();
              }
            }
          }
          if (last) continue;
          const float gk = gf[k < K ? k : 0];
          const float wl0 = f.wl0;
          const float w00 = f.wx0 * f.wy0, w10 = f.wx1 * f.wy0, w01 = f.wx0 * f.wy1, w11 = f.wx1 * f.wy1;
          float4 pv = f4_zero(), lv = f4_zero();
          f4_fma(pv, w00, tex[0]); f4_fma(pv, w10, tex[1]); f4_fma(pv, w01, tex[2]); f4_fma(pv, w11, tex[3]);
          f4_fma(lv, wl0, ltex[0]); f4_fma(lv, f.wl1, ltex[1]);
          f4_fma(acc[0], gk * w00, lv); f4_fma(acc[1], gk * w10, lv); f4_fma(acc[2], gk * w01, lv); f4_fma(acc[3], gk * w11, lv);
          f4_fma(lacc[0], gk * wl0, pv); f4_fma(lacc[1], gk * f.wl1, pv);
        }
      }
    }
  }
}

// appearance: products (plane x line) over all channels -> bf16 rows of `pitch` elements
// [products (CT = sum C) | view_dirs (3) | zero pad]: the A operand of the tensor-core colour MLP, whose first layer
// absorbs basis_matrix_color (W0' = [W0[:, :F] B | W0[:, F:]]), so no matrix-vector product is left in this kernel.
// A warp takes 32 samples.  Phase 1, lane per sample: coordinates -> per plane 4 clamped corner offsets + weights
// (zero weight = grid_sample's zero padding) and per line 2 offsets + weights, into shared memory.  Phase 2, lane per
// (sample, 4-channel group): 6 float4 texel loads, 28 FMAs, one 8-byte store; consecutive lanes write consecutive
// 8-byte pieces of the row-major output, so stores are fully coalesced and no lane idles on a short channel list.
constexpr int CF_WARPS = 8;
constexpr int CB_WARPS = 4, CB_SAMPLES = 64;       // run-merged kernels: a warp takes 64 consecutive samples

struct alignas(16) SampleRec {
  int poff[3][4];          // element offsets of the bilinear corners in the channels-last plane
  float pw[3][4];
  int line[3][4];          // {offset 0, offset 1, weight 0 bits, weight 1 bits}
};

__device__ __forceinline__ int clampi(int v, int hi) { return v < 0 ? 0 : (v >= hi ? hi - 1 : v); }

__device__ __forceinline__ void compute_record(const VmGeom& g, const VmGrid& t, int flat, SampleRec& r) {
  float pn[3];
  normalized_point(g, flat, pn);
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const Bilerp b = plane_coords(pn, t.res, i);
    const int C = t.C[i];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int x = b.x0 + (k & 1), y = b.y0 + (k >> 1);
      const bool in = x >= 0 && y >= 0 && x < b.W && y < b.H;
      r.pw[i][k] = in ? ((k & 1) ? b.wx1 : b.wx0) * ((k >> 1) ? b.wy1 : b.wy0) : 0.f;
      r.poff[i][k] = (clampi(y, b.H) * b.W + clampi(x, b.W)) * C;
    }
    int l0, L; float w0, w1;
    line_coords(pn, t.res, i, l0, L, w0, w1);
    r.line[i][0] = clampi(l0, L) * C;
    r.line[i][1] = clampi(l0 + 1, L) * C;
    r.line[i][2] = __float_as_int((l0 >= 0 && l0 < L) ? w0 : 0.f);
    r.line[i][3] = __float_as_int((l0 + 1 >= 0 && l0 + 1 < L) ? w1 : 0.f);
  }
}

struct GroupMap { int g0, g1, G, GP; unsigned inv; };     // float4 groups: [0,g0) plane 0, [g0,g1) plane 1, [g1,G) plane 2

__device__ __forceinline__ uint32_t ptx_pack_bf16(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}


// plane and line values of one 4-channel group at one sample
__device__ __forceinline__ void fetch_group(const VmGrid& t, const SampleRec& r, int i, int c, float4& pv, float4& lv, int4& po, float4& pw,
                                            int4& lr) {
  po = *reinterpret_cast<const int4*>(r.poff[i]);
  pw = *reinterpret_cast<const float4*>(r.pw[i]);
  lr = *reinterpret_cast<const int4*>(r.line[i]);
  const float* pl = (i == 0 ? t.plane[0] : (i == 1 ? t.plane[1] : t.plane[2])) + c;      // selects: no local copy of the params
  const float* ln = (i == 0 ? t.line[0] : (i == 1 ? t.line[1] : t.line[2])) + c;
  const float4 a0 = ldg4(pl + po.x), a1 = ldg4(pl + po.y), a2 = ldg4(pl + po.z), a3 = ldg4(pl + po.w);
  const float4 b0 = ldg4(ln + lr.x), b1 = ldg4(ln + lr.y);
  const float w0 = __int_as_float(lr.z), w1 = __int_as_float(lr.w);
  pv.x = a0.x * pw.x + a1.x * pw.y + a2.x * pw.z + a3.x * pw.w;
  pv.y = a0.y * pw.x + a1.y * pw.y + a2.y * pw.z + a3.y * pw.w;
  pv.z = a0.z * pw.x + a1.z * pw.y + a2.z * pw.z + a3.z * pw.w;
  pv.w = a0.w * pw.x + a1.w * pw.y + a2.w * pw.z + a3.w * pw.w;
  lv.x = b0.x * w0 + b1.x * w1; lv.y = b0.y * w0 + b1.y * w1; lv.z = b0.z * w0 + b1.z * w1; lv.w = b0.w * w0 + b1.w * w1;
}

__global__ void __launch_bounds__(CF_WARPS * 32) vm_color_features_fwd_kernel(VmGeom g, VmGrid t, GroupMap m, const float* __restrict__ view_dirs,
                                                                              uint2* __restrict__ rows) {
  __shared__ SampleRec s_rec[CF_WARPS][32];
  __shared__ uint2 s_vd[CF_WARPS][32];
  const int n = g.count[0];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gw = blockIdx.x * CF_WARPS + warp, nw = gridDim.x * CF_WARPS;
  for (int base = gw * 32; base < n; base += nw * 32) {
    const int cnt = min(32, n - base);
    if (lane < cnt) {
      const int flat = g.idx[base + lane];
      compute_record(g, t, flat, s_rec[warp][lane]);
      const float* vd = view_dirs + (size_t)(flat / g.S) * 3;
      s_vd[warp][lane] = make_uint2(ptx_pack_bf16(vd[0], vd[1]), ptx_pack_bf16(vd[2], 0.f));
    }
    __syncwarp();
    uint2* out = rows + (size_t)base * m.GP;
    for (int item = lane; item < cnt * m.GP; item += 32) {
      const int sidx = (int)(((unsigned)item * m.inv) >> 16);
      const int gq = item - sidx * m.GP;
      uint2 v = make_uint2(0u, 0u);
      if (gq < m.G) {
        const int i = (gq >= m.g0) + (gq >= m.g1);
        const int c = (gq - (i == 0 ? 0 : (i == 1 ? m.g0 : m.g1))) << 2;
        float4 pv, lv, pw; int4 po, lr;
        fetch_group(t, s_rec[warp][sidx], i, c, pv, lv, po, pw, lr);
        v = make_uint2(ptx_pack_bf16(pv.x * lv.x, pv.y * lv.y), ptx_pack_bf16(pv.z * lv.z, pv.w * lv.w));
      } else if (gq == m.G) {
        v = s_vd[warp][sidx];
      }
      out[item] = v;
    }
    __syncwarp();
  }
}

// run-merged appearance forward: consecutive compacted samples lie half a voxel apart on one ray, so they mostly share their
// bilinear footprints (in NDC the rays run along z: the (x, y) plane keeps its 2x2 texels for 5-20 samples, the lines along x / y
// likewise).  A warp takes 64 consecutive samples (records in shared memory, as the run-merged backward does); a lane owns a
// (run of FR samples, 4-channel group) item and walks the run, re-loading the four plane texels / two line texels only when the
// footprint changes: ~2 instead of 6 texel loads per sample and group.  The arithmetic per sample is the one of fetch_group().
template <int FR>
__global__ void __launch_bounds__(CB_WARPS * 32) vm_color_features_fwd_runs_kernel(VmGeom g, VmGrid t, GroupMap m, const float* __restrict__ view_dirs,
                                                                                   uint2* __restrict__ rows) {
  __shared__ SampleRec s_rec[CB_WARPS][CB_SAMPLES];
  __shared__ uint2 s_vd[CB_WARPS][CB_SAMPLES];
  const int n = g.count[0];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gw = blockIdx.x * CB_WARPS + warp, nw = gridDim.x * CB_WARPS;
  for (int base = gw * CB_SAMPLES; base < n; base += nw * CB_SAMPLES) {
    const int cnt = min(CB_SAMPLES, n - base);
    for (int sidx = lane; sidx < cnt; sidx += 32) {
      const int flat = g.idx[base + sidx];
      compute_record(g, t, flat, s_rec[warp][sidx]);
      const float* vd = view_dirs + (size_t)(flat / g.S) * 3;
      s_vd[warp][sidx] = make_uint2(ptx_pack_bf16(vd[0], vd[1]), ptx_pack_bf16(vd[2], 0.f));
    }
    __syncwarp();
    uint2* out = rows + (size_t)base * m.GP;
    const int nrun = (cnt + FR - 1) / FR;
    for (int item = lane; item < nrun * m.GP; item += 32) {
      const int run = (int)(((unsigned)item * m.inv) >> 16);          // m.inv describes GP groups per run here
      const int gq = item - run * m.GP;
      if (gq > m.G) {                                                   // zero padding behind the view directions
#pragma unroll
        for (int k = 0; k < FR; ++k)
          if (run * FR + k < cnt) out[(size_t)(run * FR + k) * m.GP + gq] = make_uint2(0u, 0u);
        continue;
      }
      if (gq == m.G) {
#pragma unroll
        for (int k = 0; k < FR; ++k)
          if (run * FR + k < cnt) out[(size_t)(run * FR + k) * m.GP + gq] = s_vd[warp][run * FR + k];
        continue;
      }
      const int i = (gq >= m.g0) + (gq >= m.g1);
      const int c = (gq - (i == 0 ? 0 : (i == 1 ? m.g0 : m.g1))) << 2;
      const float* pl = (i == 0 ? t.plane[0] : (i == 1 ? t.plane[1] : t.plane[2])) + c;
      const float* ln = (i == 0 ? t.line[0] : (i == 1 ? t.line[1] : t.line[2])) + c;
      int4 kpo = make_int4(-1, -1, -1, -1);
      int kl0 = -1, kl1 = -1;
      float4 a0 = f4_zero(), a1 = f4_zero(), a2 = f4_zero(), a3 = f4_zero(), b0 = f4_zero(), b1 = f4_zero();
#pragma unroll
      for (int k = 0; k < FR; ++k) {
        const int sidx = run * FR + k;
        if (sidx >= cnt) break;
        const SampleRec& r = s_rec[warp][sidx];
        const int4 po = *reinterpret_cast<const int4*>(r.poff[i]);
        const float4 pw = *reinterpret_cast<const float4*>(r.pw[i]);
        const int4 lr = *reinterpret_cast<const int4*>(r.line[i]);
        if (po.x != kpo.x || po.y != kpo.y || po.z != kpo.z || po.w != kpo.w) {
          kpo = po;
          a0 = ldg4(pl + po.x); a1 = ldg4(pl + po.y); a2 = ldg4(pl + po.z); a3 = ldg4(pl + po.w);
        }
        if (lr.x != kl0 || lr.y != kl1) {
          kl0 = lr.x; kl1 = lr.y;
          b0 = ldg4(ln + lr.x); b1 = ldg4(ln + lr.y);
        }
        const float w0 = __int_as_float(lr.z), w1 = __int_as_float(lr.w);
        float4 pv, lv;
        pv.x = a0.x * pw.x + a1.x * pw.y + a2.x * pw.z + a3.x * pw.w;
        pv.y = a0.y * pw.x + a1.y * pw.y + a2.y * pw.z + a3.y * pw.w;
        pv.z = a0.z * pw.x + a1.z * pw.y + a2.z * pw.z + a3.z * pw.w;
        pv.w = a0.w * pw.x + a1.w * pw.y + a2.w * pw.z + a3.w * pw.w;
        lv.x = b0.x * w0 + b1.x * w1; lv.y = b0.y * w0 + b1.y * w1; lv.z = b0.z * w0 + b1.z * w1; lv.w = b0.w * w0 + b1.w * w1;
        out[(size_t)sidx * m.GP + gq] = make_uint2(ptx_pack_bf16(pv.x * lv.x, pv.y * lv.y), ptx_pack_bf16(pv.z * lv.z, pv.w * lv.w));
      }
    }
    __syncwarp();
  }
}

// backward: g_rows[:, :CT] (fp32, row pitch `pitch` floats, a multiple of 4) scattered into the zero-initialised
// channels-last plane / line gradients with 16-byte vector reductions; same (sample, group) work split as the forward
__global__ void __launch_bounds__(CF_WARPS * 32) vm_color_features_bwd_kernel(VmGeom g, VmGrid t, GroupMap m, const float* __restrict__ g_rows, int pitch,
                                                                              float* gp0, float* gp1, float* gp2, float* gl0, float* gl1, float* gl2) {
  __shared__ SampleRec s_rec[CF_WARPS][32];
  const int n = g.count[0];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gw = blockIdx.x * CF_WARPS + warp, nw = gridDim.x * CF_WARPS;
  for (int base = gw * 32; base < n; base += nw * 32) {
    const int cnt = min(32, n - base);
    if (lane < cnt) compute_record(g, t, g.idx[base + lane], s_rec[warp][lane]);
    __syncwarp();
    for (int item = lane; item < cnt * m.G; item += 32) {
      const int sidx = (int)(((unsigned)item * m.inv) >> 16);        // m.inv / m.GP describe G groups per sample here
      const int gq = item - sidx * m.G;
      const float4 go = ldg4(g_rows + (size_t)(base + sidx) * pitch + gq * 4);
      if (go.x == 0.f && go.y == 0.f && go.z == 0.f && go.w == 0.f) continue;
      const int i = (gq >= m.g0) + (gq >= m.g1);
      const int c = (gq - (i == 0 ? 0 : (i == 1 ? m.g0 : m.g1))) << 2;
      float4 pv, lv, pw; int4 po, lr;
      fetch_group(t, s_rec[warp][sidx], i, c, pv, lv, po, pw, lr);
      float* gpl = (i == 0 ? gp0 : (i == 1 ? gp1 : gp2)) + c;
      float* gln = (i == 0 ? gl0 : (i == 1 ? gl1 : gl2)) + c;
      const float4 gpv = make_float4(go.x * lv.x, go.y * lv.y, go.z * lv.z, go.w * lv.w);
      const float4 glv = make_float4(go.x * pv.x, go.y * pv.y, go.z * pv.z, go.w * pv.w);
      if (pw.x != 0.f) red_add4(gpl + po.x, gpv.x * pw.x, gpv.y * pw.x, gpv.z * pw.x, gpv.w * pw.x);
      if (pw.y != 0.f) red_add4(gpl + po.y, gpv.x * pw.y, gpv.y * pw.y, gpv.z * pw.y, gpv.w * pw.y);
      if (pw.z != 0.f) red_add4(gpl + po.z, gpv.x * pw.z, gpv.y * pw.z, gpv.z * pw.z, gpv.w * pw.z);
      if (pw.w != 0.f) red_add4(gpl + po.w, gpv.x * pw.w, gpv.y * pw.w, gpv.z * pw.w, gpv.w * pw.w);
      const float w0 = __int_as_float(lr.z), w1 = __int_as_float(lr.w);
      if (w0 != 0.f) red_add4(gln + lr.x, glv.x * w0, glv.y * w0, glv.z * w0, glv.w * w0);
      if (w1 != 0.f) red_add4(gln + lr.y, glv.x * w1, glv.y * w1, glv.z * w1, glv.w * w1);
    }
    __syncwarp();
  }
}

// run-merged appearance backward (see vm_density_bwd_runs_kernel for the reasoning: the scatter is bound by the number of
// global reductions, and consecutive compacted samples share their bilinear footprints).  A warp takes 64 consecutive
// samples: their records go to shared memory (lane per sample, two rounds), then a lane owns a (run of 8 samples, 4-channel
// group) item and walks the run with the four corner gradients + two line gradients in registers, emitting reductions only
// when the footprint changes.  18 groups x 8 runs = 144 items per 64 samples (4.5 rounds of 32 lanes).

template <int CB_K>
__global__ void __launch_bounds__(CB_WARPS * 32) vm_color_features_bwd_runs_kernel(VmGeom g, VmGrid t, GroupMap m, const float* __restrict__ g_rows,
                                                                                   int pitch, float* gp0, float* gp1, float* gp2, float* gl0,
                                                                                   float* gl1, float* gl2) {
  __shared__ SampleRec s_rec[CB_WARPS][CB_SAMPLES];
  const int n = g.count[0];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gw = blockIdx.x * CB_WARPS + warp, nw = gridDim.x * CB_WARPS;
  for (int base = gw * CB_SAMPLES; base < n; base += nw * CB_SAMPLES) {
    const int cnt = min(CB_SAMPLES, n - base);
    for (int sidx = lane; sidx < cnt; sidx += 32) compute_record(g, t, g.idx[base + sidx], s_rec[warp][sidx]);
    __syncwarp();
    const int nrun = (cnt + CB_K - 1) / CB_K;
    for (int item = lane; item < nrun * m.G; item += 32) {
      const int run = (int)(((unsigned)item * m.inv) >> 16);          // m.inv describes G groups per run here
      const int gq = item - run * m.G;
      const int i = (gq >= m.g0) + (gq >= m.g1);
      const int c = (gq - (i == 0 ? 0 : (i == 1 ? m.g0 : m.g1))) << 2;
      const float* pl = (i == 0 ? t.plane[0] : (i == 1 ? t.plane[1] : t.plane[2])) + c;
      const float* ln = (i == 0 ? t.line[0] : (i == 1 ? t.line[1] : t.line[2])) + c;
      float* gpl = (i == 0 ? gp0 : (i == 1 ? gp1 : gp2)) + c;
      float* gln = (i == 0 ? gl0 : (i == 1 ? gl1 : gl2)) + c;
      int4 kpo = make_int4(-1, -1, -1, -1);
      int kl0 = -1, kl1 = -1;
      float4 tex[4], acc[4], ltex[2], lacc[2];
#pragma unroll
      for (int q = 0; q < 4; ++q) { tex[q] = f4_zero(); acc[q] = f4_zero(); }
      ltex[0] = ltex[1] = lacc[0] = lacc[1] = f4_zero();
#pragma unroll
      for (int k = 0; k <= CB_K; ++k) {
        const int sidx = run * CB_K + k;
        const bool last = k == CB_K || sidx >= cnt;
        float4 go = f4_zero();
        int4 po = kpo, lr = make_int4(kl0, kl1, 0, 0);
        float4 pw = f4_zero();
        if (!last) {
          go = ldg4(g_rows + (size_t)(base + sidx) * pitch + gq * 4);
          if (!f4_any(go)) continue;
          const SampleRec& r = s_rec[warp][sidx];
          po = *reinterpret_cast<const int4*>(r.poff[i]);
          pw = *reinterpret_cast<const float4*>(r.pw[i]);
          lr = *reinterpret_cast<const int4*>(r.line[i]);
        }
        if (last || po.x != kpo.x || po.y != kpo.y || po.z != kpo.z || po.w != kpo.w) {
          if (kpo.x >= 0) {
            if (f4_any(acc[0])) red_add4(gpl + kpo.x, acc[0].x, acc[0].y, acc[0].z, acc[0].w);
            if (f4_any(acc[1])) red_add4(gpl + kpo.y, acc[1].x, acc[1].y, acc[1].z, acc[1].w);
            if (f4_any(acc[2])) red_add4(gpl + kpo.z, acc[2].x, acc[2].y, acc[2].z, acc[2].w);
            if (f4_any(acc[3])) red_add4(gpl + kpo.w, acc[3].x, acc[3].y, acc[3].z, acc[3].w);
          }
#pragma unroll
          for (int q = 0; q < 4; ++q) acc[q] = f4_zero();
          if (!last) {
            kpo = po;
            tex[0] = ldg4(pl + po.x); tex[1] = ldg4(pl + po.y); tex[2] = ldg4(pl + po.z); tex[3] = ldg4(pl + po.w);
          }
        }
        if (last || lr.x != kl0 || lr.y != kl1) {
          if (kl0 >= 0) {
            if (f4_any(lacc[0])) red_add4(gln + kl0, lacc[0].x, lacc[0].y, lacc[0].z, lacc[0].w);
            if (f4_any(lacc[1])) red_add4(gln + kl1, lacc[1].x, lacc[1].y, lacc[1].z, lacc[1].w);
          }
          lacc[0] = lacc[1] = f4_zero();
          if (!last) {
            kl0 = lr.x; kl1 = lr.y;
            ltex[0] = ldg4(ln + lr.x); ltex[1] = ldg4(ln + lr.y);
          }
        }
        if (last) break;
        const float w0 = __int_as_float(lr.z), w1 = __int_as_float(lr.w);
        float4 pv = f4_zero(), lv = f4_zero();
        f4_fma(pv, pw.x, tex[0]); f4_fma(pv, pw.y, tex[1]); f4_fma(pv, pw.z, tex[2]); f4_fma(pv, pw.w, tex[3]);
        f4_fma(lv, w0, ltex[0]); f4_fma(lv, w1, ltex[1]);
        const float4 gpv = make_float4(go.x * lv.x, go.y * lv.y, go.z * lv.z, go.w * lv.w);
        const float4 glv = make_float4(go.x * pv.x, go.y * pv.y, go.z * pv.z, go.w * pv.w);
        f4_fma(acc[0], pw.x, gpv); f4_fma(acc[1], pw.y, gpv); f4_fma(acc[2], pw.z, gpv); f4_fma(acc[3], pw.w, gpv);
        f4_fma(lacc[0], w0, glv); f4_fma(lacc[1], w1, glv);
      }
    }
    __syncwarp();
  }
}

// scatter compacted rows [n, width] back to the dense [total, width] tensor (the reference's rgb[mask] = ..., :1271)
__global__ void scatter_rows_kernel(const int* __restrict__ idx, const int* __restrict__ count, const float* __restrict__ src,
                                    int width, float* __restrict__ dst) {
  const int n = count[0];
  const long long total = (long long)n * width;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(e / width), c = (int)(e % width);
    dst[(size_t)idx[j] * width + c] = src[e];
  }
}
// gather dense rows -> compacted rows (backward of the scatter)
__global__ void gather_rows_kernel(const int* __restrict__ idx, const int* __restrict__ count, const float* __restrict__ src,
                                   int width, float* __restrict__ dst) {
  const int n = count[0];
  const long long total = (long long)n * width;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(e / width), c = (int)(e % width);
    dst[e] = src[(size_t)idx[j] * width + c];
  }
}

inline int blocks_for(long long n, int per_block, int cap_mult = 32) {
  long long b = (n + per_block - 1) / per_block;
  const long long cap = (long long)sm_count() * cap_mult;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

}  // namespace srf

using namespace srf;

SRF_API int srf_pack_alpha_bits(const float* volume, int64_t num_voxels, uint32_t* bits, void* stream) {
  if (num_voxels == 0) return 0;
  SRF_REQUIRE(volume && bits, "srf_pack_alpha_bits", "null pointer");
  const long long words = (num_voxels + 31) / 32;
  pack_alpha_kernel<<<(unsigned)((words + 255) / 256), 256, 0, (cudaStream_t)stream>>>(volume, num_voxels, bits);
  return check_launch("srf_pack_alpha_bits");
}

SRF_API int srf_compaction_blocks(int64_t total) { return (int)((total + CMP_BLOCK - 1) / CMP_BLOCK); }

SRF_API int srf_tensorf_mask(const float* rays_o, const float* rays_d, const float* z, int64_t num_rays, int num_samples,
                             const float* bbox, const uint32_t* alpha_bits, const int* alpha_res, const float* alpha_box_min,
                             const float* alpha_box_size, uint8_t* mask, int* block_counts, void* stream) {
  if (num_rays == 0) return 0;
  SRF_REQUIRE(rays_o && rays_d && z && bbox && mask && block_counts, "srf_tensorf_mask", "null pointer");
  SRF_REQUIRE(alpha_bits == nullptr || (alpha_res && alpha_box_min && alpha_box_size), "srf_tensorf_mask", "alpha box missing");
  MaskParams p{};
  p.rays_o = rays_o; p.rays_d = rays_d; p.z = z; p.alpha_bits = alpha_bits; p.mask = mask; p.block_counts = block_counts;
  p.total = (long long)num_rays * num_samples; p.S = num_samples;
  for (int a = 0; a < 3; ++a) { p.bb0[a] = bbox[a]; p.bb1[a] = bbox[3 + a]; }     // HOST pointers: 6 floats
  if (alpha_bits) {
    for (int a = 0; a < 3; ++a) { p.ab0[a] = alpha_box_min[a]; p.asize[a] = alpha_box_size[a]; }
    p.ax = alpha_res[0]; p.ay = alpha_res[1]; p.az = alpha_res[2];
    SRF_REQUIRE(p.ax > 0 && p.ay > 0 && p.az > 0 && (long long)p.ax * p.ay * p.az < (1ll << 31), "srf_tensorf_mask",
                "alpha volume must hold fewer than 2^31 voxels");
  }
  tensorf_mask_kernel<<<srf_compaction_blocks(p.total), 256, 0, (cudaStream_t)stream>>>(p);
  return check_launch("srf_tensorf_mask");
}

SRF_API int srf_threshold_mask(const float* values, float threshold, int64_t total, uint8_t* mask, int* block_counts, void* stream) {
  if (total == 0) return 0;
  SRF_REQUIRE(values && mask && block_counts, "srf_threshold_mask", "null pointer");
  threshold_mask_kernel<<<srf_compaction_blocks(total), 256, 0, (cudaStream_t)stream>>>(values, threshold, total, mask, block_counts);
  return check_launch("srf_threshold_mask");
}

SRF_API int srf_compact(const uint8_t* mask, int64_t total, int* block_counts, int* block_offsets, int* indices, int* count,
                        void* stream) {
  SRF_REQUIRE(count, "srf_compact", "null pointer");
  if (total == 0) return cudaMemsetAsync(count, 0, sizeof(int), (cudaStream_t)stream) == cudaSuccess ? 0 : fail("srf_compact", "memset");
  SRF_REQUIRE(mask && block_counts && block_offsets && indices, "srf_compact", "null pointer");
  SRF_REQUIRE(total < (1ll << 31), "srf_compact", "more than 2^31 samples in one call");
  const int nb = srf_compaction_blocks(total);
  scan_blocks_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(block_counts, nb, block_offsets, count);
  compact_scatter_kernel<<<nb, 256, 0, (cudaStream_t)stream>>>(mask, total, block_counts, block_offsets, indices);
  return check_launch("srf_compact");
}

namespace {
int fill_geom(VmGeom& g, const float* rays_o, const float* rays_d, const float* z, const int* idx, const int* count, int S,
              const float* box_min, const float* box_size) {
  g.rays_o = rays_o; g.rays_d = rays_d; g.z = z; g.idx = idx; g.count = count; g.S = S; g.z_shared = 0;
  for (int a = 0; a < 3; ++a) { g.bb0[a] = box_min[a]; g.bsize[a] = box_size[a]; }
  return 0;
}
int fill_grid(VmGrid& t, const float* const* planes, const float* const* lines, const int* channels, const int* res, const char* where) {
  for (int i = 0; i < 3; ++i) {
    t.plane[i] = planes[i]; t.line[i] = lines[i]; t.C[i] = channels[i]; t.res[i] = res[i];
    if (!planes[i] || !lines[i]) return fail(where, "null plane/line pointer");
    if (channels[i] <= 0 || (channels[i] & 3)) return fail(where, "channel counts must be positive multiples of 4");
  }
  return 0;
}
}  // namespace

SRF_API int srf_vm_density_fwd(const float* rays_o, const float* rays_d, const float* z, int num_samples, const int* indices,
                               const int* count, int64_t max_count, const float* box_min, const float* box_size,
                               const float* const* planes, const float* const* lines, const int* channels, const int* resolution,
                               int softplus, float density_offset, float* sigma, float* features, void* stream) {
  if (max_count == 0) return 0;
  SRF_REQUIRE(rays_o && rays_d && z && indices && count && box_min && box_size && sigma, "srf_vm_density_fwd", "null pointer");
  VmGeom g; VmGrid t;
  fill_geom(g, rays_o, rays_d, z, indices, count, num_samples, box_min, box_size);
  if (fill_grid(t, planes, lines, channels, resolution, "srf_vm_density_fwd")) return 1;
  vm_density_fwd_kernel<<<blocks_for(max_count, 256), 256, 0, (cudaStream_t)stream>>>(g, t, softplus, density_offset, sigma, features);
  return check_launch("srf_vm_density_fwd");
}

SRF_API int srf_vm_density_bwd(const float* rays_o, const float* rays_d, const float* z, int num_samples, const int* indices,
                               const int* count, int64_t max_count, const float* box_min, const float* box_size,
                               const float* const* planes, const float* const* lines, const int* channels, const int* resolution,
                               int softplus, float density_offset, const float* g_sigma, const float* features,
                               float* const* g_planes, float* const* g_lines, void* stream) {
  if (max_count == 0) return 0;
  SRF_REQUIRE(rays_o && rays_d && z && indices && count && g_sigma && features && g_planes && g_lines, "srf_vm_density_bwd", "null pointer");
  VmGeom g; VmGrid t;
  fill_geom(g, rays_o, rays_d, z, indices, count, num_samples, box_min, box_size);
  if (fill_grid(t, planes, lines, channels, resolution, "srf_vm_density_bwd")) return 1;
  // run length K of the run-merged scatter, measured on the 331x368x220 training step (profiles/r02_tensorf_scatter.md):
  // one thread per sample (r1) 1.05 ms, K = 4: 0.41 ms, K = 8: 0.45 ms, K = 16: 0.58 ms (longer runs merge more but leave fewer threads)
  static const int variant = getenv("SRF_VM_BWD_VARIANT") ? atoi(getenv("SRF_VM_BWD_VARIANT")) : 4;     // 0: one thread per sample (r1)
  if (variant == 0) {
    vm_density_bwd_kernel<<<blocks_for(max_count, 256), 256, 0, (cudaStream_t)stream>>>(
        g, t, softplus, density_offset, g_sigma, features, g_planes[0], g_planes[1], g_planes[2], g_lines[0], g_lines[1], g_lines[2]);
  } else if (variant == 16) {
    vm_density_bwd_runs_kernel<16><<<blocks_for((max_count + 15) / 16, 128, 16), 128, 0, (cudaStream_t)stream>>>(
        g, t, softplus, density_offset, g_sigma, features, g_planes[0], g_planes[1], g_planes[2], g_lines[0], g_lines[1], g_lines[2]);
  } else if (variant == 2) {
    vm_density_bwd_runs_kernel<2><<<blocks_for((max_count + 1) / 2, 128, 16), 128, 0, (cudaStream_t)stream>>>(
        g, t, softplus, density_offset, g_sigma, features, g_planes[0], g_planes[1], g_planes[2], g_lines[0], g_lines[1], g_lines[2]);
  } else if (variant == 4) {
    vm_density_bwd_runs_kernel<4><<<blocks_for((max_count + 3) / 4, 128, 16), 128, 0, (cudaStream_t)stream>>>(
        g, t, softplus, density_offset, g_sigma, features, g_planes[0], g_planes[1], g_planes[2], g_lines[0], g_lines[1], g_lines[2]);
  } else {
    vm_density_bwd_runs_kernel<8><<<blocks_for((max_count + 7) / 8, 128, 16), 128, 0, (cudaStream_t)stream>>>(
        g, t, softplus, density_offset, g_sigma, features, g_planes[0], g_planes[1], g_planes[2], g_lines[0], g_lines[1], g_lines[2]);
  }
  return check_launch("srf_vm_density_bwd");
}

static int fill_groups(GroupMap& m, const int* channels, int groups_per_row, const char* where) {
  SRF_REQUIRE(channels[0] % 4 == 0 && channels[1] % 4 == 0 && channels[2] % 4 == 0, where, "component counts must be multiples of 4");
  m.g0 = channels[0] / 4; m.g1 = m.g0 + channels[1] / 4; m.G = m.g1 + channels[2] / 4;
  m.GP = groups_per_row;
  SRF_REQUIRE(m.GP >= 1 && m.GP <= 32, where, "row width must be 4..128 elements");
  m.inv = (65536u + (unsigned)m.GP - 1u) / (unsigned)m.GP;      // item / GP == (item * inv) >> 16 for item < 1024
  return 0;
}

SRF_API int srf_vm_color_features_fwd(const float* rays_o, const float* rays_d, const float* z, int num_samples, const int* indices,
                                      const int* count, int64_t max_count, const float* box_min, const float* box_size,
                                      const float* const* planes, const float* const* lines, const int* channels,
                                      const int* resolution, const float* view_dirs, void* rows, int row_pitch, int z_is_ladder,
                                      void* stream) {
  if (max_count == 0) return 0;
  SRF_REQUIRE(rays_o && rays_d && z && indices && count && view_dirs && rows, "srf_vm_color_features_fwd", "null pointer");
  VmGeom g; VmGrid t; GroupMap m;
  fill_geom(g, rays_o, rays_d, z, indices, count, num_samples, box_min, box_size);
  g.z_shared = z_is_ladder ? 1 : 0;
  if (fill_grid(t, planes, lines, channels, resolution, "srf_vm_color_features_fwd")) return 1;
  SRF_REQUIRE(row_pitch % 8 == 0 && row_pitch <= 128, "srf_vm_color_features_fwd", "row_pitch must be a multiple of 8, <= 128");
  if (fill_groups(m, channels, row_pitch / 4, "srf_vm_color_features_fwd")) return 1;
  SRF_REQUIRE(m.G + 1 <= m.GP, "srf_vm_color_features_fwd", "row_pitch must hold sum(C) + 3 elements");
  // run-merged gather (a lane walks K = 8 consecutive samples of one channel group, re-using texels while the footprint repeats);
  // measured on the 576x1024 frame of the bench scene: one (sample, group) item per lane 7.75 ms, K = 4: 6.58, K = 8: 6.13, K = 16: 6.51
  // (SRF_VM_CFWD_RUNS=0 / 4 / 8 selects the variant)
  static const int runs = getenv("SRF_VM_CFWD_RUNS") ? atoi(getenv("SRF_VM_CFWD_RUNS")) : SRF_VM_CFWD_RUNS_DEFAULT;
  if (runs == 4) {
    vm_color_features_fwd_runs_kernel<4><<<blocks_for(max_count, CB_WARPS * CB_SAMPLES, 12), CB_WARPS * 32, 0, (cudaStream_t)stream>>>(
        g, t, m, view_dirs, reinterpret_cast<uint2*>(rows));
  } else if (runs == 8) {
    vm_color_features_fwd_runs_kernel<8><<<blocks_for(max_count, CB_WARPS * CB_SAMPLES, 12), CB_WARPS * 32, 0, (cudaStream_t)stream>>>(
        g, t, m, view_dirs, reinterpret_cast<uint2*>(rows));
  } else {
    vm_color_features_fwd_kernel<<<blocks_for(max_count, CF_WARPS * 32, 16), CF_WARPS * 32, 0, (cudaStream_t)stream>>>(
        g, t, m, view_dirs, reinterpret_cast<uint2*>(rows));
  }
  return check_launch("srf_vm_color_features_fwd");
}

SRF_API int srf_vm_color_features_bwd(const float* rays_o, const float* rays_d, const float* z, int num_samples, const int* indices,
                                      const int* count, int64_t max_count, const float* box_min, const float* box_size,
                                      const float* const* planes, const float* const* lines, const int* channels,
                                      const int* resolution, const float* g_rows, int g_row_pitch, float* const* g_planes,
                                      float* const* g_lines, void* stream) {
  if (max_count == 0) return 0;
  SRF_REQUIRE(rays_o && rays_d && z && indices && count && g_rows && g_planes && g_lines, "srf_vm_color_features_bwd", "null pointer");
  VmGeom g; VmGrid t; GroupMap m;
  fill_geom(g, rays_o, rays_d, z, indices, count, num_samples, box_min, box_size);
  if (fill_grid(t, planes, lines, channels, resolution, "srf_vm_color_features_bwd")) return 1;
  const int CT = channels[0] + channels[1] + channels[2];
  SRF_REQUIRE(CT <= 128 && g_row_pitch >= CT && g_row_pitch % 4 == 0, "srf_vm_color_features_bwd",
              "need sum(C) <= 128 and a pitch >= sum(C) that is a multiple of 4");
  if (fill_groups(m, channels, CT / 4, "srf_vm_color_features_bwd")) return 1;
  static const int variant = getenv("SRF_VM_BWD_VARIANT") ? atoi(getenv("SRF_VM_BWD_VARIANT")) : 8;     // 0: one (sample, group) item per lane (r1)
  static const int color_k = getenv("SRF_VM_CBWD_K") ? atoi(getenv("SRF_VM_CBWD_K")) : 8;
  if (variant == 0) {
    vm_color_features_bwd_kernel<<<blocks_for(max_count, CF_WARPS * 32, 16), CF_WARPS * 32, 0, (cudaStream_t)stream>>>(
        g, t, m, g_rows, g_row_pitch, g_planes[0], g_planes[1], g_planes[2], g_lines[0], g_lines[1], g_lines[2]);
  } else if (color_k == 4) {
    vm_color_features_bwd_runs_kernel<4><<<blocks_for(max_count, CB_WARPS * CB_SAMPLES, 12), CB_WARPS * 32, 0, (cudaStream_t)stream>>>(
        g, t, m, g_rows, g_row_pitch, g_planes[0], g_planes[1], g_planes[2], g_lines[0], g_lines[1], g_lines[2]);
  } else {
    vm_color_features_bwd_runs_kernel<8><<<blocks_for(max_count, CB_WARPS * CB_SAMPLES, 12), CB_WARPS * 32, 0, (cudaStream_t)stream>>>(
        g, t, m, g_rows, g_row_pitch, g_planes[0], g_planes[1], g_planes[2], g_lines[0], g_lines[1], g_lines[2]);
  }
  return check_launch("srf_vm_color_features_bwd");
}

SRF_API int srf_scatter_rows(const int* indices, const int* count, int64_t max_count, const float* src, int width, float* dst,
                             void* stream) {
  if (max_count == 0) return 0;
  SRF_REQUIRE(indices && count && src && dst && width > 0, "srf_scatter_rows", "null pointer");
  scatter_rows_kernel<<<blocks_for(max_count * width, 256), 256, 0, (cudaStream_t)stream>>>(indices, count, src, width, dst);
  return check_launch("srf_scatter_rows");
}

SRF_API int srf_gather_rows(const int* indices, const int* count, int64_t max_count, const float* src, int width, float* dst,
                            void* stream) {
  if (max_count == 0) return 0;
  SRF_REQUIRE(indices && count && src && dst && width > 0, "srf_gather_rows", "null pointer");
  gather_rows_kernel<<<blocks_for(max_count * width, 256), 256, 0, (cudaStream_t)stream>>>(indices, count, src, width, dst);
  return check_launch("srf_gather_rows");
}
