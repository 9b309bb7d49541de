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

#include "common.cuh"

namespace srf {

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

struct MaskParams {
  const float* rays_o; const float* rays_d; const float* z;
  const uint32_t* alpha_bits;          // nullptr: no alphaMask
  uint8_t* mask; int* block_counts;
  long long total; int S;
  float bb0[3], bb1[3];                // tensor bounding box
  float ab0[3], asize[3];              // alpha-volume box: min corner and size (fp32, as the reference stores them)
  int ax, ay, az;                      // alpha-volume resolution
};

// trilinear grid_sample(align_corners=True, zero padding) of the {0,1} volume is > 0 iff some in-range corner with a
// positive fp32 weight product holds a 1 (ATen's evaluation order for coordinates and weights is reproduced exactly)
__device__ __forceinline__ bool alpha_hit(const MaskParams& p, const float (&pt)[3]) {
  const int dims[3] = {p.ax, p.ay, p.az};
  int i0[3];
  float w[3][2];
  bool cand[3][2];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    // normalise ((p - b0) / size) * 2 - 1  (:1347-1349), unnormalise ((c + 1) / 2) * (dim - 1) (ATen GridSampler.h);
    // the division by 2 is an exact scaling, so a multiplication by 0.5 gives the same bits
    const float c = __fadd_rn(__fmul_rn(__fdiv_rn(__fadd_rn(pt[a], -p.ab0[a]), p.asize[a]), 2.f), -1.f);
    const float ix = __fmul_rn(__fmul_rn(__fadd_rn(c, 1.f), 0.5f), (float)(dims[a] - 1));
    const float f0 = floorf(ix);
    i0[a] = (int)f0;
    w[a][0] = __fadd_rn(__fadd_rn(f0, 1.f), -ix);
    w[a][1] = __fadd_rn(ix, -f0);
    cand[a][0] = i0[a] >= 0 && i0[a] < dims[a] && w[a][0] > 0.f;
    cand[a][1] = i0[a] + 1 >= 0 && i0[a] + 1 < dims[a] && w[a][1] > 0.f;
  }
  const int v0 = (i0[2] * p.ay + i0[1]) * p.ax + i0[0];        // < 2^31 voxels (checked by the host)
  const int sy = p.ax, sz = p.ax * p.ay;
  bool hit = false;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const int dx = c & 1, dy = (c >> 1) & 1, dz = c >> 2;
    if (!(cand[0][dx] && cand[1][dy] && cand[2][dz])) continue;
    if (!(__fmul_rn(__fmul_rn(w[0][dx], w[1][dy]), w[2][dz]) > 0.f)) continue;      // the product itself may underflow to 0
    const int v = v0 + dx + dy * sy + dz * sz;
    hit |= (__ldg(p.alpha_bits + (v >> 5)) >> (v & 31)) & 1u;
  }
  return hit;
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

// ------------------------------------------------------------------------------------------ VM gathers
struct VmGrid {
  const float* plane[3];   // channels-last [H][W][C]
  const float* line[3];    // [L][C]
  int C[3];                // channels per plane/line pair (multiples of 4)
  int res[3];              // tensor resolution (X, Y, Z)
};

struct VmGeom {
  const float* rays_o; const float* rays_d; const float* z;
  const int* idx; const int* count;
  int S;
  float bb0[3], bsize[3];
};

// matrix_axes = [[0,1],[0,2],[1,2]], vector_axes = [2,1,0]  (:1131-1132); grid x -> W = res[a0], y -> H = res[a1]
// (functions, not __constant__ tables: after unrolling the axis is a compile-time constant and nothing is indexed dynamically)
__host__ __device__ __forceinline__ constexpr int axis0(int i) { return i == 2 ? 1 : 0; }
__host__ __device__ __forceinline__ constexpr int axis1(int i) { return i == 0 ? 1 : 2; }
__host__ __device__ __forceinline__ constexpr int axisv(int i) { return 2 - i; }

struct Bilerp {
  int x0, y0, W, H;
  float wx0, wx1, wy0, wy1;
};

__device__ __forceinline__ void normalized_point(const VmGeom& g, int flat, float (&pn)[3]) {
  const int r = flat / g.S;
  const float zz = g.z[flat];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const float pt = __fadd_rn(g.rays_o[r * 3 + a], __fmul_rn(g.rays_d[r * 3 + a], zz));
    pn[a] = __fadd_rn(__fmul_rn(__fdiv_rn(__fadd_rn(pt, -g.bb0[a]), g.bsize[a]), 2.f), -1.f);
  }
}

__device__ __forceinline__ float unnorm(float c, int size) { return __fmul_rn(__fdiv_rn(__fadd_rn(c, 1.f), 2.f), (float)(size - 1)); }

__device__ __forceinline__ Bilerp plane_coords(const float (&pn)[3], const int (&res)[3], int i) {
  Bilerp b;
  b.W = res[axis0(i)]; b.H = res[axis1(i)];
  const float ix = unnorm(pn[axis0(i)], b.W), iy = unnorm(pn[axis1(i)], b.H);
  const float fx = floorf(ix), fy = floorf(iy);
  b.x0 = (int)fx; b.y0 = (int)fy;
  b.wx1 = ix - fx; b.wx0 = (fx + 1.f) - ix;
  b.wy1 = iy - fy; b.wy0 = (fy + 1.f) - iy;
  return b;
}

// value of channel group [c, c+4) of a plane at the bilinear position (zeros outside, as grid_sample pads)
__device__ __forceinline__ float4 plane_fetch4(const float* plane, const Bilerp& b, int C, int c) {
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int x = b.x0 + (k & 1), y = b.y0 + (k >> 1);
    if (x < 0 || y < 0 || x >= b.W || y >= b.H) continue;
    const float w = ((k & 1) ? b.wx1 : b.wx0) * ((k >> 1) ? b.wy1 : b.wy0);
    const float4 t = __ldg(reinterpret_cast<const float4*>(plane + ((size_t)y * b.W + x) * C + c));
    acc.x += t.x * w; acc.y += t.y * w; acc.z += t.z * w; acc.w += t.w * w;
  }
  return acc;
}

__device__ __forceinline__ void line_coords(const float (&pn)[3], const int (&res)[3], int i, int& l0, int& L, float& w0, float& w1) {
  L = res[axisv(i)];
  const float iy = unnorm(pn[axisv(i)], L);
  const float fy = floorf(iy);
  l0 = (int)fy;
  w1 = iy - fy; w0 = (fy + 1.f) - iy;
}

__device__ __forceinline__ float4 line_fetch4(const float* line, int l0, int L, float w0, float w1, int C, int c) {
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (l0 >= 0 && l0 < L) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(line + (size_t)l0 * C + c));
    acc.x += t.x * w0; acc.y += t.y * w0; acc.z += t.z * w0; acc.w += t.w * w0;
  }
  if (l0 + 1 >= 0 && l0 + 1 < L) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(line + (size_t)(l0 + 1) * C + c));
    acc.x += t.x * w1; acc.y += t.y * w1; acc.z += t.z * w1; acc.w += t.w * w1;
  }
  return acc;
}

__device__ __forceinline__ void red_add4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
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

// appearance: products (plane x line) over all channels -> bf16 rows of `pitch` elements
// [products (CT = sum C) | view_dirs (3) | zero pad]: the A operand of the tensor-core colour MLP, whose first layer
// absorbs basis_matrix_color (W0' = [W0[:, :F] B | W0[:, F:]]), so no matrix-vector product is left in this kernel.
// A warp takes 32 samples.  Phase 1, lane per sample: coordinates -> per plane 4 clamped corner offsets + weights
// (zero weight = grid_sample's zero padding) and per line 2 offsets + weights, into shared memory.  Phase 2, lane per
// (sample, 4-channel group): 6 float4 texel loads, 28 FMAs, one 8-byte store; consecutive lanes write consecutive
// 8-byte pieces of the row-major output, so stores are fully coalesced and no lane idles on a short channel list.
constexpr int CF_WARPS = 8;

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

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

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
  g.rays_o = rays_o; g.rays_d = rays_d; g.z = z; g.idx = idx; g.count = count; g.S = S;
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
  vm_density_bwd_kernel<<<blocks_for(max_count, 256), 256, 0, (cudaStream_t)stream>>>(
      g, t, softplus, density_offset, g_sigma, features, g_planes[0], g_planes[1], g_planes[2], g_lines[0], g_lines[1], g_lines[2]);
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
                                      const int* resolution, const float* view_dirs, void* rows, int row_pitch, void* stream) {
  if (max_count == 0) return 0;
  SRF_REQUIRE(rays_o && rays_d && z && indices && count && view_dirs && rows, "srf_vm_color_features_fwd", "null pointer");
  VmGeom g; VmGrid t; GroupMap m;
  fill_geom(g, rays_o, rays_d, z, indices, count, num_samples, box_min, box_size);
  if (fill_grid(t, planes, lines, channels, resolution, "srf_vm_color_features_fwd")) return 1;
  SRF_REQUIRE(row_pitch % 8 == 0 && row_pitch <= 128, "srf_vm_color_features_fwd", "row_pitch must be a multiple of 8, <= 128");
  if (fill_groups(m, channels, row_pitch / 4, "srf_vm_color_features_fwd")) return 1;
  SRF_REQUIRE(m.G + 1 <= m.GP, "srf_vm_color_features_fwd", "row_pitch must hold sum(C) + 3 elements");
  vm_color_features_fwd_kernel<<<blocks_for(max_count, CF_WARPS * 32, 16), CF_WARPS * 32, 0, (cudaStream_t)stream>>>(
      g, t, m, view_dirs, reinterpret_cast<uint2*>(rows));
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
  vm_color_features_bwd_kernel<<<blocks_for(max_count, CF_WARPS * 32, 16), CF_WARPS * 32, 0, (cudaStream_t)stream>>>(
      g, t, m, g_rows, g_row_pitch, g_planes[0], g_planes[1], g_planes[2], g_lines[0], g_lines[1], g_lines[2]);
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
