// Shared device helpers of the Simple-TensoRF kernels (tensorf.cu, tensorf_march.cu): occupancy test against the bit-packed
// alpha volume with ATen's grid_sampler_3d coordinate arithmetic, and the VM plane / line addressing of
// src/models/SimpleTensoRF09.py:1131-1132, :763-765, :1214-1263.
#pragma once
#include "common.cuh"

namespace srf {

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
  int z_shared;            // z is ONE [S] ladder shared by all rays (test time) instead of [R,S]
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
  const float zz = g.z[g.z_shared ? flat - r * g.S : flat];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const float pt = __fadd_rn(g.rays_o[r * 3 + a], __fmul_rn(g.rays_d[r * 3 + a], zz));
    pn[a] = __fadd_rn(__fmul_rn(__fdiv_rn(__fadd_rn(pt, -g.bb0[a]), g.bsize[a]), 2.f), -1.f);
  }
}

// ATen: ((c + 1) / 2) * (size - 1); the division by 2 is an exact scaling, so a multiplication by 0.5 gives the same bits
__device__ __forceinline__ float unnorm(float c, int size) { return __fmul_rn(__fmul_rn(__fadd_rn(c, 1.f), 0.5f), (float)(size - 1)); }

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

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

__device__ __forceinline__ void red_add4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

}  // namespace srf
