// Simple-TensoRF CANDECOMP/PARAFAC (CP) tensor: density and appearance from three line factors per component.
//
// Replaces (reference file:line, relative to the upstream checkout):
//   srf_cp_density_fwd / _bwd          src/models/SimpleTensoRF09.py:763-765 (normalise), :1043-1062 (CpDecomposedTensor.get_volume_density:
//                                      three 1-D grid_samples, their product, the sum over components, the density predictor) and its autograd
//   srf_cp_color_features_fwd / _bwd   :1064-1078 (the same products for the appearance lines, the input of basis_matrix_color) and its autograd
//
// `decomposition_type = "CandecompParafac"` (:537-539) is selected by no shipped configuration; it shares everything but these two
// gathers with the vector-matrix tensor (occupancy test, compaction, compositing, colour MLP with basis_matrix_color folded into its
// first layer).  Lines are read from the channels-last derived caches [L][C] of tensorf_ops.to_channels_last, line i running along axis
// vector_axes[i] = 2 - i (:969).  All three lines of a tensor are a few hundred KB at most: every read is an L1 / L2 hit, and the
// gradient scatter lands on a few thousand distinct 16-byte vectors.
#include <cuda_bf16.h>

#include "tensorf_common.cuh"

namespace srf {

struct CpGrid {
  const float* line[3];    // [L_i][C], L_i = res[axisv(i)]
  int C;                   // components (a multiple of 4)
  int res[3];              // tensor resolution (X, Y, Z)
};

// the two taps of one line: element offsets (clamped into the line) and weights (zero outside: grid_sample's zero padding)
struct LineTap { int o0, o1; float w0, w1; };

__device__ __forceinline__ LineTap line_tap(const float (&pn)[3], const int (&res)[3], int i, int C) {
  int l0, L; float w0, w1;
  line_coords(pn, res, i, l0, L, w0, w1);
  LineTap t;
  const bool in0 = l0 >= 0 && l0 < L, in1 = l0 + 1 >= 0 && l0 + 1 < L;
  t.o0 = (in0 ? l0 : 0) * C; t.o1 = (in1 ? l0 + 1 : 0) * C;
  t.w0 = in0 ? w0 : 0.f; t.w1 = in1 ? w1 : 0.f;
  return t;
}

__device__ __forceinline__ float4 tap_fetch4(const float* line, const LineTap& t, int c) {
  const float4 a = ldg4(line + t.o0 + c), b = ldg4(line + t.o1 + c);
  // 0 + a w0 + b w1, in grid_sample's accumulation order
  return make_float4(a.x * t.w0 + b.x * t.w1, a.y * t.w0 + b.y * t.w1, a.z * t.w0 + b.z * t.w1, a.w * t.w0 + b.w * t.w1);
}

__device__ __forceinline__ float4 mul4(const float4& a, const float4& b) { return make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }

// density: sigma = act(sum_c l0_c l1_c l2_c); one thread per compacted sample
__global__ void __launch_bounds__(256) cp_density_fwd_kernel(VmGeom g, CpGrid t, int softplus, float offset, float* __restrict__ sigma,
                                                             float* __restrict__ feat_out) {
  const int n = g.count[0];
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
    const int flat = g.idx[j];
    float pn[3];
    normalized_point(g, flat, pn);
    const LineTap t0 = line_tap(pn, t.res, 0, t.C), t1 = line_tap(pn, t.res, 1, t.C), t2 = line_tap(pn, t.res, 2, t.C);
    float feat = 0.f;
    for (int c = 0; c < t.C; c += 4) {
      const float4 v = mul4(mul4(tap_fetch4(t.line[0], t0, c), tap_fetch4(t.line[1], t1, c)), tap_fetch4(t.line[2], t2, c));
      feat += (v.x + v.y) + (v.z + v.w);
    }
    if (feat_out) feat_out[j] = feat;
    float s;
    if (softplus) { const float x = feat + offset; s = x > 20.f ? x : log1pf(expf(x)); }
    else s = fmaxf(feat, 0.f);
    sigma[flat] = s;
  }
}

__device__ __forceinline__ void tap_scatter4(float* gline, const LineTap& t, int c, const float4& gv) {
  if (t.w0 != 0.f) red_add4(gline + t.o0 + c, gv.x * t.w0, gv.y * t.w0, gv.z * t.w0, gv.w * t.w0);
  if (t.w1 != 0.f) red_add4(gline + t.o1 + c, gv.x * t.w1, gv.y * t.w1, gv.z * t.w1, gv.w * t.w1);
}

__global__ void __launch_bounds__(256) cp_density_bwd_kernel(VmGeom g, CpGrid t, int softplus, float offset, const float* __restrict__ g_sigma,
                                                             const float* __restrict__ feat_in, float* gl0, float* gl1, float* gl2) {
  const int n = g.count[0];
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
    const int flat = g.idx[j];
    const float feat = feat_in[j];
    float gf = g_sigma[flat];
    if (softplus) { const float x = feat + offset; gf *= 1.f / (1.f + expf(-x)); }
    else gf = feat > 0.f ? gf : 0.f;
    if (gf == 0.f) continue;
    float pn[3];
    normalized_point(g, flat, pn);
    const LineTap t0 = line_tap(pn, t.res, 0, t.C), t1 = line_tap(pn, t.res, 1, t.C), t2 = line_tap(pn, t.res, 2, t.C);
    const float4 g4 = make_float4(gf, gf, gf, gf);
    for (int c = 0; c < t.C; c += 4) {
      const float4 a = tap_fetch4(t.line[0], t0, c), b = tap_fetch4(t.line[1], t1, c), d = tap_fetch4(t.line[2], t2, c);
      tap_scatter4(gl0, t0, c, mul4(g4, mul4(b, d)));
      tap_scatter4(gl1, t1, c, mul4(g4, mul4(a, d)));
      tap_scatter4(gl2, t2, c, mul4(g4, mul4(a, b)));
    }
  }
}

// appearance rows: bf16 [products (C) | view_dirs (3) | zero pad], `GP` 4-element groups per row.  A warp takes 32 samples: lane per sample
// computes the six taps into shared memory, then lane per (sample, 4-channel group): six float4 loads, one 8-byte store, coalesced
constexpr int CPF_WARPS = 8;

struct alignas(16) CpRec { int off[3][2]; float w[3][2]; };

__device__ __forceinline__ void cp_record(const VmGeom& g, const CpGrid& t, int flat, CpRec& r) {
  float pn[3];
  normalized_point(g, flat, pn);
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const LineTap tp = line_tap(pn, t.res, i, t.C);
    r.off[i][0] = tp.o0; r.off[i][1] = tp.o1; r.w[i][0] = tp.w0; r.w[i][1] = tp.w1;
  }
}

__device__ __forceinline__ float4 rec_fetch4(const float* line, const CpRec& r, int i, int c) {
  const float4 a = ldg4(line + r.off[i][0] + c), b = ldg4(line + r.off[i][1] + c);
  const float w0 = r.w[i][0], w1 = r.w[i][1];
  return make_float4(a.x * w0 + b.x * w1, a.y * w0 + b.y * w1, a.z * w0 + b.z * w1, a.w * w0 + b.w * w1);
}

__device__ __forceinline__ uint32_t cp_pack_bf16(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

__global__ void __launch_bounds__(CPF_WARPS * 32) cp_color_features_fwd_kernel(VmGeom g, CpGrid t, int G, int GP, unsigned inv,
                                                                               const float* __restrict__ view_dirs, uint2* __restrict__ rows) {
  __shared__ CpRec s_rec[CPF_WARPS][32];
  __shared__ uint2 s_vd[CPF_WARPS][32];
  const int n = g.count[0];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gw = blockIdx.x * CPF_WARPS + warp, nw = gridDim.x * CPF_WARPS;
  for (int base = gw * 32; base < n; base += nw * 32) {
    const int cnt = min(32, n - base);
    if (lane < cnt) {
      const int flat = g.idx[base + lane];
      cp_record(g, t, flat, s_rec[warp][lane]);
      const float* vd = view_dirs + (size_t)(flat / g.S) * 3;
      s_vd[warp][lane] = make_uint2(cp_pack_bf16(vd[0], vd[1]), cp_pack_bf16(vd[2], 0.f));
    }
    __syncwarp();
    uint2* out = rows + (size_t)base * GP;
    for (int item = lane; item < cnt * GP; item += 32) {
      const int sidx = (int)(((unsigned)item * inv) >> 16);
      const int gq = item - sidx * GP;
      uint2 v = make_uint2(0u, 0u);
      if (gq < G) {
        const CpRec& r = s_rec[warp][sidx];
        const int c = gq << 2;
        const float4 p = mul4(mul4(rec_fetch4(t.line[0], r, 0, c), rec_fetch4(t.line[1], r, 1, c)), rec_fetch4(t.line[2], r, 2, c));
        v = make_uint2(cp_pack_bf16(p.x, p.y), cp_pack_bf16(p.z, p.w));
      } else if (gq == G) {
        v = s_vd[warp][sidx];
      }
      out[item] = v;
    }
    __syncwarp();
  }
}

__device__ __forceinline__ void rec_scatter4(float* gline, const CpRec& r, int i, int c, const float4& gv) {
  const float w0 = r.w[i][0], w1 = r.w[i][1];
  if (w0 != 0.f) red_add4(gline + r.off[i][0] + c, gv.x * w0, gv.y * w0, gv.z * w0, gv.w * w0);
  if (w1 != 0.f) red_add4(gline + r.off[i][1] + c, gv.x * w1, gv.y * w1, gv.z * w1, gv.w * w1);
}

// backward: g_rows[:, :C] (fp32, row pitch `pitch` floats) scattered into the zero-initialised channels-last line gradients
__global__ void __launch_bounds__(CPF_WARPS * 32) cp_color_features_bwd_kernel(VmGeom g, CpGrid t, int G, unsigned inv,
                                                                               const float* __restrict__ g_rows, int pitch, float* gl0,
                                                                               float* gl1, float* gl2) {
  __shared__ CpRec s_rec[CPF_WARPS][32];
  const int n = g.count[0];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gw = blockIdx.x * CPF_WARPS + warp, nw = gridDim.x * CPF_WARPS;
  for (int base = gw * 32; base < n; base += nw * 32) {
    const int cnt = min(32, n - base);
    if (lane < cnt) cp_record(g, t, g.idx[base + lane], s_rec[warp][lane]);
    __syncwarp();
    for (int item = lane; item < cnt * G; item += 32) {
      const int sidx = (int)(((unsigned)item * inv) >> 16);
      const int gq = item - sidx * G;
      const float4 go = ldg4(g_rows + (size_t)(base + sidx) * pitch + gq * 4);
      if (go.x == 0.f && go.y == 0.f && go.z == 0.f && go.w == 0.f) continue;
      const CpRec& r = s_rec[warp][sidx];
      const int c = gq << 2;
      const float4 a = rec_fetch4(t.line[0], r, 0, c), b = rec_fetch4(t.line[1], r, 1, c), d = rec_fetch4(t.line[2], r, 2, c);
      rec_scatter4(gl0, r, 0, c, mul4(go, mul4(b, d)));
      rec_scatter4(gl1, r, 1, c, mul4(go, mul4(a, d)));
      rec_scatter4(gl2, r, 2, c, mul4(go, mul4(a, b)));
    }
    __syncwarp();
  }
}

namespace {
int cp_blocks(long long n, int per_block, int cap_mult) {
  long long b = (n + per_block - 1) / per_block;
  const long long cap = (long long)sm_count() * cap_mult;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}
void cp_fill_geom(VmGeom& g, const float* rays_o, const float* rays_d, const float* z, const int* idx, const int* count, int S,
                  const float* box_min, const float* box_size) {
  g.rays_o = rays_o; g.rays_d = rays_d; g.z = z; g.idx = idx; g.count = count; g.S = S; g.z_shared = 0;
  for (int a = 0; a < 3; ++a) { g.bb0[a] = box_min[a]; g.bsize[a] = box_size[a]; }
}
int cp_fill_grid(CpGrid& t, const float* const* lines, int components, const int* res, const char* where) {
  if (!lines || !res) return fail(where, "null pointer");
  if (components <= 0 || (components & 3)) return fail(where, "the component count must be a positive multiple of 4");
  t.C = components;
  for (int i = 0; i < 3; ++i) {
    if (!lines[i]) return fail(where, "null line pointer");
    if (res[i] <= 0) return fail(where, "empty grid");
    t.line[i] = lines[i]; t.res[i] = res[i];
  }
  return 0;
}
}  // namespace

}  // namespace srf

using namespace srf;

SRF_API int srf_cp_density_fwd(const float* rays_o, const float* rays_d, const float* z, int num_samples, const int* indices, const int* count,
                               int64_t max_count, const float* box_min, const float* box_size, const float* const* lines, int components,
                               const int* resolution, int softplus, float density_offset, float* sigma, float* features, void* stream) {
  if (max_count == 0) return 0;
  SRF_REQUIRE(rays_o && rays_d && z && indices && count && box_min && box_size && sigma, "srf_cp_density_fwd", "null pointer");
  VmGeom g; CpGrid t;
  cp_fill_geom(g, rays_o, rays_d, z, indices, count, num_samples, box_min, box_size);
  if (cp_fill_grid(t, lines, components, resolution, "srf_cp_density_fwd")) return 1;
  cp_density_fwd_kernel<<<cp_blocks(max_count, 256, 32), 256, 0, (cudaStream_t)stream>>>(g, t, softplus, density_offset, sigma, features);
  return check_launch("srf_cp_density_fwd");
}

SRF_API int srf_cp_density_bwd(const float* rays_o, const float* rays_d, const float* z, int num_samples, const int* indices, const int* count,
                               int64_t max_count, const float* box_min, const float* box_size, const float* const* lines, int components,
                               const int* resolution, int softplus, float density_offset, const float* g_sigma, const float* features,
                               float* const* g_lines, void* stream) {
  if (max_count == 0) return 0;
  SRF_REQUIRE(rays_o && rays_d && z && indices && count && box_min && box_size && g_sigma && features && g_lines, "srf_cp_density_bwd",
              "null pointer");
  SRF_REQUIRE(g_lines[0] && g_lines[1] && g_lines[2], "srf_cp_density_bwd", "null gradient pointer");
  VmGeom g; CpGrid t;
  cp_fill_geom(g, rays_o, rays_d, z, indices, count, num_samples, box_min, box_size);
  if (cp_fill_grid(t, lines, components, resolution, "srf_cp_density_bwd")) return 1;
  cp_density_bwd_kernel<<<cp_blocks(max_count, 256, 32), 256, 0, (cudaStream_t)stream>>>(g, t, softplus, density_offset, g_sigma, features,
                                                                                        g_lines[0], g_lines[1], g_lines[2]);
  return check_launch("srf_cp_density_bwd");
}

SRF_API int srf_cp_color_features_fwd(const float* rays_o, const float* rays_d, const float* z, int num_samples, const int* indices,
                                      const int* count, int64_t max_count, const float* box_min, const float* box_size,
                                      const float* const* lines, int components, const int* resolution, const float* view_dirs, void* rows,
                                      int row_pitch, void* stream) {
  if (max_count == 0) return 0;
  SRF_REQUIRE(rays_o && rays_d && z && indices && count && box_min && box_size && view_dirs && rows, "srf_cp_color_features_fwd", "null pointer");
  VmGeom g; CpGrid t;
  cp_fill_geom(g, rays_o, rays_d, z, indices, count, num_samples, box_min, box_size);
  if (cp_fill_grid(t, lines, components, resolution, "srf_cp_color_features_fwd")) return 1;
  SRF_REQUIRE(row_pitch % 8 == 0 && row_pitch <= 128, "srf_cp_color_features_fwd", "row_pitch must be a multiple of 8, <= 128");
  const int G = components / 4, GP = row_pitch / 4;
  SRF_REQUIRE(G + 1 <= GP, "srf_cp_color_features_fwd", "row_pitch must hold components + 3 elements");
  const unsigned inv = (65536u + (unsigned)GP - 1u) / (unsigned)GP;         // item / GP == (item * inv) >> 16 for item < 1024
  cp_color_features_fwd_kernel<<<cp_blocks(max_count, CPF_WARPS * 32, 16), CPF_WARPS * 32, 0, (cudaStream_t)stream>>>(
      g, t, G, GP, inv, view_dirs, reinterpret_cast<uint2*>(rows));
  return check_launch("srf_cp_color_features_fwd");
}

SRF_API int srf_cp_color_features_bwd(const float* rays_o, const float* rays_d, const float* z, int num_samples, const int* indices,
                                      const int* count, int64_t max_count, const float* box_min, const float* box_size,
                                      const float* const* lines, int components, const int* resolution, const float* g_rows,
                                      int g_row_pitch, float* const* g_lines, void* stream) {
  if (max_count == 0) return 0;
  SRF_REQUIRE(rays_o && rays_d && z && indices && count && box_min && box_size && g_rows && g_lines, "srf_cp_color_features_bwd",
              "null pointer");
  SRF_REQUIRE(g_lines[0] && g_lines[1] && g_lines[2], "srf_cp_color_features_bwd", "null gradient pointer");
  VmGeom g; CpGrid t;
  cp_fill_geom(g, rays_o, rays_d, z, indices, count, num_samples, box_min, box_size);
  if (cp_fill_grid(t, lines, components, resolution, "srf_cp_color_features_bwd")) return 1;
  SRF_REQUIRE(components <= 128 && g_row_pitch >= components && g_row_pitch % 4 == 0, "srf_cp_color_features_bwd",
              "need components <= 128 and a pitch >= components that is a multiple of 4");
  const int G = components / 4;
  const unsigned inv = (65536u + (unsigned)G - 1u) / (unsigned)G;
  cp_color_features_bwd_kernel<<<cp_blocks(max_count, CPF_WARPS * 32, 16), CPF_WARPS * 32, 0, (cudaStream_t)stream>>>(
      g, t, G, inv, g_rows, g_row_pitch, g_lines[0], g_lines[1], g_lines[2]);
  return check_launch("srf_cp_color_features_bwd");
}
