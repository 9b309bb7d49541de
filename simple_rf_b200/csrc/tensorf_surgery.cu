// Simple-TensoRF grid surgery ("next" row f3): the alpha-mask rebuild and the plane / line resampling that
// run on a handful of the 25 000 training iterations and invalidate every hot-path cache.
//
// Replaces (reference file:line, relative to the upstream checkout):
//   srf_alpha_grid_occupancy   src/models/SimpleTensoRF09.py:849-859 (dense grid of world points, compute_alpha :878-897:
//                              previous-mask test, normalise, get_volume_density of the VM (:1214-1239) or CP (:1043-1062) tensor,
//                              1 - exp(-sigma * step)), :862 clamp and the
//                              threshold of :866-867 applied BEFORE the pooling (max over a window >= t  <=>  any member >= t)
//   srf_alpha_grid_dilate      :864-865 (3x3x3 max-pool, stride 1, -inf padding) on the 1-bit volume, :869 (the new
//                              AlphaGridMask volume), :871-875 (per-axis projection of the occupied voxels -> new bounding box)
//   srf_pack_alpha_bits_u8     derived 1-bit cache of a bool / uint8 alpha volume (AlphaGridMask.alpha_volume, :1333)
//   srf_resample_plane         :1284-1295 F.interpolate(mode='bilinear', align_corners=True) of planes [1,C,H,W] and lines
//                              [1,C,L,1], and the window slicing of shrink_tensor (:1303-1319) as the same kernel with an
//                              identity scale (a pure copy of the window)
//
// No fp32 [Z,Y,X] volume exists at any point: density is evaluated straight into one bit per voxel (a warp = 32 consecutive
// x of one grid row, one ballot), the pooling is bit arithmetic on 3 x 3 neighbouring row words, and the bounding box comes
// from the per-axis projections of the occupied set.
#include <limits.h>

#include "tensorf_common.cuh"

namespace srf {

struct OccupancyParams {
  VmGrid grid;
  MaskParams prev;                   // previous alpha mask (alpha_bits == nullptr: none); only the alpha_* fields are read
  const float* coord[3];             // world coordinates of the grid planes along x, y, z: bb0 (1 - s) + bb1 s, s = linspace(0, 1, n)
  float bb0[3], bsize[3];            // tensor bounding box (normalisation, :763-765)
  int n[3];                          // grid size (X, Y, Z) == tensor resolution at the rebuild
  int pitch;                         // words per grid row = ceil(X / 32)
  int softplus; float offset;        // density activation (:672-676)
  float length;                      // step_size
  float threshold;                   // alpha_mask_threshold
  uint32_t* raw;                     // [Z][Y][pitch] words, bit x & 31 of word x >> 5
};

// warp = 32 consecutive x of one (y, z) row; lanes read neighbouring texels of the xy / xz planes (coalesced, channels
// last) and the same texels of the yz plane (broadcast)
__global__ void __launch_bounds__(256) alpha_occupancy_kernel(const OccupancyParams p) {
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long rows = (long long)p.n[1] * p.n[2];
  const long long total = rows * p.pitch;
  if (warp >= total) return;
  const int xw = (int)(warp % p.pitch);
  const long long row = warp / p.pitch;
  const int y = (int)(row % p.n[1]), z = (int)(row / p.n[1]);
  const int x = xw * 32 + lane_id();
  bool occ = false;
  if (x < p.n[0]) {
    const float pt[3] = {__ldg(p.coord[0] + x), __ldg(p.coord[1] + y), __ldg(p.coord[2] + z)};
    bool live = true;
    if (p.prev.alpha_bits != nullptr) live = alpha_hit(p.prev, pt);           // :879-883: only the previous mask, no box test
    float sigma = 0.f;
    if (live) {
      float pn[3];
#pragma unroll
      for (int a = 0; a < 3; ++a) pn[a] = __fadd_rn(__fmul_rn(__fdiv_rn(__fadd_rn(pt[a], -p.bb0[a]), p.bsize[a]), 2.f), -1.f);
      float feat = 0.f;
      if (p.grid.plane[0] == nullptr) {
        // CANDECOMP/PARAFAC tensor (:1043-1062): sum over components of the product of the three line factors
        int l0[3], L[3]; float w0[3], w1[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) line_coords(pn, p.grid.res, i, l0[i], L[i], w0[i], w1[i]);
        const int C = p.grid.C[0];
        for (int c = 0; c < C; c += 4) {
          const float4 a = line_fetch4(p.grid.line[0], l0[0], L[0], w0[0], w1[0], C, c);
          const float4 b = line_fetch4(p.grid.line[1], l0[1], L[1], w0[1], w1[1], C, c);
          const float4 d = line_fetch4(p.grid.line[2], l0[2], L[2], w0[2], w1[2], C, c);
          feat += (a.x * b.x * d.x + a.y * b.y * d.y) + (a.z * b.z * d.z + a.w * b.w * d.w);
        }
      } else {
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          const Bilerp b = plane_coords(pn, p.grid.res, i);
          int l0, L; float w0, w1;
          line_coords(pn, p.grid.res, i, l0, L, w0, w1);
          float part = 0.f;
          for (int c = 0; c < p.grid.C[i]; c += 4) {
            const float4 pv = plane_fetch4(p.grid.plane[i], b, p.grid.C[i], c);
            const float4 lv = line_fetch4(p.grid.line[i], l0, L, w0, w1, p.grid.C[i], c);
            part += pv.x * lv.x + pv.y * lv.y + pv.z * lv.z + pv.w * lv.w;
          }
          feat += part;
        }
      }
      if (p.softplus) { const float v = feat + p.offset; sigma = v > 20.f ? v : log1pf(expf(v)); }
      else sigma = fmaxf(feat, 0.f);
    }
    float alpha = 1.f - expf(-sigma * p.length);
    alpha = fminf(fmaxf(alpha, 0.f), 1.f);
    occ = alpha >= p.threshold;
  }
  const uint32_t word = __ballot_sync(FULL, occ);
  if (lane_id() == 0) p.raw[warp] = word;
}

// thread = one output word (32 voxels of one row): OR of the 3 x 3 neighbouring rows' words, widened by one voxel in x.
// `projection` = occupancy projected on each axis: [pitch] words for x (bit layout of a row), then Y flags, then Z flags
__global__ void __launch_bounds__(256) alpha_dilate_kernel(const uint32_t* __restrict__ raw, int X, int Y, int Z, int pitch,
                                                           uint8_t* __restrict__ volume, uint32_t* __restrict__ projection) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)Y * Z * pitch;
  if (t >= total) return;
  const int xw = (int)(t % pitch);
  const long long row = t / pitch;
  const int y = (int)(row % Y), z = (int)(row / Y);
  uint32_t mid = 0, left = 0, right = 0;
  for (int dz = -1; dz <= 1; ++dz) {
    const int zz = z + dz;
    if (zz < 0 || zz >= Z) continue;
    for (int dy = -1; dy <= 1; ++dy) {
      const int yy = y + dy;
      if (yy < 0 || yy >= Y) continue;
      const uint32_t* r = raw + ((long long)zz * Y + yy) * pitch;
      mid |= __ldg(r + xw);
      if (xw > 0) left |= __ldg(r + xw - 1);
      if (xw + 1 < pitch) right |= __ldg(r + xw + 1);
    }
  }
  uint32_t out = mid | (mid << 1) | (mid >> 1) | (left >> 31) | (right << 31);
  const int valid = min(32, X - xw * 32);
  if (valid < 32) out &= (1u << valid) - 1u;
  uint8_t* dst = volume + ((long long)z * Y + y) * X + xw * 32;
  for (int b = 0; b < valid; ++b) dst[b] = (out >> b) & 1u;
  if (out) {
    if ((__ldg(projection + xw) & out) != out) atomicOr(projection + xw, out);
    projection[pitch + y] = 1u;
    projection[pitch + Y + z] = 1u;
  }
}

__global__ void pack_alpha_u8_kernel(const uint8_t* __restrict__ vol, long long n, uint32_t* __restrict__ bits) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool on = i < n && vol[i] != 0;
  const uint32_t word = __ballot_sync(FULL, on);
  if (lane_id() == 0 && i < n) bits[i >> 5] = word;
}

// out[c][y][x] over a window (y0, x0, h, w) of src [C][H][W]: ATen's upsample_bilinear2d arithmetic with align_corners
// (scale = (in - 1) / (out - 1), source = scale * index, the far neighbour clamped to the window); an unchanged size is a copy
struct ResampleParams {
  const float* src; float* dst;
  int C, H, W;                       // source extents
  int y0, x0, h, w;                  // window of the source
  int oh, ow;                        // output extents
  float sy, sx;                      // (h - 1) / (oh - 1), (w - 1) / (ow - 1); 0 when the output extent is 1
};

__global__ void __launch_bounds__(256) resample_plane_kernel(const ResampleParams p) {
  const long long per = (long long)p.oh * p.ow;
  const long long total = per * p.C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i / per);
    const int rem = (int)(i - (long long)c * per);
    const int oy = rem / p.ow, ox = rem - oy * p.ow;
    const float* s = p.src + ((long long)c * p.H + p.y0) * p.W + p.x0;
    float v;
    if (p.oh == p.h && p.ow == p.w) {
      v = s[(long long)oy * p.W + ox];
    } else {
      const float fy = p.sy * oy, fx = p.sx * ox;
      const int iy = (int)fy, ix = (int)fx;
      const int py = iy < p.h - 1 ? 1 : 0, px = ix < p.w - 1 ? 1 : 0;
      const float ly1 = fy - iy, ly0 = 1.f - ly1, lx1 = fx - ix, lx0 = 1.f - lx1;
      const float v00 = s[(long long)iy * p.W + ix], v01 = s[(long long)iy * p.W + ix + px];
      const float v10 = s[(long long)(iy + py) * p.W + ix], v11 = s[(long long)(iy + py) * p.W + ix + px];
      v = ly0 * (lx0 * v00 + lx1 * v01) + ly1 * (lx0 * v10 + lx1 * v11);
    }
    p.dst[i] = v;
  }
}

}  // namespace srf

using namespace srf;

SRF_API int srf_alpha_grid_words(const int* resolution) {
  return (int)(((long long)(resolution[0] + 31) / 32) * resolution[1] * resolution[2]);
}

SRF_API int srf_alpha_grid_occupancy(const float* const* planes, const float* const* lines, const int* channels, const int* resolution,
                                     const float* box_min, const float* box_size, const float* coord_x, const float* coord_y,
                                     const float* coord_z, const uint32_t* prev_bits, const int* prev_res, const float* prev_box_min,
                                     const float* prev_box_size, int softplus, float density_offset, float step_size,
                                     float threshold, uint32_t* raw_words, void* stream) {
  // planes == NULL: a CANDECOMP/PARAFAC tensor (lines only, channels[0] components in each of the three lines)
  SRF_REQUIRE(lines && channels && resolution && box_min && box_size && coord_x && coord_y && coord_z && raw_words,
              "srf_alpha_grid_occupancy", "null pointer");
  SRF_REQUIRE(prev_bits == nullptr || (prev_res && prev_box_min && prev_box_size), "srf_alpha_grid_occupancy", "previous alpha box missing");
  SRF_REQUIRE(planes != nullptr || (channels[1] == channels[0] && channels[2] == channels[0]), "srf_alpha_grid_occupancy",
              "a CP tensor has the same component count in all three lines");
  OccupancyParams p{};
  for (int i = 0; i < 3; ++i) {
    p.grid.plane[i] = planes ? planes[i] : nullptr; p.grid.line[i] = lines[i]; p.grid.C[i] = channels[i]; p.grid.res[i] = resolution[i];
    SRF_REQUIRE((planes == nullptr || planes[i]) && lines[i], "srf_alpha_grid_occupancy", "null plane/line pointer");
    SRF_REQUIRE(channels[i] > 0 && !(channels[i] & 3), "srf_alpha_grid_occupancy", "channel counts must be positive multiples of 4");
    SRF_REQUIRE(resolution[i] > 0, "srf_alpha_grid_occupancy", "empty grid");
    p.bb0[i] = box_min[i]; p.bsize[i] = box_size[i]; p.n[i] = resolution[i];
  }
  p.coord[0] = coord_x; p.coord[1] = coord_y; p.coord[2] = coord_z;
  p.prev.alpha_bits = prev_bits;
  if (prev_bits) {
    for (int a = 0; a < 3; ++a) { p.prev.ab0[a] = prev_box_min[a]; p.prev.asize[a] = prev_box_size[a]; }
    p.prev.ax = prev_res[0]; p.prev.ay = prev_res[1]; p.prev.az = prev_res[2];
    SRF_REQUIRE(p.prev.ax > 0 && p.prev.ay > 0 && p.prev.az > 0 && (long long)p.prev.ax * p.prev.ay * p.prev.az < (1ll << 31),
                "srf_alpha_grid_occupancy", "alpha volume must hold fewer than 2^31 voxels");
  }
  SRF_REQUIRE((long long)p.n[0] * p.n[1] * p.n[2] < (1ll << 31), "srf_alpha_grid_occupancy", "grid must hold fewer than 2^31 voxels");
  p.pitch = (p.n[0] + 31) / 32;
  p.softplus = softplus; p.offset = density_offset; p.length = step_size; p.threshold = threshold; p.raw = raw_words;
  const long long warps = (long long)p.pitch * p.n[1] * p.n[2];
  alpha_occupancy_kernel<<<(unsigned)((warps + 7) / 8), 256, 0, (cudaStream_t)stream>>>(p);
  return check_launch("srf_alpha_grid_occupancy");
}

SRF_API int srf_alpha_grid_dilate(const uint32_t* raw_words, const int* resolution, uint8_t* volume, uint32_t* projection, void* stream) {
  SRF_REQUIRE(raw_words && resolution && volume && projection, "srf_alpha_grid_dilate", "null pointer");
  const int X = resolution[0], Y = resolution[1], Z = resolution[2];
  SRF_REQUIRE(X > 0 && Y > 0 && Z > 0, "srf_alpha_grid_dilate", "empty grid");
  const int pitch = (X + 31) / 32;
  const long long words = (long long)pitch * Y * Z;
  alpha_dilate_kernel<<<(unsigned)((words + 255) / 256), 256, 0, (cudaStream_t)stream>>>(raw_words, X, Y, Z, pitch, volume, projection);
  return check_launch("srf_alpha_grid_dilate");
}

SRF_API int srf_pack_alpha_bits_u8(const uint8_t* volume, int64_t num_voxels, uint32_t* bits, void* stream) {
  if (num_voxels == 0) return 0;
  SRF_REQUIRE(volume && bits, "srf_pack_alpha_bits_u8", "null pointer");
  const long long threads = (num_voxels + 31) / 32 * 32;
  pack_alpha_u8_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(volume, num_voxels, bits);
  return check_launch("srf_pack_alpha_bits_u8");
}

SRF_API int srf_resample_plane(const float* src, int channels, int height, int width, int y0, int x0, int h, int w, float* dst,
                               int out_height, int out_width, void* stream) {
  SRF_REQUIRE(src && dst, "srf_resample_plane", "null pointer");
  SRF_REQUIRE(channels > 0 && h > 0 && w > 0 && out_height > 0 && out_width > 0, "srf_resample_plane", "empty plane");
  SRF_REQUIRE(y0 >= 0 && x0 >= 0 && y0 + h <= height && x0 + w <= width, "srf_resample_plane", "window outside the source plane");
  ResampleParams p{};
  p.src = src; p.dst = dst; p.C = channels; p.H = height; p.W = width; p.y0 = y0; p.x0 = x0; p.h = h; p.w = w;
  p.oh = out_height; p.ow = out_width;
  p.sy = out_height > 1 ? (float)(h - 1) / (float)(out_height - 1) : 0.f;
  p.sx = out_width > 1 ? (float)(w - 1) / (float)(out_width - 1) : 0.f;
  const long long total = (long long)channels * out_height * out_width;
  const long long blocks = (total + 255) / 256;
  resample_plane_kernel<<<(unsigned)(blocks < 65535 * 16 ? blocks : 65535 * 16), 256, 0, (cudaStream_t)stream>>>(p);
  return check_launch("srf_resample_plane");
}
