// Fused test-time ray march of one Simple-TensoRF VM tensor (NDC): per ray, front to back, in ONE kernel
//   sample depth from the shared ladder -> point -> box test -> alphaMask test -> VM density gather -> alpha ->
//   transmittance (warp scan) -> weights -> acc / depth / depth variance (NDC and world) -> surface test (w > threshold)
// i.e. src/models/SimpleTensoRF09.py:701-726 + :767-819 of LowRankTensor.forward without any [R,S] intermediate: the
// unfused path (kept for training and for `retraw`) materialises the depths z[R,S] (they are ONE ladder at test time), a byte
// mask, a compacted index list, dense sigma and dense weights, and walks them in ~8 launches.  Here the only per-sample
// traffic is the list of surface samples (8 % of the samples on the benchmark scene), written in ray order.
//
//   srf_tensorf_march          warp per ray; chunks of 32 consecutive samples are TESTED (box, occupancy — one bit of the
//                              corner-OR volume away from voxel boundaries, the exact trilinear test next to them) and the valid ones queued; density + compositing run on full batches of 32
//                              valid samples; a ray stops once its transmittance is below 1e-7 (everything behind
//                              contributes < 1e-7 to any map; the surface threshold is 1e-4)
//   srf_alpha_corner_or_bits   derived cache: per cell of the alpha grid, the OR of its 8 corner bits
//   srf_tensorf_march_compact  per-ray surface lists -> one flat list in row-major (ray, sample) order — exactly the order
//                              of the reference's boolean-mask indexing (:1248) — + per-ray offsets
//   srf_ray_accumulate         rgb_map[r] = sum_k w_k rgb_k over the ray's surface samples (+ white background), fixed order
//
// The arithmetic of point, box test, occupancy test, normalisation and bilinear weights is shared with tensorf.cu
// (tensorf_common.cuh), so validity and surface sets equal the unfused path's.
#include "tensorf_common.cuh"

namespace srf {

constexpr int MARCH_WARPS = 8;
constexpr int MARCH_MAX_SMEM_LADDER = 4096;

struct MarchParams {
  MaskParams m;                 // box + alpha volume (rays / z / mask members unused)
  VmGrid t;                     // density planes / lines
  const float* o_ndc; const float* d_ndc; const float* rays_o; const float* rays_d;
  const float* ladder;          // [S] sample depths shared by all rays
  float bsize[3];               // box size, as the host stores it (bounding_box_size buffer)
  const uint32_t* alpha_or8;    // nullable: corner-OR volume of the alpha bits, (ax+1) x (ay+1) x (az+1) (srf_alpha_corner_or_bits)
  float cscale[3];              // (alpha resolution - 1) / alpha box size: world offset -> voxel coordinate (approximate)
  int R, S, softplus;
  float offset, distance_scale, threshold;
  float* acc; float* depth; float* depth_var; float* depth_ndc; float* depth_var_ndc;   // [R]
  int* ray_count;               // [R] surface samples of the ray
  int* entry_sample;            // [R,S]: the first ray_count[r] entries of row r are the sample indices, ascending
  float* entry_weight;          // [R,S]
};

__device__ __forceinline__ float march_rcp(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float march_sqrt(float x) {
  float y;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// sigma of one point (the arithmetic of vm_density_fwd_kernel)
__device__ __forceinline__ float density_at(const MarchParams& p, const float (&pt)[3]) {
  float pn[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) pn[a] = __fadd_rn(__fmul_rn(__fdiv_rn(__fadd_rn(pt[a], -p.m.bb0[a]), p.bsize[a]), 2.f), -1.f);
  float feat = 0.f;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const Bilerp b = plane_coords(pn, p.t.res, i);
    int l0, L; float w0, w1;
    line_coords(pn, p.t.res, i, l0, L, w0, w1);
    float part = 0.f;
    for (int c = 0; c < p.t.C[i]; c += 4) {
      const float4 pv = plane_fetch4(p.t.plane[i], b, p.t.C[i], c);
      const float4 lv = line_fetch4(p.t.line[i], l0, L, w0, w1, p.t.C[i], c);
      part += pv.x * lv.x + pv.y * lv.y + pv.z * lv.z + pv.w * lv.w;
    }
    feat += part;
  }
  if (p.softplus) { const float x = feat + p.offset; return x > 20.f ? x : log1pf(expf(x)); }
  return fmaxf(feat, 0.f);
}

// running weighted mean / second moment, merged chunk by chunk (Chan's parallel update): one pass, no cancellation
struct Moments {
  float W, mean, M2;
  __device__ __forceinline__ void merge(float s0, float mc, float m2c) {
    const float Wn = W + s0;
    const float d = mc - mean;
    const float f = s0 * march_rcp(Wn);
    mean += d * f;
    M2 += m2c + d * d * W * f;
    W = Wn;
  }
};

// Fast occupancy test on the corner-OR volume.  Away from voxel boundaries the exact trilinear test of alpha_hit() reduces to
// "some in-range corner of the cell holds a 1": both weights of every axis are strictly positive, so every corner's weight
// product is.  The corner-OR volume stores exactly that per cell (cell i0 = -1 .. dim-1 per axis, index i0 + 1), and the cell of a
// point follows from a cheap approximate coordinate whenever its fractional part is at least EDGE away from 0 and 1 (the
// approximate and the exact fp32 coordinate differ by < 1e-3 voxel for any resolution the < 2^31-voxel bound admits).
// Returns +1 hit, 0 miss, -1 undecided (within EDGE of a boundary: run the exact test).
__device__ __forceinline__ int corner_or_test(const MarchParams& p, const float (&pt)[3]) {
  constexpr float EDGE = 1.f / 256.f;
  const int dims[3] = {p.m.ax, p.m.ay, p.m.az};
  int c[3];
  bool outside = false;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const float f = (pt[a] - p.m.ab0[a]) * p.cscale[a];
    const float fl = floorf(f);
    const float t = f - fl;
    if (!(t > EDGE && t < 1.f - EDGE)) return -1;                // also catches NaN
    const int i = (int)fl;
    outside = outside || i < -1 || i > dims[a] - 1;
    c[a] = i + 1;
  }
  if (outside) return 0;
  const int v = (c[2] * (p.m.ay + 1) + c[1]) * (p.m.ax + 1) + c[0];
  return (int)((__ldg(p.alpha_or8 + (v >> 5)) >> (v & 31)) & 1u);
}

// Warp per ray.  Two alternating phases: FILL tests chunks of 32 consecutive ladder samples (box, occupancy) and appends the valid sample indices to a per-warp queue; DRAIN takes 32 queued samples (lane = one VALID
// sample: the density gather and the compositing arithmetic run with every lane busy), evaluates density, alpha with the
// sample's own ladder interval, the transmittance scan and the per-ray sums.  Samples that fail the tests have alpha == 0 and
// multiply the transmittance by fl(1 + 1e-10) == 1: skipping them changes nothing.  A ray stops once T < 1e-7.
__global__ void __launch_bounds__(MARCH_WARPS * 32, 3) tensorf_march_kernel(const MarchParams p) {
  __shared__ float s_ladder[MARCH_MAX_SMEM_LADDER];
  __shared__ unsigned short s_queue[MARCH_WARPS][64];
  const bool ladder_in_smem = p.S <= MARCH_MAX_SMEM_LADDER;
  if (ladder_in_smem)
    for (int i = threadIdx.x; i < p.S; i += blockDim.x) s_ladder[i] = p.ladder[i];
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const unsigned lt = (1u << lane) - 1u;
  unsigned short* queue = s_queue[warp];
  auto depth_of = [&](int s) { return ladder_in_smem ? s_ladder[s] : __ldg(p.ladder + s); };
  for (int r = blockIdx.x * MARCH_WARPS + warp; r < p.R; r += gridDim.x * MARCH_WARPS) {
    float o[3], d[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) { o[a] = __ldg(p.o_ndc + (size_t)r * 3 + a); d[a] = __ldg(p.d_ndc + (size_t)r * 3 + a); }
    // per-ray constants of the compositing (csrc/composite.cu: ray_consts / ndc_to_world)
    const float ns = march_sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]) * p.distance_scale;
    const float oz = __ldg(p.rays_o + (size_t)r * 3 + 2), wz = __ldg(p.rays_d + (size_t)r * 3 + 2);
    const float iw = march_rcp(wz);
    const float tn = -(1.f + oz) * iw;
    const float A = (oz + tn * wz) * iw;
    float T = 1.f;
    Moments mn{0.f, 0.f, 0.f}, mw{0.f, 0.f, 0.f};
    int kcount = 0, queued = 0, c0 = 0;
    const size_t row = (size_t)r * p.S;
    while (true) {
      // ---- fill: until a full batch is queued or the ladder is exhausted
      while (queued < 32 && c0 < p.S) {
        const int s = c0 + lane;
        bool ok = s < p.S;
        const float z = ok ? depth_of(s) : 1.f;
        float pt[3];
#pragma unroll
        for (int a = 0; a < 3; ++a) {
          pt[a] = __fadd_rn(o[a], __fmul_rn(d[a], z));
          ok = ok && (p.m.bb0[a] <= pt[a]) && (pt[a] <= p.m.bb1[a]);
        }
        if (ok && p.m.alpha_bits != nullptr) {
          const int fast = p.alpha_or8 != nullptr ? corner_or_test(p, pt) : -1;
          ok = fast >= 0 ? fast != 0 : alpha_hit(p.m, pt);
        }
        const unsigned bal = __ballot_sync(FULL, ok);
        if (ok) queue[queued + __popc(bal & lt)] = (unsigned short)s;
        queued += __popc(bal);
        c0 += 32;
      }
      __syncwarp();
      const int n = min(queued, 32);
      if (n == 0) break;
      // ---- drain: one batch of valid samples, in ladder order
      const bool valid = lane < n;
      const int s = valid ? (int)queue[lane] : p.S - 1;
      const float z = depth_of(s);
      float pt[3];
#pragma unroll
      for (int a = 0; a < 3; ++a) pt[a] = __fadd_rn(o[a], __fmul_rn(d[a], z));
      const float sigma = valid ? density_at(p, pt) : 0.f;
      const float zn = s + 1 < p.S ? depth_of(s + 1) : 1.f;            // last interval ends at inf_depth = 1 in NDC (:781)
      const float al = valid ? 1.f - __expf(-sigma * ((zn - z) * ns)) : 0.f;
      const float q = valid ? (1.f - al + 1e-10f) : 1.f;
      float incl = q;
#pragma unroll
      for (int sh = 1; sh < 32; sh <<= 1) {
        const float tq = __shfl_up_sync(FULL, incl, sh);
        if (lane >= sh) incl *= tq;
      }
      float excl = __shfl_up_sync(FULL, incl, 1);
      if (lane == 0) excl = 1.f;
      const float w = al * (T * excl);
      T *= __shfl_sync(FULL, incl, 31);
      // per-ray sums
      const float eps = z == 1.f ? 1e-3f : 0.f;
      const float zw = fmaf(A, march_rcp(__fadd_rn(__fadd_rn(1.f, -z), eps)) - 1.f, tn);
      const float s0 = warp_sum(w);
      if (s0 > 0.f) {
        const float inv0 = march_rcp(s0);
        const float mcn = warp_sum(w * z) * inv0, mcw = warp_sum(w * zw) * inv0;
        const float dn = z - mcn, dw = zw - mcw;
        mn.merge(s0, mcn, warp_sum(w * dn * dn));
        mw.merge(s0, mcw, warp_sum(w * dw * dw));
      }
      // surface samples of this batch, in sample order
      const bool surf = valid && w > p.threshold;
      const unsigned bal = __ballot_sync(FULL, surf);
      if (surf) {
        const size_t e = row + kcount + __popc(bal & lt);
        p.entry_sample[e] = s;
        p.entry_weight[e] = w;
      }
      kcount += __popc(bal);
      if (T < 1e-7f) break;
      // ---- keep what the batch did not take
      const int left = queued - n;
      const unsigned short keep = lane < left ? queue[32 + lane] : (unsigned short)0;
      __syncwarp();
      if (lane < left) queue[lane] = keep;
      queued = left;
      __syncwarp();
    }
    if (lane == 0) {
      const float acc = mn.W;
      const float inv = march_rcp(acc + 1e-6f);
      const float dn = mn.mean * acc * inv, dw = mw.mean * acc * inv;          // sum(w z) / (acc + 1e-6)  (:797, :803)
      p.acc[r] = acc;
      p.depth_ndc[r] = dn;
      p.depth_var_ndc[r] = mn.M2 + acc * (mn.mean - dn) * (mn.mean - dn);     // sum w (z - depth)^2 about the biased mean
      p.depth[r] = dw;
      p.depth_var[r] = mw.M2 + acc * (mw.mean - dw) * (mw.mean - dw);
      p.ray_count[r] = kcount;
    }
    __syncwarp();
  }
}

// corner-OR volume: bit (x, y, z), 0 <= x <= ax etc., = OR of the alpha bits at (x - 1 .. x, y - 1 .. y, z - 1 .. z) that are in range
__global__ void alpha_or8_kernel(const uint32_t* __restrict__ bits, int ax, int ay, int az, uint32_t* __restrict__ or8) {
  const long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)(ax + 1) * (ay + 1) * (az + 1);
  bool any = false;
  if (v < total) {
    const int x = (int)(v % (ax + 1)), y = (int)((v / (ax + 1)) % (ay + 1)), z = (int)(v / ((long long)(ax + 1) * (ay + 1)));
    for (int dz = -1; dz <= 0; ++dz)
      for (int dy = -1; dy <= 0; ++dy)
        for (int dx = -1; dx <= 0; ++dx) {
          const int xx = x + dx, yy = y + dy, zz = z + dz;
          if (xx < 0 || yy < 0 || zz < 0 || xx >= ax || yy >= ay || zz >= az) continue;
          const int f = (zz * ay + yy) * ax + xx;
          any = any || ((__ldg(bits + (f >> 5)) >> (f & 31)) & 1u);
        }
  }
  const uint32_t word = __ballot_sync(FULL, any);
  if ((threadIdx.x & 31) == 0 && v < ((total + 31) & ~31ll)) or8[v >> 5] = word;
}

// exclusive scan of the per-ray counts inside blocks of 1024 rays + block totals
__global__ void __launch_bounds__(1024) ray_count_scan_kernel(const int* __restrict__ counts, int R, int* __restrict__ local_off,
                                                              int* __restrict__ block_tot) {
  __shared__ int s_warp[32];
  const int i = blockIdx.x * 1024 + threadIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int v = i < R ? counts[i] : 0;
  int incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(FULL, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    int w = s_warp[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(FULL, w, o);
      if (lane >= o) w += t;
    }
    s_warp[lane] = w;
  }
  __syncthreads();
  const int before = warp == 0 ? 0 : s_warp[warp - 1];
  if (i < R) local_off[i] = before + incl - v;
  if (threadIdx.x == 1023) block_tot[blockIdx.x] = before + incl;
}

// flat list in (ray, sample) order: idx[off[r] + k] = r S + sample_k, weight likewise; off[r] and the total are written too
__global__ void __launch_bounds__(256) march_compact_kernel(const int* __restrict__ counts, const int* __restrict__ local_off,
                                                            const int* __restrict__ block_tot, int nblocks, int R, int S,
                                                            const int* __restrict__ entry_sample, const float* __restrict__ entry_weight,
                                                            int* __restrict__ ray_off, int* __restrict__ idx, float* __restrict__ weight,
                                                            int* __restrict__ total) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (blockIdx.x == 0 && warp == 0) {
    int t = 0;
    for (int b = lane; b < nblocks; b += 32) t += block_tot[b];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(FULL, t, o);
    if (lane == 0) total[0] = t;
  }
  for (int r = blockIdx.x * 8 + warp; r < R; r += gridDim.x * 8) {
    const int blk = r >> 10;
    int base = 0;
    for (int b = lane; b < blk; b += 32) base += block_tot[b];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) base += __shfl_xor_sync(FULL, base, o);
    const int off = base + local_off[r];
    const int n = counts[r];
    if (lane == 0) ray_off[r] = off;
    const size_t row = (size_t)r * S;
    for (int k = lane; k < n; k += 32) {
      idx[off + k] = (int)(row + entry_sample[row + k]);
      weight[off + k] = entry_weight[row + k];
    }
  }
}

// rgb_map[r] = sum over the ray's surface samples of w * rgb (+ 1 - acc on a white background), lane-strided partial sums
// combined by a fixed butterfly: deterministic
__global__ void __launch_bounds__(256) ray_accumulate_kernel(const float* __restrict__ rgb_rows, const float* __restrict__ weight,
                                                             const int* __restrict__ ray_off, const int* __restrict__ counts,
                                                             const float* __restrict__ acc, int R, int white, float* __restrict__ rgb_map) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int r = blockIdx.x * 8 + warp; r < R; r += gridDim.x * 8) {
    const int off = ray_off[r], n = counts[r];
    float c0 = 0.f, c1 = 0.f, c2 = 0.f;
    for (int k = lane; k < n; k += 32) {
      const float w = weight[off + k];
      const float* c = rgb_rows + (size_t)(off + k) * 3;
      c0 += w * c[0]; c1 += w * c[1]; c2 += w * c[2];
    }
    c0 = warp_sum(c0); c1 = warp_sum(c1); c2 = warp_sum(c2);
    if (lane == 0) {
      if (white) { const float bg = 1.f - acc[r]; c0 += bg; c1 += bg; c2 += bg; }
      rgb_map[(size_t)r * 3 + 0] = c0; rgb_map[(size_t)r * 3 + 1] = c1; rgb_map[(size_t)r * 3 + 2] = c2;
    }
  }
}

}  // namespace srf

using namespace srf;

SRF_API int srf_tensorf_march(const float* rays_o_ndc, const float* rays_d_ndc, const float* rays_o, const float* rays_d,
                              const float* ladder, int64_t num_rays, int num_samples, const float* bbox, const float* box_size,
                              const uint32_t* alpha_bits, const uint32_t* alpha_corner_or, const int* alpha_res, const float* alpha_box_min,
                              const float* alpha_box_size, const float* const* planes, const float* const* lines, const int* channels, const int* resolution,
                              int softplus, float density_offset, float distance_scale, float weight_threshold,
                              float* acc, float* depth, float* depth_var, float* depth_ndc, float* depth_var_ndc,
                              int* ray_count, int* entry_sample, float* entry_weight, void* stream) {
  if (num_rays == 0) return 0;
  SRF_REQUIRE(rays_o_ndc && rays_d_ndc && rays_o && rays_d && ladder && bbox && box_size, "srf_tensorf_march", "null pointer");
  SRF_REQUIRE(acc && depth && depth_var && depth_ndc && depth_var_ndc && ray_count && entry_sample && entry_weight, "srf_tensorf_march",
              "null output pointer");
  SRF_REQUIRE(alpha_bits == nullptr || (alpha_res && alpha_box_min && alpha_box_size), "srf_tensorf_march", "alpha box missing");
  SRF_REQUIRE(num_samples > 0 && num_rays * (int64_t)num_samples < (1ll << 31), "srf_tensorf_march", "more than 2^31 samples in one call");
  MarchParams p{};
  p.m.alpha_bits = alpha_bits;
  for (int a = 0; a < 3; ++a) { p.m.bb0[a] = bbox[a]; p.m.bb1[a] = bbox[3 + a]; p.bsize[a] = box_size[a]; }
  if (alpha_bits) {
    for (int a = 0; a < 3; ++a) { p.m.ab0[a] = alpha_box_min[a]; p.m.asize[a] = alpha_box_size[a]; }
    p.m.ax = alpha_res[0]; p.m.ay = alpha_res[1]; p.m.az = alpha_res[2];
    SRF_REQUIRE(p.m.ax > 0 && p.m.ay > 0 && p.m.az > 0 && (long long)p.m.ax * p.m.ay * p.m.az < (1ll << 31), "srf_tensorf_march",
                "alpha volume must hold fewer than 2^31 voxels");
    p.alpha_or8 = alpha_corner_or;
    SRF_REQUIRE((long long)(p.m.ax + 1) * (p.m.ay + 1) * (p.m.az + 1) < (1ll << 31), "srf_tensorf_march", "alpha volume too large");
    for (int a = 0; a < 3; ++a) p.cscale[a] = (float)(alpha_res[a] - 1) / alpha_box_size[a];
  }
  SRF_REQUIRE(num_samples <= 65535, "srf_tensorf_march", "more than 65535 samples per ray");
  for (int i = 0; i < 3; ++i) {
    SRF_REQUIRE(planes[i] && lines[i], "srf_tensorf_march", "null plane/line pointer");
    SRF_REQUIRE(channels[i] > 0 && (channels[i] & 3) == 0, "srf_tensorf_march", "channel counts must be positive multiples of 4");
    p.t.plane[i] = planes[i]; p.t.line[i] = lines[i]; p.t.C[i] = channels[i]; p.t.res[i] = resolution[i];
  }
  p.o_ndc = rays_o_ndc; p.d_ndc = rays_d_ndc; p.rays_o = rays_o; p.rays_d = rays_d; p.ladder = ladder;
  p.R = (int)num_rays; p.S = num_samples; p.softplus = softplus; p.offset = density_offset;
  p.distance_scale = distance_scale; p.threshold = weight_threshold;
  p.acc = acc; p.depth = depth; p.depth_var = depth_var; p.depth_ndc = depth_ndc; p.depth_var_ndc = depth_var_ndc;
  p.ray_count = ray_count; p.entry_sample = entry_sample; p.entry_weight = entry_weight;
  long long blocks = (num_rays + MARCH_WARPS - 1) / MARCH_WARPS;
  const long long cap = (long long)sm_count() * 24;      // short blocks: the rays of a frame differ a lot in cost
  if (blocks > cap) blocks = cap;
  tensorf_march_kernel<<<(int)blocks, MARCH_WARPS * 32, 0, (cudaStream_t)stream>>>(p);
  return check_launch("srf_tensorf_march");
}

SRF_API int srf_alpha_corner_or_words(const int* alpha_res) {
  return (int)(((long long)(alpha_res[0] + 1) * (alpha_res[1] + 1) * (alpha_res[2] + 1) + 31) / 32);
}

SRF_API int srf_alpha_corner_or_bits(const uint32_t* alpha_bits, const int* alpha_res, uint32_t* corner_or, void* stream) {
  SRF_REQUIRE(alpha_bits && alpha_res && corner_or, "srf_alpha_corner_or_bits", "null pointer");
  const int ax = alpha_res[0], ay = alpha_res[1], az = alpha_res[2];
  SRF_REQUIRE(ax > 0 && ay > 0 && az > 0 && (long long)(ax + 1) * (ay + 1) * (az + 1) < (1ll << 31), "srf_alpha_corner_or_bits",
              "bad alpha resolution");
  const long long threads = (long long)srf_alpha_corner_or_words(alpha_res) * 32;
  alpha_or8_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(alpha_bits, ax, ay, az, corner_or);
  return check_launch("srf_alpha_corner_or_bits");
}

SRF_API int srf_tensorf_march_blocks(int64_t num_rays) { return (int)((num_rays + 1023) / 1024); }

SRF_API int srf_tensorf_march_compact(const int* ray_count, int64_t num_rays, int num_samples, const int* entry_sample,
                                      const float* entry_weight, int* scratch, int* ray_offset, int* indices, float* weights,
                                      int* count, void* stream) {
  SRF_REQUIRE(count, "srf_tensorf_march_compact", "null pointer");
  if (num_rays == 0) return cudaMemsetAsync(count, 0, sizeof(int), (cudaStream_t)stream) == cudaSuccess ? 0 : fail("srf_tensorf_march_compact", "memset");
  SRF_REQUIRE(ray_count && entry_sample && entry_weight && scratch && ray_offset && indices && weights, "srf_tensorf_march_compact", "null pointer");
  const int nb = srf_tensorf_march_blocks(num_rays);
  int* local_off = scratch;                 // [num_rays]
  int* block_tot = scratch + num_rays;      // [nb]
  ray_count_scan_kernel<<<nb, 1024, 0, (cudaStream_t)stream>>>(ray_count, (int)num_rays, local_off, block_tot);
  long long blocks = (num_rays + 7) / 8;
  const long long cap = (long long)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  march_compact_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(ray_count, local_off, block_tot, nb, (int)num_rays, num_samples,
                                                                      entry_sample, entry_weight, ray_offset, indices, weights, count);
  return check_launch("srf_tensorf_march_compact");
}

SRF_API int srf_ray_accumulate(const float* rgb_rows, const float* weights, const int* ray_offset, const int* ray_count,
                               const float* acc, int64_t num_rays, int white_bkgd, float* rgb_map, void* stream) {
  if (num_rays == 0) return 0;
  SRF_REQUIRE(rgb_rows && weights && ray_offset && ray_count && acc && rgb_map, "srf_ray_accumulate", "null pointer");
  long long blocks = (num_rays + 7) / 8;
  const long long cap = (long long)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  ray_accumulate_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(rgb_rows, weights, ray_offset, ray_count, acc, (int)num_rays,
                                                                       white_bkgd, rgb_map);
  return check_launch("srf_ray_accumulate");
}
