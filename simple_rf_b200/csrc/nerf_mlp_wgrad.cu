// Weight gradients of the fused NeRF MLP on tcgen05 tensor cores:  dW_l = dZ_l^T X_l  and  db_l = colsum(dZ_l).
//
// Replaces what autograd derives for the nn.Linear layers of src/models/SimpleNeRF17.py:644-666 (weights,
// biases, heads).  Both operands are the 128-sample x 64-channel bf16 swizzled tile images written by the forward
// (activations X, srf_nerf_mlp_fwd with save_acts) and by the dgrad chain (pre-activation gradients dZ,
// srf_nerf_mlp_dgrad): with the reduction dimension K = samples, such an image IS the canonical MN-major
// SWIZZLE_128B UMMA operand (rows = K, 64 contiguous channels = one MN atom), so the images are bulk-copied from HBM
// into shared memory and fed to tcgen05.mma unchanged — no transposes.
//
// One CTA per (work item, split) - the SMs are dealt out to the items in proportion to the bytes they stream per tile, so
// the launch is one balanced wave: the item names the dZ images (M = 128 or 256 output channels) and the X images
// (N <= 256 input channels); the split is a contiguous range of sample tiles.  fp32 accumulators live in TMEM
// (2 x 256 columns) across the whole range and are added to the global gradient with red.global at the end.
// HBM-bound by design: 128 FLOP per byte streamed.
#include "common.cuh"
#include "tcgen05.cuh"

namespace srf {

constexpr int WG_MAX_ITEMS = 48;

struct WgradItem {            // mirrors srf_wgrad_item in include/simple_rf_b200.h
  int32_t dz_slot;            // first dZ image of the output-channel range
  int32_t dz_images;          // 2 (M = 128) or 4 (M = 256)
  int32_t x_slot;             // first X image
  int32_t x_images;           // 1..4 (N = 64 * x_images)
  int32_t out_rows;           // rows of dW actually written (<= 64 * dz_images)
  int32_t in_col0;            // first image column that maps to a weight column
  int32_t in_cols;            // number of mapped columns
  int32_t w_col0;             // weight column of image column in_col0
  int32_t w_stride;           // row pitch of dW (= in features of the layer), floats
  int32_t bias;               // 1: also accumulate db (first out_rows channels of dZ)
  int64_t dw_offset;          // float offset of dW in the gradient buffer
  int64_t db_offset;          // float offset of db
};

struct WgradParams {
  const uint8_t* acts; const uint8_t* dz;
  float* grads;
  int act_slots, dz_slots;
  int num_tiles;
  int cta_first[WG_MAX_ITEMS + 1];   // CTAs [cta_first[i], cta_first[i + 1]) share item i: the SMs are dealt out in proportion to the
                                     // bytes an item streams per tile (dz + x images), one wave, no tail
  const int* count;           // device-side row count, or nullptr: only tiles below ceil(count / 128) are accumulated
  int num_items;
  WgradItem items[WG_MAX_ITEMS];
};

constexpr int WG_THREADS = 192;           // warp 0 producer, warp 1 MMA, warps 2-5 bias sums + epilogue
constexpr int WG_STAGES = 3;
constexpr int HALF_IMAGE = 64 * 128;      // 64 samples x 64 channels bf16
constexpr int WG_STAGE_BYTES = 8 * HALF_IMAGE;

struct alignas(1024) WgradSmem {
  uint8_t stage[WG_STAGES][WG_STAGE_BYTES];      // [dz half-images (<=4) | x half-images (<=4)]
  uint64_t full[WG_STAGES], empty[WG_STAGES];
  uint64_t done;
  uint32_t tmem_base;
};

// MN-major SWIZZLE_128B operand: 64-channel atoms LBO bytes apart, 8-sample groups 1024 bytes apart
__device__ __forceinline__ uint64_t make_mn_desc(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

__global__ void __launch_bounds__(WG_THREADS, 1) nerf_mlp_wgrad_kernel(const __grid_constant__ WgradParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  WgradSmem& sm = *reinterpret_cast<WgradSmem*>(smem_raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if ((ptx::smem_u32(smem_raw) & 1023u) != 0) __trap();
  int item = 0;
  while (item + 1 < p.num_items && (int)blockIdx.x >= p.cta_first[item + 1]) ++item;
  const WgradItem& it = p.items[item];
  const int split = (int)blockIdx.x - p.cta_first[item], splits = p.cta_first[item + 1] - p.cta_first[item];
  int num_tiles = p.num_tiles;
  if (p.count != nullptr) { const int c = (*p.count + 127) / 128; num_tiles = c < num_tiles ? c : num_tiles; }
  const int t0 = (int)((long long)num_tiles * split / splits), t1 = (int)((long long)num_tiles * (split + 1) / splits);
  if (t1 == t0) return;                              // nothing to accumulate for this CTA (uniform exit, before any barrier)
  const int halves = it.dz_images >> 1;            // accumulators (M = 128 each)
  const int N = it.x_images * 64;
  const int bias_arrivals = it.bias ? 4 : 0;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < WG_STAGES; ++s) { ptx::mbar_init(&sm.full[s], 1); ptx::mbar_init(&sm.empty[s], 1 + bias_arrivals); }
    ptx::mbar_init(&sm.done, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc(&sm.tmem_base, 512);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = sm.tmem_base;
  const int num_stages_total = (t1 - t0) * 2;        // two 64-sample halves per tile
  const uint32_t stage_bytes = (uint32_t)(it.dz_images + it.x_images) * HALF_IMAGE;

  if (warp == 0) {
    if (lane == 0) {
      for (int i = 0; i < num_stages_total; ++i) {
        const int s = i % WG_STAGES, ph = (i / WG_STAGES) & 1;
        const int tile = t0 + (i >> 1), half = i & 1;
        ptx::mbar_wait(&sm.empty[s], ph ^ 1);
        ptx::mbar_arrive_expect_tx(&sm.full[s], stage_bytes);
        for (int a = 0; a < it.dz_images; ++a)
          ptx::bulk_g2s(sm.stage[s] + a * HALF_IMAGE,
                        p.dz + ((size_t)tile * p.dz_slots + it.dz_slot + a) * (2 * HALF_IMAGE) + half * HALF_IMAGE, HALF_IMAGE, &sm.full[s]);
        for (int b = 0; b < it.x_images; ++b)
          ptx::bulk_g2s(sm.stage[s] + (4 + b) * HALF_IMAGE,
                        p.acts + ((size_t)tile * act_tile_images(p.act_slots) + it.x_slot + b) * (2 * HALF_IMAGE) + half * HALF_IMAGE, HALF_IMAGE, &sm.full[s]);
      }
    }
  } else if (warp == 1) {
    // MMA issuer: converged warp, one elected lane issues.  A and B both MN-major (bits 15, 16), fp32 accumulate, bf16 operands
    const uint32_t idesc = ptx::make_idesc_bf16(128, (uint32_t)N) | (1u << 15) | (1u << 16);
    uint32_t s = 0, ph = 0;
    for (int i = 0; i < num_stages_total; ++i) {
      ptx::mbar_wait(&sm.full[s], ph);
      ptx::tc_fence_after();
      const uint32_t base = ptx::smem_u32(sm.stage[s]);
      const uint32_t issue = ptx::elect_one();
#pragma unroll
      for (int k = 0; k < 4; ++k) {                 // 16 samples per MMA
        const uint64_t b_desc = make_mn_desc(base + 4 * HALF_IMAGE + k * 2048, HALF_IMAGE);
        ptx::umma_bf16_if(issue, tmem, make_mn_desc(base + k * 2048, HALF_IMAGE), b_desc, idesc, (i == 0 && k == 0) ? 0u : 1u);
        if (halves > 1)
          ptx::umma_bf16_if(issue, tmem + 256, make_mn_desc(base + 2 * HALF_IMAGE + k * 2048, HALF_IMAGE), b_desc, idesc,
                            (i == 0 && k == 0) ? 0u : 1u);
      }
      ptx::umma_commit_if(issue, &sm.empty[s]);
      if (++s == WG_STAGES) { s = 0; ph ^= 1; }
    }
    ptx::umma_commit_if(ptx::elect_one(), &sm.done);
  } else {
    // ------------------------------------------------------------ bias column sums, then the TMEM -> global epilogue
    const int t = threadIdx.x - 64;                  // 0..127
    float bsum[2] = {0.f, 0.f};
    if (it.bias) {
      for (int i = 0; i < num_stages_total; ++i) {
        const int s = i % WG_STAGES, ph = (i / WG_STAGES) & 1;
        ptx::mbar_wait(&sm.full[s], ph);
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int ch = t + 128 * j;                // output channel 0..255
          if (ch < it.dz_images * 64) {
            const uint8_t* img = sm.stage[s] + (ch >> 6) * HALF_IMAGE;
            const int c = ch & 63;
            float acc = 0.f;
            for (int r = 0; r < 64; ++r)
              acc += __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(img + ptx::sw128_offset(r, c >> 3) + ((c & 7) << 1)));
            bsum[j] += acc;
          }
        }
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&sm.empty[s]);
      }
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int ch = t + 128 * j;
        if (ch < it.out_rows && num_stages_total > 0) atomicAdd(p.grads + it.db_offset + ch, bsum[j]);
      }
    }
    ptx::mbar_wait(&sm.done, 0);
    ptx::tc_fence_after();
    if (num_stages_total > 0) {
      const int quarter = warp & 3;                  // TMEM lane quarter of this warp
      for (int h = 0; h < halves; ++h) {
        const int m = h * 128 + quarter * 32 + lane; // output channel = TMEM lane
        for (int c0 = 0; c0 < N; c0 += 32) {
          uint32_t v[32];
          ptx::tmem_ld32(tmem + ((uint32_t)(quarter * 32) << 16) + h * 256 + c0, v);
          ptx::tmem_ld_wait(v);
          if (m < it.out_rows) {
            float* row = p.grads + it.dw_offset + (size_t)m * it.w_stride;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int c = c0 + j - it.in_col0;
              if (c >= 0 && c < it.in_cols) atomicAdd(row + it.w_col0 + c, __uint_as_float(v[j]));
            }
          }
        }
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem, 512);
  }
}

}  // namespace srf

using namespace srf;

SRF_API int srf_nerf_mlp_wgrad(const void* items, int num_items, const void* acts, int act_slots, const void* dz, int dz_slots,
                               int64_t num_tiles, const int* count, float* grads, void* stream) {
  if (num_tiles == 0 || num_items == 0) return 0;
  SRF_REQUIRE(items && acts && dz && grads, "srf_nerf_mlp_wgrad", "null pointer");
  SRF_REQUIRE(num_items <= WG_MAX_ITEMS, "srf_nerf_mlp_wgrad", "too many work items");
  WgradParams p{};
  p.acts = reinterpret_cast<const uint8_t*>(acts); p.dz = reinterpret_cast<const uint8_t*>(dz); p.grads = grads;
  p.act_slots = act_slots; p.dz_slots = dz_slots; p.num_tiles = (int)num_tiles; p.num_items = num_items; p.count = count;
  const WgradItem* src = reinterpret_cast<const WgradItem*>(items);
  for (int i = 0; i < num_items; ++i) {
    const WgradItem& it = src[i];
    SRF_REQUIRE((it.dz_images == 2 || it.dz_images == 4) && it.x_images >= 1 && it.x_images <= 4, "srf_nerf_mlp_wgrad", "bad image counts");
    SRF_REQUIRE(it.dz_slot >= 0 && it.dz_slot + it.dz_images <= dz_slots && it.x_slot >= 0 && it.x_slot + it.x_images <= act_slots,
                "srf_nerf_mlp_wgrad", "image slot out of range");
    SRF_REQUIRE(it.out_rows >= 1 && it.out_rows <= 64 * it.dz_images && it.in_col0 >= 0 && it.in_cols >= 1 &&
                    it.in_col0 + it.in_cols <= 64 * it.x_images, "srf_nerf_mlp_wgrad", "bad row / column mapping");
    p.items[i] = it;
  }
  // deal the SMs out in proportion to the bytes per tile of every item (largest-remainder rounding, at least one CTA each)
  int ctas = sm_count();
  if ((long long)ctas > (long long)num_items * num_tiles) ctas = (int)((long long)num_items * num_tiles);
  if (ctas < num_items) ctas = num_items;
  int weight[WG_MAX_ITEMS], share[WG_MAX_ITEMS], total_w = 0, used = 0;
  for (int i = 0; i < num_items; ++i) { weight[i] = p.items[i].dz_images + p.items[i].x_images; total_w += weight[i]; }
  for (int i = 0; i < num_items; ++i) {
    share[i] = (int)((long long)ctas * weight[i] / total_w);
    if (share[i] < 1) share[i] = 1;
    if (share[i] > num_tiles) share[i] = (int)num_tiles;
    used += share[i];
  }
  for (int guard = 0; used != ctas && guard < 4 * ctas; ++guard) {      // hand the remainder to / take the excess from the
    int best = -1;                                                      // items with the most / least work per CTA
    for (int i = 0; i < num_items; ++i) {
      if (used < ctas ? share[i] >= num_tiles : share[i] <= 1) continue;
      if (best < 0 || (used < ctas ? (long long)weight[i] * share[best] > (long long)weight[best] * share[i]
                                   : (long long)weight[i] * share[best] < (long long)weight[best] * share[i]))
        best = i;
    }
    if (best < 0) break;
    share[best] += used < ctas ? 1 : -1;
    used += used < ctas ? 1 : -1;
  }
  p.cta_first[0] = 0;
  for (int i = 0; i < num_items; ++i) p.cta_first[i + 1] = p.cta_first[i] + share[i];
  const int grid = p.cta_first[num_items];
  const size_t smem = sizeof(WgradSmem);
  static unsigned long long configured = 0ull;      // one bit per device
  if (first_use_on_this_device(configured)) {
    cudaError_t e = cudaFuncSetAttribute(nerf_mlp_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail("srf_nerf_mlp_wgrad", cudaGetErrorString(e));
  }
  nerf_mlp_wgrad_kernel<<<grid, WG_THREADS, smem, (cudaStream_t)stream>>>(p);
  return check_launch("srf_nerf_mlp_wgrad");
}

SRF_API int srf_wgrad_item_bytes(void) { return (int)sizeof(WgradItem); }
