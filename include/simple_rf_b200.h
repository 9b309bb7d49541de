/* simple_rf_b200 — C ABI of the B200 (sm_100a) per-ray rendering hot path of Simple-RF.
 *
 * The reference (NagabhushanSN95/Simple-RF) is pure PyTorch and has no FFI of its own; the seam a
 * maintainer binds to is the set of torch-op sequences inside its model classes.  Every entry point
 * below names the reference lines it replaces (paths relative to the upstream checkout).
 *
 * Conventions
 *  - every pointer is a DEVICE pointer to densely packed row-major data, fp32 unless stated;
 *  - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, nothing synchronises;
 *  - return value 0 = success; otherwise srf_last_error() describes the failure (per host thread);
 *  - "nullable" outputs may be NULL to skip that output;
 *  - no CPU fallback exists: without a CUDA device every call fails with a CUDA error.
 */
#ifndef SIMPLE_RF_B200_H
#define SIMPLE_RF_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

const char* srf_last_error(void);
int srf_abi_version(void);

/* Ray generation + NDC warp + unit view directions, one fused kernel.
 * Replaces src/utils/CommonUtils04.py:73-95 (get_rays_tr), :120-138 (get_ndc_rays_tr), :147-149
 * (get_view_dirs_tr) and the x-flip of src/models/SimpleTensoRF09.py:205-207.
 *   pixel_id int32 [R,3] = (view, x, y);  k_inv [F,9] per-view inverse intrinsics;  c2w [F,16];
 *   focal [F,2] = (fx, fy);  near, two_near = (float)near, (float)(2.0*near) as the reference's
 *   python-double scalars round;  half_pixel: +0.5 px (`mip_nerf_used`);  flip_x: TensoRF x-flip;
 *   viewdirs_from_ndc: normalise the NDC direction (SimpleTensoRF09.py:238) instead of the world one.
 *   outputs [R,3]; rays_o_ndc / rays_d_ndc are required when ndc != 0. */
int srf_raygen(const int32_t* pixel_id, int64_t num_rays, const float* k_inv, const float* c2w,
               const float* focal, int num_views, int height, int width, float near, float two_near,
               int half_pixel, int flip_x, int ndc, int viewdirs_from_ndc, float* rays_o, float* rays_d,
               float* rays_o_ndc, float* rays_d_ndc, float* view_dirs, void* stream);

/* Stratified depths.  Replaces src/models/SimpleNeRF17.py:347-357 (== SimpleTensoRF09.py:371-381).
 *   ladder [S] = the un-jittered depths (:341-345);  jitter [R,S] = the reference's torch.rand draws
 *   (parity mode), or NULL with use_philox=1 (in-kernel Philox4x32-10, NOT seed-compatible with the
 *   reference) or NULL with use_philox=0 (eval: z = ladder broadcast).  z [R,S]. */
int srf_stratified_z(const float* ladder, int num_samples, int64_t num_rays, const float* jitter,
                     int use_philox, uint64_t seed, float* z, void* stream);

/* Box-march depths of Simple-TensoRF without NDC.  Replaces src/models/SimpleTensoRF09.py:388-400: entry distance of the ray into
 * the tensor's bounding box (zero direction components replaced by 1e-6), clamped to [near, far], then
 *   z[r, s] = t_entry[r] + step_size * (s + jitter[r])      (jitter [R] = the reference's one torch.rand per ray, or NULL).
 * bbox HOST [2,3] = [min xyz | max xyz]; rays_o / rays_d are the WORLD rays; z [R,S].  Bit-exact w.r.t. the reference's fp32 ops. */
int srf_box_march_z(const float* rays_o, const float* rays_d, int64_t num_rays, int num_samples, const float* bbox,
                    float near, float far, float step_size, const float* jitter, float* z, void* stream);

/* Hierarchical resampling + merge.  Replaces src/models/SimpleNeRF17.py:360-371 (get_z_vals_fine) and
 * :385-417 (sample_pdf).  Bit-exact against the reference's CPU path given identical inputs.
 *   z_coarse, weights [R,S];  u [R,N] (u_row_stride = N), one shared row [N] (u_row_stride = 0, the
 *   deterministic linspace of :394) or NULL (in-kernel Philox, not seed-compatible);
 *   z_fine [R,S+N] sorted;  nullable: samples [R,N], below / above int64 [R,N]. */
int srf_sample_pdf_merge(const float* z_coarse, const float* weights, const float* u, int64_t u_row_stride,
                         uint64_t seed, int64_t num_rays, int num_coarse, int num_fine, float* z_fine,
                         float* samples, int64_t* below, int64_t* above, void* stream);

/* Volume-rendering compositing, forward.  Replaces src/models/SimpleNeRF17.py:486-539 and
 * src/models/SimpleTensoRF09.py:767-819 (distance_scale = 25 there, 1 for NeRF).
 *   sigma, z [R,S];  rgb [R,S,3] nullable (then rgb_map must be NULL);  rays_* [R,3];
 *   ndc != 0: z are NDC depths, the last interval ends at 1, depth/depth_var are world depths
 *   (CommonUtils04.py:208-224) and depth_ndc/depth_var_ndc the NDC ones;  otherwise the last interval
 *   ends at 1e10 and the *_ndc outputs are ignored.
 *   nullable: alpha, visibility [R,S];  weights [R,S] is required. */
int srf_composite_fwd(const float* sigma, const float* rgb, const float* z, const float* rays_o,
                      const float* rays_d, const float* rays_d_ndc, int64_t num_rays, int num_samples,
                      int ndc, int white_bkgd, float distance_scale, float* alpha, float* visibility,
                      float* weights, float* rgb_map, float* acc, float* depth, float* depth_var,
                      float* depth_ndc, float* depth_var_ndc, void* stream);

/* Compositing backward (what autograd derives from the lines above).  `visibility`, acc, depth,
 * depth_ndc are the forward's outputs; upstream gradients are nullable; g_sigma [R,S];
 * g_rgb_samples [R,S,3] nullable. */
int srf_composite_bwd(const float* sigma, const float* rgb, const float* z, const float* visibility,
                      const float* rays_o, const float* rays_d, const float* rays_d_ndc, const float* acc,
                      const float* depth, const float* depth_ndc, const float* g_rgb, const float* g_acc,
                      const float* g_depth, const float* g_depth_ndc, const float* g_depth_var,
                      const float* g_depth_var_ndc, const float* g_weights, int64_t num_rays,
                      int num_samples, int ndc, int white_bkgd, float distance_scale, float* g_sigma,
                      float* g_rgb_samples, void* stream);

/* Fused NeRF MLP forward (tcgen05 tensor cores, bf16 operands, fp32 accumulation in TMEM).
 * Replaces src/models/SimpleNeRF17.py:213-215 (sample points), :581-613 (positional encoding), :419-484
 * (run_network / batchify) and :696-785 (MLP.forward + heads) for one MLP.
 *
 * The MLP variant is described by a layer program (HOST pointer, copied at launch).  The A operand of a
 * layer is a concatenation of 64-column shared-memory K-blocks ("regions"): 0 = point encoding E
 * (3*(2*points_degree+1) columns, zero padded), 1..4 = the four 64-column blocks of the 256 hidden
 * units, 5 = view-direction encoding V.  Weights: for every layer, for every 128-row half of its n output
 * units, for every K-block in program order, an image of 128 rows x 64 bf16 (row = output unit, column =
 * input column of that block) in which the 16-byte unit u of row r is stored at unit position
 * (u ^ (r & 7)) — the UMMA 128-byte swizzle — so one 16 KB bulk copy lands it in shared memory ready for
 * the tensor cores.  `side` is an fp32 table holding, per layer, the bias [n] and, for head layers, head
 * weights [rows][n] followed by head biases [rows]; bias_offset / head_offset are multiples of 4. */
typedef struct {
  int32_t num_kblocks;
  int32_t kblock_region[6];
  int32_t kblock_ksteps[6];   /* 16-wide MMA K steps to issue for the block (1..4) */
  int32_t n;                  /* 256 or 128 */
  int32_t relu;
  int32_t write_h;            /* keep the activation as the next layer's H blocks (n must be 256) */
  int32_t head;               /* 0 none; 1 sigma = relu(w.h + b [+ noise]); 2 sigma + sigmoid rgb (4 rows); 3 sigmoid rgb (3 rows) */
  int32_t bias_offset;        /* float offsets into `side` */
  int32_t head_offset;
  int32_t save_slot;          /* training: first image slot of this layer's output in the saved tile, or -1 */
  int64_t weight_offset;      /* byte offset into `weights` */
} srf_mlp_layer;

typedef struct {
  int32_t num_layers;         /* <= 12 */
  int32_t points_degree;      /* <= 10 */
  int32_t views_degree;       /* <= 4, or < 0 when the variant has no view branch */
  int32_t side_count;
  srf_mlp_layer layers[12];
  /* 0: bf16 operands (fp32 accumulate).  > 0: split-bf16 operands for the 1e-3 fp32 contract (inference only): activations and
   * weights are pairs hi = bf16(x), lo = bf16(x - hi), a K block contributes A_hi W_hi + A_lo W_hi + A_hi W_lo, and the lo image
   * of the weight image at byte offset o of `weights` sits at o + lo_offset (a multiple of 16). */
  int64_t lo_offset;
} srf_mlp_program;

/*   rays_o, rays_d [R,3]: origin / direction the sample points are built from (the NDC pair when ndc);
 *   z [R,S];  view_dirs [R,3] (NULL iff views_degree < 0);  noise [R*S] nullable: the reference's
 *   randn * raw_noise_std (SimpleNeRF17.py:739-741), added before the sigma ReLU;
 *   sigma [R*S], rgb [R*S,3]: post-activation outputs.
 *   Training (save_acts != NULL): every A-operand tile is also written to HBM as [tile][act_slots + M][128 x 64 bf16
 *   swizzled image] (encodings at e_slot / v_slot, layer outputs at layers[l].save_slot); the images feed
 *   srf_nerf_mlp_wgrad as GEMM operands.  The M = ceil(act_slots / 16) trailing "mask images" hold, for every saved layer
 *   output image s, one 32-bit word per (32-column group g, row r) at byte (act_slots + s / 16) * 16384 + (s % 16) * 1024 +
 *   g * 512 + r * 4 of the tile: bit i = pre-activation value of column 32 g + i is positive.  srf_nerf_mlp_dgrad reads
 *   these words as ReLU masks (128 bytes per warp instead of the activation rows).  `act_slots` counts the data images
 *   only, in all three entry points. */
int srf_nerf_mlp_fwd(const void* program, const void* weights, const float* side, const float* rays_o,
                     const float* rays_d, const float* z, const float* view_dirs, const float* noise,
                     int64_t num_rays, int num_samples, float* sigma, float* rgb, void* save_acts,
                     int act_slots, int e_slot, int v_slot, void* stream);
int srf_nerf_mlp_program_bytes(void);   /* sizeof(srf_mlp_program) as compiled, for binding self-checks */
/* bf16 inference launches of srf_nerf_mlp_fwd as clusters of two CTAs sharing every MMA (tcgen05 cta_group::2, M = 256, each CTA
 * holding half of every weight image): 1 on, 0 off, -1 back to the default (environment SRF_MLP_PAIR, else off).  Returns the
 * previous setting.  Results are bit-identical either way. */
int srf_mlp_set_pairing(int mode);

/* ---------------------------------------------------------------------------------------------------------
 * Simple-TensoRF vector-matrix tensor.  Small parameter blocks marked HOST are host pointers read at launch.
 * Planes / lines are channels-last derived caches: plane i is [res[a1]][res[a0]][C_i] with
 * (a0,a1) = (0,1),(0,2),(1,2) and line i is [res[v_i]][C_i] with v = (2,1,0) (SimpleTensoRF09.py:1131-1132,
 * :1159); C_i must be multiples of 4.  Compacted sample lists are int32 flat indices r*S+s in ascending
 * (row-major, stable) order with the element count kept on the DEVICE (no host synchronisation). */

/* 1 bit per voxel (x fastest, 32 voxels per word) from the fp32 {0,1} volume [Z,Y,X] (SimpleTensoRF09.py:1333). */
int srf_pack_alpha_bits(const float* volume, int64_t num_voxels, uint32_t* bits, void* stream);
int srf_compaction_blocks(int64_t total);   /* length of the block_counts / block_offsets scratch arrays */

/* pts = o + d z (SimpleTensoRF09.py:263), inside-box test (:705) and alphaMask test (:707-710, :1342-1349;
 * bit-exact w.r.t. F.grid_sample(...) > 0).  bbox HOST [2,3]; alpha_bits nullable; alpha_res HOST (X,Y,Z);
 * alpha_box_min / alpha_box_size HOST [3].  mask uint8 [R,S]; block_counts int32 [srf_compaction_blocks]. */
int srf_tensorf_mask(const float* rays_o, const float* rays_d, const float* z, int64_t num_rays, int num_samples,
                     const float* bbox, const uint32_t* alpha_bits, const int* alpha_res, const float* alpha_box_min,
                     const float* alpha_box_size, uint8_t* mask, int* block_counts, void* stream);
/* mask = values > threshold (SimpleTensoRF09.py:726). */
int srf_threshold_mask(const float* values, float threshold, int64_t total, uint8_t* mask, int* block_counts, void* stream);
/* Stable stream compaction of a mask produced by the two calls above (replaces pts[mask], :1221, :1248). */
int srf_compact(const uint8_t* mask, int64_t total, int* block_counts, int* block_offsets, int* indices, int* count,
                void* stream);

/* VM density (SimpleTensoRF09.py:763-765, :1214-1239): sigma[idx] = act(sum_i sum_c plane_i,c * line_i,c), act =
 * ReLU or softplus(x + density_offset); sigma [R*S] must be zero-filled by the caller; features [max_count]
 * (pre-activation, compacted order) is kept for the backward.  planes / lines: HOST arrays of 3 device pointers;
 * channels, resolution (X,Y,Z), box_min, box_size: HOST. */
int srf_vm_density_fwd(const float* rays_o, const float* rays_d, const float* z, int num_samples, const int* indices,
                       const int* count, int64_t max_count, const float* box_min, const float* box_size,
                       const float* const* planes, const float* const* lines, const int* channels, const int* resolution,
                       int softplus, float density_offset, float* sigma, float* features, void* stream);
/* its autograd: g_sigma [R*S] -> scatter-add into zero-initialised channels-last gradient buffers. */
int srf_vm_density_bwd(const float* rays_o, const float* rays_d, const float* z, int num_samples, const int* indices,
                       const int* count, int64_t max_count, const float* box_min, const float* box_size,
                       const float* const* planes, const float* const* lines, const int* channels, const int* resolution,
                       int softplus, float density_offset, const float* g_sigma, const float* features,
                       float* const* g_planes, float* const* g_lines, void* stream);

/* VM appearance features (SimpleTensoRF09.py:1241-1263): (plane x line) products over sum(C) <= 96 channels written as
 * bf16 rows [max_count, row_pitch] = [products | 3 view_dirs | zero pad] (row_pitch a multiple of 8 that holds
 * sum(C) + 3 elements, <= 128; every C a multiple of 4) — the A operand of the colour MLP.  basis_matrix_color (:1151, :1263) is linear and is folded into the MLP's first layer by the host
 * (W0' = [W0[:, :F] B | W0[:, F:]]), so this kernel is a pure gather.  The backward scatters g_rows[:, :sum(C)]
 * (fp32, row pitch g_row_pitch floats) into zero-initialised channels-last plane / line gradients. */
int srf_vm_color_features_fwd(const float* rays_o, const float* rays_d, const float* z, int num_samples, const int* indices,
                              const int* count, int64_t max_count, const float* box_min, const float* box_size,
                              const float* const* planes, const float* const* lines, const int* channels,
                              const int* resolution, const float* view_dirs, void* rows, int row_pitch, int z_is_ladder, void* stream);
int srf_vm_color_features_bwd(const float* rays_o, const float* rays_d, const float* z, int num_samples, const int* indices,
                              const int* count, int64_t max_count, const float* box_min, const float* box_size,
                              const float* const* planes, const float* const* lines, const int* channels,
                              const int* resolution, const float* g_rows, int g_row_pitch, float* const* g_planes,
                              float* const* g_lines, void* stream);

/* CANDECOMP/PARAFAC tensor (csrc/tensorf_cp.cu; `decomposition_type = "CandecompParafac"`, SimpleTensoRF09.py:537-539, :964-1124 —
 * selected by no shipped configuration).  Same contracts as the four VM calls above with lines only: `lines` is a HOST array of 3
 * device pointers to channels-last [L_i][components] lines, line i running along axis vector_axes[i] = 2 - i (:969), `components` a
 * multiple of 4 (<= 128 for the appearance calls).
 *   srf_cp_density_fwd / _bwd          :1043-1062: sigma[idx] = act(sum_c line_0,c line_1,c line_2,c), and its autograd
 *   srf_cp_color_features_fwd / _bwd   :1064-1078: bf16 rows [products (components) | 3 view_dirs | zero pad], and the scatter of
 *                                      g_rows[:, :components] into zero-initialised line gradients */
int srf_cp_density_fwd(const float* rays_o, const float* rays_d, const float* z, int num_samples, const int* indices, const int* count,
                       int64_t max_count, const float* box_min, const float* box_size, const float* const* lines, int components,
                       const int* resolution, int softplus, float density_offset, float* sigma, float* features, void* stream);
int srf_cp_density_bwd(const float* rays_o, const float* rays_d, const float* z, int num_samples, const int* indices, const int* count,
                       int64_t max_count, const float* box_min, const float* box_size, const float* const* lines, int components,
                       const int* resolution, int softplus, float density_offset, const float* g_sigma, const float* features,
                       float* const* g_lines, void* stream);
int srf_cp_color_features_fwd(const float* rays_o, const float* rays_d, const float* z, int num_samples, const int* indices,
                              const int* count, int64_t max_count, const float* box_min, const float* box_size,
                              const float* const* lines, int components, const int* resolution, const float* view_dirs, void* rows,
                              int row_pitch, void* stream);
int srf_cp_color_features_bwd(const float* rays_o, const float* rays_d, const float* z, int num_samples, const int* indices,
                              const int* count, int64_t max_count, const float* box_min, const float* box_size,
                              const float* const* lines, int components, const int* resolution, const float* g_rows,
                              int g_row_pitch, float* const* g_lines, void* stream);

/* dst[indices[j], :] = src[j, :] (the scatter-back of :1238 / :1271) and its transpose. */
int srf_scatter_rows(const int* indices, const int* count, int64_t max_count, const float* src, int width, float* dst,
                     void* stream);
int srf_gather_rows(const int* indices, const int* count, int64_t max_count, const float* src, int width, float* dst,
                    void* stream);

/* ---- "next" row f3: grid surgery (csrc/tensorf_surgery.cu).
 * Alpha-mask rebuild, LowRankTensor.update_alpha_mask (SimpleTensoRF09.py:849-876) without any fp32 volume:
 *   1. srf_alpha_grid_occupancy: for every voxel of the (X,Y,Z) = resolution grid, world point = (coord_x[x], coord_y[y],
 *      coord_z[z]) (DEVICE arrays holding bb0 (1 - s) + bb1 s, s = linspace(0, 1, n), :850-855), previous alpha-mask test
 *      (prev_bits nullable; :879-883), normalise (:763-765), VM density (:1214-1239), alpha = 1 - exp(-sigma step_size) (:895),
 *      clamp (:862) and `>= threshold` -> raw_words [Z][Y][ceil(X/32)] (bit x & 31 of word x >> 5; srf_alpha_grid_words words).
 *   2. srf_alpha_grid_dilate: the 3x3x3 max-pool of :864-865 as a bit dilation (max over a window >= t <=> any member >= t)
 *      -> volume uint8 {0,1} [Z,Y,X] (the new AlphaGridMask.alpha_volume, :869) and projection uint32[ceil(X/32) + Y + Z]
 *      (CALLER-zeroed): the occupied set projected on each axis — x as row-layout bit words, then one flag per y, per z —
 *      from which the host takes the new bounding box (:871-875: amin / amax of the occupied voxels' coordinates).
 * box_min / box_size / prev_box_* / resolution / prev_res / channels are HOST arrays; planes / lines as for srf_vm_density_fwd.
 * planes == NULL evaluates a CANDECOMP/PARAFAC tensor instead (:1043-1062): lines only, channels[0] components in every line. */
int srf_alpha_grid_words(const int* resolution);
int srf_alpha_grid_occupancy(const float* const* planes, const float* const* lines, const int* channels, const int* resolution,
                             const float* box_min, const float* box_size, const float* coord_x, const float* coord_y,
                             const float* coord_z, const uint32_t* prev_bits, const int* prev_res, const float* prev_box_min,
                             const float* prev_box_size, int softplus, float density_offset, float step_size, float threshold,
                             uint32_t* raw_words, void* stream);
int srf_alpha_grid_dilate(const uint32_t* raw_words, const int* resolution, uint8_t* volume, uint32_t* projection, void* stream);
/* srf_pack_alpha_bits for a bool / uint8 volume (the in-memory form of the drop-in's AlphaGridMask.alpha_volume). */
int srf_pack_alpha_bits_u8(const uint8_t* volume, int64_t num_voxels, uint32_t* bits, void* stream);
/* dst [C, out_height, out_width] = bilinear resampling (F.interpolate(mode='bilinear', align_corners=True), ATen's
 * upsample_bilinear2d arithmetic; SimpleTensoRF09.py:1284-1295) of the window [y0, y0+h) x [x0, x0+w) of src [C, height, width];
 * with (out_height, out_width) == (h, w) it is the window copy of shrink_tensor (:1303-1319). */
int srf_resample_plane(const float* src, int channels, int height, int width, int y0, int x0, int h, int w, float* dst,
                       int out_height, int out_width, void* stream);

/* Colour MLP on the tensor cores: same kernel and program format as srf_nerf_mlp_fwd, but regions 0 and 5 are filled from
 * precomputed bf16 rows [max_rows, row_pitch] (columns 0..63 / 64..row_pitch-1, zero beyond) instead of encodings (views_degree = -2 in the
 * program; -1 when only region 0 is used); the row count is read from the device. */
int srf_mlp_rows_fwd(const void* program, const void* weights, const float* side, const void* rows, int row_pitch, const int* count,
                     int64_t max_rows, float* rgb, void* save_acts, int act_slots, int e_slot, int v_slot, void* stream);

/* Weight gradients of the fused MLP (what autograd derives for the nn.Linear layers of
 * src/models/SimpleNeRF17.py:644-666): dW = dZ^T X and db = colsum(dZ) on tcgen05, K = samples.  `acts` are the
 * activation tile images saved by srf_nerf_mlp_fwd, `dz` the pre-activation gradient images written by
 * srf_nerf_mlp_dgrad (same 128 x 64 bf16 swizzled format, [tile][slots][16 KB]).  Each work item (HOST array)
 * multiplies dz images [dz_slot, dz_slot + dz_images) (64 output channels each; 2 or 4 images) with act images
 * [x_slot, x_slot + x_images) and ADDS rows [0, out_rows) x image columns [in_col0, in_col0 + in_cols) into
 * grads[dw_offset + row * w_stride + w_col0 + (col - in_col0)] (+ column sums into grads[db_offset + row] if bias).
 * `count` (device, nullable; also in srf_nerf_mlp_dgrad): the number of valid rows when the buffers are sized for a worst case
 * (TensoRF surface samples: the count stays on the device, no read-back synchronisation in the training step). */
typedef struct {
  int32_t dz_slot, dz_images, x_slot, x_images;
  int32_t out_rows, in_col0, in_cols, w_col0, w_stride, bias;
  int64_t dw_offset, db_offset;
} srf_wgrad_item;
int srf_nerf_mlp_wgrad(const void* items, int num_items, const void* acts, int act_slots, const void* dz, int dz_slots,
                       int64_t num_tiles, const int* count, float* grads, void* stream);
int srf_wgrad_item_bytes(void);

/* Data-gradient chain of the fused MLP (what autograd derives for src/models/SimpleNeRF17.py:726-785): from
 * g_sigma [M], g_rgb [M,3] (nullable) through the heads and every hidden layer down to layer 1, on tcgen05 with the
 * transposed weights streamed as swizzled images ([layer][128-row half of the 256 inputs][K block of outputs]).
 * Reads the mask words saved next to the activation tiles (see srf_nerf_mlp_fwd) and the sigma / rgb outputs of
 * srf_nerf_mlp_fwd; writes every layer's pre-activation
 * gradient as tile images into dz [tile][dz_slots][16 KB] for srf_nerf_mlp_wgrad.
 * Every backward layer has its own output width n_out (128 or 256).  The TensoRF colour MLP
 * (src/models/SimpleTensoRF09.py:1389-1393, srf_mlp_rows_fwd) uses the same chain with 128-wide layers; its last layer
 * sets dz_slot = -1 and rows_cols = sum(C): the gradient of the product rows is written as fp32 g_rows [M, g_row_pitch]
 * for srf_vm_color_features_bwd. */
typedef struct {
  int32_t num_kblocks, mask_slot, rank1_offset, dz_slot, n_out, rows_cols;
  int64_t weight_offset;
} srf_dgrad_layer;
typedef struct {
  int32_t num_layers, num_fwd_layers, top_width, top_mask_slot, top_slot, head_slot, head_kind, head_w_offset, side_count, pad_;
  srf_dgrad_layer layers[12];
} srf_dgrad_program;
int srf_nerf_mlp_dgrad(const void* program, const void* weights_t, const float* side, const void* acts, int act_slots,
                       const float* sigma, const float* rgb, const float* g_sigma, const float* g_rgb, int64_t num_rows,
                       const int* count, void* dz, int dz_slots, float* g_rows, int g_row_pitch, void* stream);
int srf_dgrad_program_bytes(void);

/* Gradient of one fused-MLP evaluation with respect to its INPUTS, for learnable cameras (the pose correction r, t of
 * src/models/SimpleNeRF17.py:817-842 receives its gradient through the rays: pts = o + z d at :210-214 and view_dirs at :190; autograd
 * derives it through the positional encoding :669-693 and the first / skip / view layers :726-765).  Reads the dZ images
 * srf_nerf_mlp_dgrad wrote and the fp32 parameter vector: for every layer that consumes an encoding image (`sources`, HOST array),
 * g_enc[row, j] += sum_n dZ[row, n] W[n, cols[j]]; then the encoding backward
 * g_x[c] = g_enc[c] + sum_k 2^k (cos(2^k x_c) g_enc[3 + 6k + c] - sin(2^k x_c) g_enc[6 + 6k + c]).
 * Writes g_points [num_rows, 3] and g_views [num_rows, 3] (nullable when no source has target 1); rays_o / rays_d [num_rows / num_samples, 3],
 * z [num_rows] and view_dirs are the arrays srf_nerf_mlp_fwd was given.  Rows mode (num_samples == 0: the TensoRF colour MLP of
 * src/models/SimpleTensoRF09.py:1411-1421 fed by srf_mlp_rows_fwd; rays_o / rays_d / z may be NULL): no encoding, g_points[row, c] = g_enc[c]
 * for the three weight columns a source maps to image columns 0..2 (the view directions behind the products); `count` (device,
 * nullable) = number of valid rows, rows beyond it are written as zero.  Off the hot path of every shipped configuration (cameras frozen). */
typedef struct {
  int32_t dz_slot, dz_images, in_total, target;   /* target 0: points encoding image, 1: view encoding image (32 columns) */
  int64_t w_offset;                               /* element offset of the weight matrix [64 * dz_images, in_total] in params */
  int32_t cols[64];                               /* encoding image column -> weight column, -1: not consumed */
} srf_input_grad_source;
int srf_nerf_mlp_input_grad(const void* sources, int num_sources, const float* params, const void* dz, int dz_slots,
                            const float* rays_o, const float* rays_d, const float* z, const float* view_dirs,
                            int64_t num_rows, const int* count, int num_samples, int points_degree, int views_degree,
                            float* g_points, float* g_views, void* stream);
int srf_input_grad_source_bytes(void);

/* ---- fused test-time ray march of one VM tensor (NDC), rows IX-XII + the weights half of VIII of SURVEY.md §8a in one kernel:
 * sample depths from the shared [num_samples] ladder (src/models/SimpleTensoRF09.py:363-376 at test time) -> points (:263) ->
 * box test (:705) -> alphaMask test (:707-710, :1342-1349) -> VM density (:1214-1239) -> alpha / transmittance / weights
 * (:767-790) -> acc, depth, depth_var, depth_ndc, depth_var_ndc (:792-811) -> surface test weights > threshold (:726).
 * No [num_rays, num_samples] intermediate is read; the surface samples of ray r are appended, in sample order, to row r of
 * entry_sample / entry_weight ([num_rays, num_samples], first ray_count[r] entries valid).  z_vals, dense sigma / weights are
 * NOT produced: callers that need them (`retraw`, training) use the unfused entry points above.  num_rays * num_samples < 2^31.
 * bbox = [min xyz | max xyz], box_size = the tensor's bounding_box_size buffer (host pointers, 6 + 3 floats); alpha_* nullable.
 * alpha_corner_or (nullable, DEVICE, srf_alpha_corner_or_words(alpha_res) words from srf_alpha_corner_or_bits): bit (x,y,z) of
 * the (X+1)(Y+1)(Z+1) volume = OR of the alpha bits at (x-1..x, y-1..y, z-1..z).  A point whose approximate voxel coordinates
 * all lie >= 1/256 voxel away from a voxel boundary is decided by ONE bit of it (both trilinear weights of every axis are then
 * strictly positive, so grid_sample(...) > 0 <=> some in-range corner is set); points next to a boundary take the exact test —
 * the validity set is unchanged.  num_samples <= 65535.
 *
 * srf_tensorf_march_compact: the per-ray lists -> one flat list in row-major (ray, sample) order — the order of the
 * reference's boolean-mask indexing (:1248) — indices[j] = ray * num_samples + sample, weights[j], ray_offset[ray], *count.
 * scratch: int32[num_rays + srf_tensorf_march_blocks(num_rays)].
 *
 * srf_ray_accumulate: rgb_map[r] = sum_k weights[ray_offset[r] + k] * rgb_rows[ray_offset[r] + k] (+ 1 - acc[r] on a white
 * background): the colour half of volume_render (:813-817) over the surface samples only (rgb is 0 elsewhere, :1271). */
int srf_tensorf_march(const float* rays_o_ndc, const float* rays_d_ndc, const float* rays_o, const float* rays_d,
                      const float* ladder, int64_t num_rays, int num_samples, const float* bbox, const float* box_size,
                      const uint32_t* alpha_bits, const uint32_t* alpha_corner_or, const int* alpha_res, const float* alpha_box_min,
                      const float* alpha_box_size, const float* const* planes, const float* const* lines, const int* channels,
                      const int* resolution, int softplus, float density_offset, float distance_scale, float weight_threshold,
                      float* acc, float* depth, float* depth_var, float* depth_ndc, float* depth_var_ndc,
                      int* ray_count, int* entry_sample, float* entry_weight, void* stream);
int srf_alpha_corner_or_words(const int* alpha_res);
int srf_alpha_corner_or_bits(const uint32_t* alpha_bits, const int* alpha_res, uint32_t* corner_or, void* stream);
int srf_tensorf_march_blocks(int64_t num_rays);
int srf_tensorf_march_compact(const int* ray_count, int64_t num_rays, int num_samples, const int* entry_sample,
                              const float* entry_weight, int* scratch, int* ray_offset, int* indices, float* weights,
                              int* count, void* stream);
int srf_ray_accumulate(const float* rgb_rows, const float* weights, const int* ray_offset, const int* ray_count,
                       const float* acc, int64_t num_rays, int white_bkgd, float* rgb_map, void* stream);

/* ---- "next" row f1: masks of the patch-reprojection depth losses
 * (src/loss_functions/AugmentationsDepthLoss11.py:105-182, src/loss_functions/CoarseFineConsistencyLoss34.py:89-164,
 * src/utils/CommonUtils04.py:227-253).  For each of num_rays image rays: the two candidate depths are reprojected into
 * the closest other training view (closest_view [V], poses [V,4,4] camera-to-world, intrinsics_first = device pointer to the 3x3
 * intrinsics of the first ray, which the reference hard-codes), patch_x x patch_y rgb patches of images [V,H,W,3] around
 * pixel_id [N,3] = (view, x, y) and the two reprojections are compared (RMSE, zeros outside the frame) and mask1 / mask2
 * [N] (0/1) mark the rays where model 1 / model 2 is the more accurate one.  both_invalid_rule = 1 adds
 * AugmentationsDepthLoss11.py:178-182 (prefer the larger depth when both reprojections leave the frame).  rmse1 / rmse2 are
 * optional [N] outputs. */
int srf_patch_reprojection_masks(const float* rays_o, const float* rays_d, const float* depth1, const float* depth2,
                                 const int32_t* pixel_id, int64_t num_rays, const int32_t* closest_view, const float* poses,
                                 const float* intrinsics_first, const float* images, int num_views, int height, int width,
                                 int patch_x, int patch_y, float rmse_threshold, int both_invalid_rule, uint8_t* mask1,
                                 uint8_t* mask2, float* rmse1, float* rmse2, void* stream);

/* ---- "next" row f4: total-variation regulariser of VM planes, loss and gradient in one launch.  Replaces
 * TotalVariationLoss04.compute_tv_loss (src/loss_functions/TotalVariationLoss04.py:97-116) and its autograd: for each of the
 * num_planes (<= 12) planes [C,H,W] (dims = {C,H,W} per plane; device pointers in host arrays)
 *   *loss += 2 * (sum_h (x[h+1]-x[h])^2 / max(C (H-1) W, 1) + sum_w (x[w+1]-x[w])^2 / max(C H (W-1), 1)) * iter_weight
 * (double accumulator, device, caller-zeroed) and grads[i] = d(that plane's term) / d x. */
int srf_tv_loss(const float* const* planes, float* const* grads, const int* dims, int num_planes, float iter_weight,
                double* loss, void* stream);

/* ---- "next" row f2: device-side batch assembly.  Replaces load_nerf_cached_batch /
 * load_sparse_depth_cached_batch (src/data_preprocessors/DataPreprocessor10.py:530-549, 568-595): row b takes the cached
 * per-pixel tables at flat index indices[b] (all tables [num_pixels, C], device); image rays (is_sparse_depth[b] == 0, or
 * is_sparse_depth NULL) get pixel_id + target_rgb, sparse-depth rays get pixel_id + depth / reprojection error / 3-D point;
 * fields a ray kind does not carry are -1, as the reference initialises them.  sd_* outputs (and their tables) are nullable.
 * Indices must lie in [0, num_pixels) (they are the reference's own shuffled index arrays); a row whose index does not is
 * filled with -1 and *error_flag (nullable; device-accessible, e.g. mapped pinned host memory) is set to 1 — the reference's
 * fancy indexing raises IndexError there. */
int srf_assemble_batch(const int64_t* indices, const uint8_t* is_sparse_depth, int64_t batch, int64_t num_pixels,
                       const int* pixel_table, const float* rgb_table, const float* depth_table, const float* error_table,
                       const float* points_table, int* pixel_id, float* target_rgb, float* sd_depth, float* sd_error,
                       float* sd_points, int* error_flag, void* stream);

/* ---- "next" row f4: output tail of a rendered frame.  Replaces retrieve_inference_outputs
 * (src/data_preprocessors/DataPreprocessor10.py:775-803) with its post_process_output / post_process_image /
 * post_process_depth helpers (:967-995): instead of copying every tensor of the output dict to the host and converting there,
 * one kernel writes the record the caller keeps —
 *   [ image uint8 [num_rays,3] | depth | depth_var | depth_ndc | depth_var_ndc  (fp32 [num_rays] each) ],
 * every section padded to 16 bytes (srf_frame_record_bytes(num_rays, 4) in total; the two NDC maps are nullable and then left
 * unwritten) — with numpy's arithmetic: clip(rgb,0,1)*255 rounded half-to-even -> uint8, negative depths -> 0.  All pointers
 * device, 16-byte aligned. */
int64_t srf_frame_record_bytes(int64_t num_rays, int num_maps);
int srf_frame_outputs(const float* rgb, const float* depth, const float* depth_var, const float* depth_ndc,
                      const float* depth_var_ndc, int64_t num_rays, uint8_t* record, void* stream);

/* ---- "next" row f4: optimiser tail.  One fused Adam step over flat fp32 arrays (device pointers, 16-byte
 * aligned), replacing torch.optim.Adam.step as created by src/optimizers/OptimizerFactory02.py:9-22 and called at
 * src/Trainer10.py:109-110; arithmetic of torch/optim/adam.py::_single_tensor_adam (amsgrad / maximize off), `step` counts
 * from 1, weight_decay is the L2 form (0 in every shipped config). */
int srf_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, float lr, float beta1,
                  float beta2, float eps, float weight_decay, int64_t step, void* stream);

/* Capturable form of the same step (what torch.optim.Adam(capturable=True) provides): the step count (int64, device) and the
 * learning rate (fp32, device) are read by the kernel, so a CUDA graph holding `srf_adam_advance` (step += 1) followed by
 * `srf_adam_step_capturable` replays correctly; bias corrections are evaluated in double on the device. */
int srf_adam_advance(int64_t* step, void* stream);
int srf_adam_step_capturable(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, const float* lr,
                             float beta1, float beta2, float eps, float weight_decay, const int64_t* step, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SIMPLE_RF_B200_H */
