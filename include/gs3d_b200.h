/*
 * gs3d_b200.h -- C ABI of the B200-native Gaussian-splatting rasteriser (libgs3d_b200.so).
 *
 * This is the drop-in boundary for the hot path of heheyas/gaussian_splatting_3d: every entry
 * point replaces one binding of the reference's `_gs` pybind module (gs/src/bindings.cpp:5-67,
 * prototypes gs/src/render.h:3-131) or one torch-level function that sits between two bindings
 * on the path (gs/renderer.py, gs/culling.py, utils/camera.py).  Plain pointers and sizes only:
 * no torch types.  All pointers are DEVICE pointers unless the name ends in `_host`.
 *
 * Conventions
 *   - every function returns 0 on success, else a GS3D_E* code; gs3d_last_error() gives the
 *     message (thread-local).  Nothing calls exit() or traps (the reference printf+exit()s,
 *     common.h:56-72).
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it (the reference runs
 *     its cull/bin kernels on the legacy stream and cudaMalloc/cudaFree/D2H-syncs inside,
 *     aabb_culling.h:204-259; this library never allocates: scratch comes from the caller).
 *   - outputs are written in place into caller-allocated buffers, like the reference.
 *   - layouts are the reference's: mean [N,3], qvec [N,4] (w,x,y,z), svec [N,3], mean2d [N,2],
 *     cov2d [N,4] row-major 2x2 with S01 and S10 kept apart, sh_coeffs channel-major
 *     [N][3][C*C], image HWC float32, start/end int32 with -1 = empty tile, keys int64
 *     (tile id << 32 | float32 depth bits).
 *   - only tile_size == 16 is supported by the compositing kernels (all reference configs,
 *     conf/ *.yaml `tile_size: 16`); other sizes return GS3D_EUNSUPPORTED.
 */
#ifndef GS3D_B200_H
#define GS3D_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GS3D_OK 0
#define GS3D_EINVAL 1       /* bad argument (null pointer, size mismatch, bad C) */
#define GS3D_ECUDA 2        /* CUDA runtime / launch failure */
#define GS3D_EUNSUPPORTED 3 /* valid in the reference, not implemented here (e.g. tile_size != 16) */
#define GS3D_ECOUNT 4       /* emitted duplicate count != n_dub (reference: assert, aabb_culling.h:228) */

/* Mirrors utils/camera.py:219-230 CameraInfo (intrinsics holder). */
typedef struct gs3d_camera {
  double fx, fy, cx, cy; /* Python floats in the reference; rounded to FP32 where torch would */
  int32_t w, h;
  double near_plane, far_plane;
} gs3d_camera;

/* Per-Gaussian staging record consumed by the compositing kernels: 12 floats = 48 B, 16-B aligned.
 *   [0] mean2d.x  [1] mean2d.y  [2] min(alpha,0.99)  [3] log2((1/255)/alpha_) (skip threshold)
 *   [4..6] -0.5*log2(e) * (c3, -(c1+c2), c0) / det   (fast conic form)   [7] depth
 *   [8..11] c0 c1 c2 c3 (raw covariance; exact-decision path and backward)            */
#define GS3D_RECORD_FLOATS 12

int gs3d_version(void);
const char *gs3d_last_error(void);
/* number of CUDA kernels this library has launched in this process (bench.py's gpu_launches) */
uint64_t gs3d_launch_count(void);

/* ---- a1  CameraInfo.get_frustum, utils/camera.py:249-283.  c2w [3,4] device; normals/pts [6,3]. */
int gs3d_get_frustum(const float *c2w, const gs3d_camera *cam_host, float *normals, float *pts,
                     void *stream);

/* ---- a2  culling_gaussian_bsphere, bindings.cpp:6 / render.cu:15-43 / culling.h:11-34.
 * mask is torch.bool storage (1 byte per Gaussian). qvec is accepted and unused, as in the reference. */
int gs3d_culling_gaussian_bsphere(uint32_t N, const float *mean, const float *qvec,
                                  const float *svec, const float *normal, const float *pts,
                                  uint8_t *mask, float thresh, void *stream);

/* ---- a4  project_gaussians, gs/renderer.py:391-419 (+ :366-387, utils/transforms.py:31-45).
 * JW [N,3,3] may be NULL.  depth [N].  Forward only; see gs3d_project_gaussians_backward. */
int gs3d_project_gaussians(uint32_t N, const float *mean, const float *qvec, const float *svec,
                           const float *c2w, float *mean2d, float *cov2d, float *JW, float *depth,
                           void *stream);

/* autograd of a4 (quirk Q6: J constant, depth detached unless detach_depth == 0).
 * grad_depth may be NULL.  Outputs are OVERWRITTEN: grad_mean [N,3], grad_qvec [N,4],
 * grad_svec [N,3]. */
int gs3d_project_gaussians_backward(uint32_t N, const float *mean, const float *qvec,
                                    const float *svec, const float *c2w, const float *grad_mean2d,
                                    const float *grad_cov2d, const float *grad_depth,
                                    int detach_depth, float *grad_mean, float *grad_qvec,
                                    float *grad_svec, void *stream);

/* ---- a5  tile_culling_aabb_count, gs/culling.py:8-37 + utils/camera.py:290-303 (quirk Q8).
 * aabb_topleft / aabb_bottomright int32 [N,2] (tile units, inclusive).  The duplicate count is
 * written to *n_dub_host after a stream synchronise (the reference's `.item()`, culling.py:33-35).
 * scratch: >= gs3d_count_scratch_bytes(N) bytes of device memory. */
size_t gs3d_count_scratch_bytes(uint32_t N);
int gs3d_tile_culling_aabb_count(uint32_t N, const float *mean2d, const float *cov2d,
                                 uint32_t tile_size, const gs3d_camera *cam_host, float D,
                                 int32_t *aabb_topleft, int32_t *aabb_bottomright,
                                 int64_t *n_dub_host, void *scratch, size_t scratch_bytes,
                                 void *stream);

/* ---- a2+a4+a5 fused (the product path; replaces sh_renderer.py:189-236 minus the mask
 * compaction): frustum planes, sphere cull, activations (svec = exp, alpha = sigmoid when the
 * *_act flags are 1; 0 = identity), projection, tile rect, duplicate count and staging record in
 * ONE pass over the parameters.  Culled Gaussians get rect tl=(0,0) br=(-1,-1) (zero duplicates)
 * and mask 0; nothing is compacted, ids stay original indices.
 * Outputs: mask [N] u8, mean2d [N,2], cov2d [N,4], depth [N], aabb tl/br int32 [N,2],
 * records [N,12] (may be NULL), svec_out [N,3] / alpha_out [N] activated values (may be NULL),
 * mean2d and cov2d may BOTH be NULL when records is not: the record already holds them (floats 0-1 = mean2d,
 * floats 8-11 = cov2d; for a culled Gaussian the record's covariance is the identity, not zero),
 * cnt int32 [N] (may be NULL): cnt[i] += 1 for kept Gaussians (sh_renderer.py:215-216).
 * *n_dub_host is valid after return (one 8-byte D2H + stream sync).  n_dub_host == NULL: no read-back and no
 * synchronisation -- the count stays on the device in the first 8 bytes of `scratch` (uint64); use it with
 * gs3d_tile_culling_aabb_start_end_capacity. */
int gs3d_project_cull_fused(uint32_t N, const float *mean, const float *qvec,
                            const float *svec_param, const float *alpha_param, int svec_act,
                            int alpha_act, const float *c2w, const gs3d_camera *cam_host,
                            float frustum_radius, int skip_frustum_culling, float tile_D,
                            uint32_t tile_size, uint8_t *mask, float *mean2d, float *cov2d,
                            float *depth, int32_t *aabb_topleft, int32_t *aabb_bottomright,
                            float *records, float *svec_out, float *alpha_out, int32_t *cnt,
                            int64_t *n_dub_host, void *scratch, size_t scratch_bytes, void *stream);

/* ---- a6  tile_culling_aabb_start_end, bindings.cpp:27 / render.cu:380-397 /
 * aabb_culling.h:15-41,70-103,192-260.  Hand-written LSD radix sort of the 64-bit key
 * (tile << 32 | depth bits): the four depth-byte passes run over the N Gaussians BEFORE
 * duplication, the tile-digit passes over the n_dub duplicates; the result equals a stable
 * ascending int64 sort of the reference's keys with ties (same tile, same depth bits) in
 * ascending Gaussian id -- one of the orders the reference's atomic emission can produce.
 * gaussian_ids int32 [n_dub], start/end int32 [n_tiles] (-1 = empty), sorted_keys int64 [n_dub]
 * optional (NULL to skip).  Limits: N, n_dub < 2^30; n_tiles_h, n_tiles_w <= 65535 and n_tiles < 2^24 (the rects
 * travel through the depth sort packed in 8 bytes; rects are expected clamped to the tile grid, as
 * gs/culling.py:29-31 and gs3d_tile_culling_aabb_count produce them).  Returns GS3D_ECOUNT (after a sync) only when check_count != 0 and the
 * rects do not add up to n_dub. */
size_t gs3d_binning_scratch_bytes(uint32_t N, uint32_t n_dub);
int gs3d_tile_culling_aabb_start_end(uint32_t N, uint32_t n_dub, uint32_t n_tiles_h,
                                     uint32_t n_tiles_w, const int32_t *aabb_topleft,
                                     const int32_t *aabb_bottomright, const float *depth,
                                     int32_t *gaussian_ids, int32_t *start, int32_t *end,
                                     int64_t *sorted_keys, int check_count, void *scratch,
                                     size_t scratch_bytes, void *stream);

/* Same binning without ANY host round trip: the reference learns the duplicate count through `.item()`
 * (gs/culling.py:33-35) and a D2H memcpy (aabb_culling.h:222-228) before it can size gaussian_ids.  Here the
 * caller passes a buffer of `capacity` ids instead; the count only ever exists on the device: n_dub_dev
 * (int64, device) receives the true count, overflow_dev (int32, device) is set to 1 when it exceeded the
 * capacity -- the lists are then truncated (the farthest duplicates are dropped) and the caller must re-run with
 * a larger buffer.  Every launch is sized by the capacity, so the call (and with gs3d_project_cull_fused's
 * n_dub_host == NULL the whole forward+backward step) can be captured in a CUDA graph.
 * Scratch: gs3d_binning_scratch_bytes(N, capacity). */
int gs3d_tile_culling_aabb_start_end_capacity(uint32_t N, uint32_t capacity, uint32_t n_tiles_h,
                                              uint32_t n_tiles_w, const int32_t *aabb_topleft,
                                              const int32_t *aabb_bottomright, const float *depth,
                                              int32_t *gaussian_ids, int32_t *start, int32_t *end,
                                              int64_t *n_dub_dev, int32_t *overflow_dev, void *scratch,
                                              size_t scratch_bytes, void *stream);

/* ---- staging records from raw arrays (used by the reference-signature wrappers below). */
int gs3d_pack_records(uint32_t N, const float *mean2d, const float *cov2d, const float *alpha,
                      const float *depth /* may be NULL */, float *records, void *stream);

/* ---- a7 / a11  tile_based_vol_rendering_sh{,_with_bg}, bindings.cpp:42,57 /
 * render.cu:483-544 / vol_render_sh.h:97-266 / vol_render_bg.h:12-100.
 * records [M,12]; sh_coeffs rows addressed as sh_coeffs + g*sh_stride_g + c*sh_stride_c + k
 * (reference layout: sh_stride_g = 3*C*C, sh_stride_c = C*C).  topleft [2], c2w [3,4], bg_rgb [3]
 * device pointers (bg_rgb NULL = no background).  out [H*W*3] is fully written for non-empty
 * tiles; pixels of empty tiles are left untouched (caller pre-zeroes, renderer.py:693) or set to
 * bg.  final_T [H*W] and n_contrib int32 [H*W] are optional extra outputs (NULL to skip).
 * exact_decisions != 0 re-evaluates the Gaussian in the reference's exact FP32 operation order
 * whenever alpha*G is within 2e-3 (relative) of the 1/255 skip threshold. */
int gs3d_composite_sh_forward(uint32_t M, const float *records, const float *sh_coeffs,
                              uint32_t sh_stride_g, uint32_t sh_stride_c, const int32_t *start,
                              const int32_t *end, const int32_t *gaussian_ids, float *out,
                              const float *topleft, const float *c2w, uint32_t tile_size,
                              uint32_t n_tiles_h, uint32_t n_tiles_w, float pixel_size_x,
                              float pixel_size_y, uint32_t H, uint32_t W, uint32_t C, float thresh,
                              const float *bg_rgb, float *final_T, int32_t *n_contrib,
                              int exact_decisions, void *stream);

/* ---- a8 / a11  tile_based_vol_rendering_backward_sh{,_with_bg}, bindings.cpp:44,60 /
 * render.cu:546-624 / vol_render_sh.h:268-480 / kernels.h:394-418.
 * Gradients are ACCUMULATED (caller pre-zeroes, renderer.py:761-764): grad_mean2d [M,2],
 * grad_cov2d [M,4], grad_alpha [M], grad_sh rows at grad_sh + g*gsh_stride_g + c*gsh_stride_c + k.
 * out is the saved forward image.  Per-warp reduction in shared memory, one vector
 * red.global.add per (tile, Gaussian) row instead of the reference's 256 shared atomics. */
int gs3d_composite_sh_backward(uint32_t M, const float *records, const float *sh_coeffs,
                               uint32_t sh_stride_g, uint32_t sh_stride_c, const int32_t *start,
                               const int32_t *end, const int32_t *gaussian_ids, const float *out,
                               const float *grad_out, float *grad_mean2d, float *grad_cov2d,
                               float *grad_sh, uint32_t gsh_stride_g, uint32_t gsh_stride_c,
                               float *grad_alpha, const float *topleft, const float *c2w,
                               uint32_t tile_size, uint32_t n_tiles_h, uint32_t n_tiles_w,
                               float pixel_size_x, float pixel_size_y, uint32_t H, uint32_t W,
                               uint32_t C, float thresh, int exact_decisions, void *stream);

/* Same kernel with the FUSED GRADIENT EXCHANGE of view-sharded training (new capability; the
 * reference has no multi-GPU code): the SH-coefficient gradient rows -- 576 of the 708 MB of leaf
 * gradients at C = 4, and sparse in practice -- are reduced by the compositing-backward kernel itself
 * into EVERY rank's gradient buffer while it computes: with one `multimem.red.add.v4.f32` per row
 * chunk on an NVSwitch multicast address (multicast_grad_sh != NULL) or with one red.global.add per
 * peer over NVLink (peer_grad_sh_host[0..n_peers): device addresses of each rank's grad_sh, this
 * rank included).  All buffers must be zeroed and a cross-rank barrier passed before the launch;
 * after a second barrier each buffer holds the sum over ranks.  n_peers == 0: identical to
 * gs3d_composite_sh_backward. */
int gs3d_composite_sh_backward_peers(uint32_t M, const float *records, const float *sh_coeffs,
                                     uint32_t sh_stride_g, uint32_t sh_stride_c, const int32_t *start,
                                     const int32_t *end, const int32_t *gaussian_ids, const float *out,
                                     const float *grad_out, float *grad_mean2d, float *grad_cov2d,
                                     float *grad_sh, uint32_t gsh_stride_g, uint32_t gsh_stride_c,
                                     float *grad_alpha, const float *topleft, const float *c2w,
                                     uint32_t tile_size, uint32_t n_tiles_h, uint32_t n_tiles_w,
                                     float pixel_size_x, float pixel_size_y, uint32_t H, uint32_t W,
                                     uint32_t C, float thresh, int exact_decisions,
                                     const uint64_t *peer_grad_sh_host, int n_peers,
                                     void *multicast_grad_sh, uint8_t *touched, void *stream);

/* ---- sparse gradient exchange (view-sharded training, new capability).  `touched` above (may be
 * NULL) receives a 1 for every Gaussian whose gradient rows the backward wrote; after the ranks have
 * OR-ed their marks, only the union rows are all-reduced.  gs3d_rows_gather packs rows row_idx[0..U)
 * of n_blocks row-major blocks (block s = [N, block_widths_host[s]] floats at device address
 * block_ptrs_host[s]; both arrays live on the HOST) side by side into packed [U, W] (W >= sum of the
 * widths, padding columns zeroed); gs3d_rows_scatter writes a packed matrix back. */
int gs3d_rows_gather(int n_blocks, const uint64_t *block_ptrs_host, const uint32_t *block_widths_host,
                     const int32_t *row_idx, uint32_t U, float *packed, uint32_t W, void *stream);
int gs3d_rows_scatter(int n_blocks, const uint64_t *block_ptrs_host, const uint32_t *block_widths_host,
                      const int32_t *row_idx, uint32_t U, const float *packed, uint32_t W, void *stream);

/* Push form of the sparse exchange over NVLink / NVSwitch symmetric memory (no NCCL, no host round
 * trip): for every Gaussian g with marks[g] != 0, the row g of each private block s ([N, width_s] at
 * src_ptrs_host[s]) is ADDED into every rank's result buffer at float offset dst_offsets_host[s] +
 * g*width_s -- with multimem.red.add on multicast_result when it is not NULL, else with one
 * red.global.add per entry of peer_result_ptrs_host[0..n_peers) (this rank included) -- and
 * union marks byte g is set on every rank (peer_union_ptrs_host, may be NULL).  The result buffers must
 * be zero in those rows and a cross-rank barrier passed before the launch; after a second barrier
 * every result buffer holds the sum over ranks.  gs3d_rows_zero_marked clears the marked rows of the
 * given blocks (and, clear_marks != 0, the marks): the per-step reset of both buffers. */
int gs3d_rows_push_marked(const uint8_t *marks, uint32_t N, int n_blocks, const uint64_t *src_ptrs_host,
                          const uint32_t *block_widths_host, const uint64_t *dst_offsets_host,
                          const uint64_t *peer_result_ptrs_host, const uint64_t *peer_union_ptrs_host,
                          int n_peers, void *multicast_result, void *stream);
int gs3d_rows_zero_marked(uint8_t *marks, uint32_t N, int n_blocks, const uint64_t *block_ptrs_host,
                          const uint32_t *block_widths_host, int clear_marks, void *stream);

/* Pull form (a sparse NVLS all-reduce; what scales to 8 ranks): gs3d_marks_broadcast sets byte g in
 * every rank's union marks for each marked g (multicast_union != NULL: the NVSwitch ORs whole 4-mark words
 * into all ranks with one multimem.red.or.b32 each; else one byte store per mark and peer); after a cross-rank barrier, gs3d_rows_pull_marked on rank
 * `rank` walks every n_peers-th 512-row chunk of its union marks and, per marked row, reads the sum of
 * ALL ranks' private rows (multimem.ld_reduce.add on multicast_private: reduced inside the NVSwitch;
 * without multicast, peer loads) and stores it into ALL ranks' result buffers (multimem.st on
 * multicast_result; else peer stores).  Private and result buffers share one layout: block s of width
 * block_widths_host[s] starts at float offset block_offsets_host[s].  A second barrier completes it. */
int gs3d_marks_broadcast(const uint8_t *marks, uint32_t N, const uint64_t *peer_union_ptrs_host, int n_peers,
                         void *multicast_union, void *stream);
int gs3d_rows_pull_marked(uint8_t *union_marks, uint32_t N, int n_blocks, const uint32_t *block_widths_host,
                          const uint64_t *block_offsets_host, const uint64_t *peer_private_ptrs_host,
                          const uint64_t *peer_result_ptrs_host, int n_peers, int rank,
                          const void *multicast_private, void *multicast_result, void *stream);

/* ---- legacy RGB path: tile_based_vol_rendering_start_end and its backward,
 * bindings.cpp:29-33 / render.cu:399-481 / vol_render.h:716-923 (batch loops :169-250, :252-352), behind
 * gs/renderer.py:536-671 `render_start_end` and GaussianRenderer.render_aabb_culling (:1219-1305).
 * Same kernels as the SH path in RGB mode: color [M,3] is composited as it is (no SH basis, no sigmoid,
 * no NaN guards), grad_color += a*T*G*grad_out, and -- exact_decisions != 0 -- skip decisions near the
 * 1/255 threshold are taken with the reference's FP64 Gaussian (kernel_gaussian_2d, kernels.h:195-214).
 * records as for the SH entry points (gs3d_pack_records / gs3d_project_cull_fused); out pre-zeroed by the
 * caller (renderer.py:558); gradients are ACCUMULATED into pre-zeroed buffers (renderer.py:609-612). */
int gs3d_composite_rgb_forward(uint32_t M, const float *records, const float *color, const int32_t *start,
                               const int32_t *end, const int32_t *gaussian_ids, float *out, const float *topleft,
                               uint32_t tile_size, uint32_t n_tiles_h, uint32_t n_tiles_w, float pixel_size_x,
                               float pixel_size_y, uint32_t H, uint32_t W, float thresh, int exact_decisions,
                               void *stream);
int gs3d_composite_rgb_backward(uint32_t M, const float *records, const float *color, const int32_t *start,
                                const int32_t *end, const int32_t *gaussian_ids, const float *out,
                                const float *grad_out, float *grad_mean2d, float *grad_cov2d, float *grad_color,
                                float *grad_alpha, const float *topleft, uint32_t tile_size, uint32_t n_tiles_h,
                                uint32_t n_tiles_w, float pixel_size_x, float pixel_size_y, uint32_t H, uint32_t W,
                                float thresh, int exact_decisions, void *stream);

/* ---- a9 + a10 fused: chain rule from (grad_mean2d, grad_cov2d, grad_alpha) to the leaf
 * parameters through projection (Q6) and the activations (sh_renderer.py:318-324), for the
 * Gaussians with mask != 0 (others get zero gradient; mask NULL = all), plus the ADC accumulator
 * of sh_renderer.py:602-623: when adc_mode != 0,
 * grad_mean_acc[i] = max(., ||grad_mean2d_i||) (adc_mode 1, split_reduction "max") or
 * += ||grad_mean2d_i|| (adc_mode 2, "mean") for split_type "2d_mean_grad".
 * Leaf gradients are OVERWRITTEN, or -- accumulate != 0, used when several views share one
 * gradient buffer (view-sharded training) -- ADDED for the Gaussians with mask != 0.
 * accumulate == 2: same as 1, and `mask` is a SPARSE row filter (the `touched` marks of
 * gs3d_composite_sh_backward_peers: a few percent of the rows): the marked rows are compacted per warp first so
 * that all lanes work (one thread per Gaussian would run the chain rule with one live lane per warp). */
int gs3d_project_backward_fused(uint32_t N, const uint8_t *mask, const float *mean,
                                const float *qvec, const float *svec_param,
                                const float *alpha_param, int svec_act, int alpha_act,
                                const float *c2w, int detach_depth, const float *grad_mean2d,
                                const float *grad_cov2d, const float *grad_alpha,
                                float *grad_mean, float *grad_qvec, float *grad_svec_param,
                                float *grad_alpha_param, float *grad_mean_acc, int adc_mode,
                                int accumulate, void *stream);

/* ---- tile-sharded render (new capability, SURVEY.md 8e / cfg 5): tiles are independent after
 * projection, so rank r bins and composites the tile rows [row_begin, row_end) of its band only.
 * gs3d_row_duplicate_counts: duplicates per tile row, row_counts int64 [n_tiles_h] (device), from the
 * rects of gs/culling.py:22-31 -- what the bands are balanced by (scratch >= 8*(n_tiles_h+1) bytes).
 * gs3d_clip_rects_to_rows: clips every rect to the band and compacts the Gaussians that still cover a
 * tile, in ascending index order (deterministic): tl_out/br_out [M',2], depth_out [M'], index_out [M']
 * (original index), all caller-allocated for N entries; counts_host[0] = M', counts_host[1] = the
 * band's duplicate count.  The band is then binned over M' Gaussians (gs3d_tile_culling_aabb_start_end)
 * and the resulting gaussian_ids are mapped back through index_out. */
int gs3d_row_duplicate_counts(uint32_t N, const int32_t *aabb_topleft, const int32_t *aabb_bottomright,
                              uint32_t n_tiles_h, int64_t *row_counts, void *scratch, size_t scratch_bytes,
                              void *stream);
size_t gs3d_clip_scratch_bytes(uint32_t N);
int gs3d_clip_rects_to_rows(uint32_t N, const int32_t *aabb_topleft, const int32_t *aabb_bottomright,
                            const float *depth, int row_begin, int row_end, int32_t *tl_out,
                            int32_t *br_out, float *depth_out, int32_t *index_out, int64_t *counts_host,
                            void *scratch, size_t scratch_bytes, void *stream);

/* ---- measurement aid: when set (device pointer to FOUR zero-initialised uint64 counters, NULL to
 * disable), every compositing launch adds the number of duplicates it actually STAGED into shared
 * memory to counters[0] (forward) / counters[1] (backward) -- one atomic per tile -- and the number of
 * CONTRIBUTING (pixel, Gaussian) pairs (alpha*G >= 1/255, pixel not yet saturated) to counters[2] (forward) /
 * counters[3] (backward).  Tiles stop staging once all their pixels are saturated, so these are the unit
 * counts behind bench.py's rooflines (bytes per staged duplicate, FLOP per pair).  Launches made while the
 * counters are set run an instrumented instantiation of the same kernels. */
int gs3d_set_stage_counters(uint64_t *counters);

/* ======== the steps either side of the rasteriser in the training loop ======== */

/* ---- `opt.step()` (main_sh.py:193) of the torch.optim.Adam built by SHRenderer.get_optimizer
 * (sh_renderer.py:720-729: one group per parameter tensor with its own lr, betas (0.9, 0.99), eps 1e-8,
 * no weight decay, no amsgrad), all groups in ONE launch.  Arithmetic follows torch's
 * `_single_tensor_adam` (lerp, mul/addcmul, sqrt/div/add, addcdiv); bias corrections are formed on the
 * host in double like torch's Python scalars.  `step` counts from 1 and is the value AFTER the increment.
 * state_mode 0: general step (moments read and written, 28 B/element);
 * state_mode 1: step == 1, moments are taken as zero and written but not read (20 B/element);
 * state_mode 2: step == 1, moments neither read nor written, exp_avg/exp_avg_sq may be NULL
 *               (12 B/element) -- the reference re-creates the optimiser after every step
 *               (main_sh.py:238), so every step it ever takes is such a first step. */
typedef struct gs3d_adam_segment {
  float *param;
  const float *grad;
  float *exp_avg;
  float *exp_avg_sq;
  uint64_t n; /* elements */
  double lr;
} gs3d_adam_segment;
int gs3d_adam_step(int n_segments, const gs3d_adam_segment *segments_host, double beta1, double beta2,
                   double eps, uint32_t step, int state_mode, void *stream);

/* ---- adaptive density control: split_gaussians / remove_low_alpha_gaussians / select_masked_gaussians
 * (sh_renderer.py:426-600, 731-741) as classify -> plan (deterministic block scan) -> apply (one fused row
 * mover).  Classes: */
#define GS3D_ADC_KEEP 0  /* row stays */
#define GS3D_ADC_CLONE 1 /* row stays and is copied once after the kept rows */
#define GS3D_ADC_SPLIT 2 /* row is replaced by two samples after the clones */
#define GS3D_ADC_DROP 3  /* row is removed */
/* sh_renderer.py:433-456: hot = grad_mean_acc > pos_grad_thresh (reduction 1, "max") or
 * grad_mean_acc / (cnt + 1e-5) > pos_grad_thresh (reduction 2, "mean"); a hot Gaussian is SPLIT when any
 * activated scale exceeds split_scale_thresh, else CLONE; everything else KEEP.  cls u8 [N]. */
int gs3d_adc_classify(uint32_t N, const float *grad_mean_acc, const int32_t *cnt, int reduction,
                      float pos_grad_thresh, const float *svec_param, int svec_act, float split_scale_thresh,
                      uint8_t *cls, void *stream);
/* sh_renderer.py:542-560: KEEP iff act(alpha_param) >= alpha_thresh, else DROP. */
int gs3d_adc_classify_alpha(uint32_t N, const float *alpha_param, int alpha_act, float alpha_thresh,
                            uint8_t *cls, void *stream);
/* Counts and per-block offsets.  cls_is_keep_mask != 0: cls is a torch.bool keep mask (non-zero = KEEP,
 * zero = DROP; select_masked_gaussians).  counts_host[0] = rows that stay (KEEP + CLONE), [1] = CLONE,
 * [2] = SPLIT, valid on return (one 24-byte D2H + stream sync: the reference's `.sum().item()`,
 * sh_renderer.py:458-459).  The new N is counts[0] + counts[1] + 2*counts[2].  The offsets are left in
 * `scratch` (>= gs3d_adc_scratch_bytes(N) bytes) for gs3d_adc_apply. */
size_t gs3d_adc_scratch_bytes(uint32_t N);
int gs3d_adc_plan(uint32_t N, const uint8_t *cls, int cls_is_keep_mask, int64_t *counts_host, void *scratch,
                  size_t scratch_bytes, void *stream);
/* Writes the new parameter tensors in the reference's order (sh_renderer.py:485-527):
 * [KEEP and CLONE rows in index order | CLONE rows again | first split samples | second split samples].
 * Split samples: mean + R(q)^T (noise * svec), scales svec / scale_shrink_factor through the inverse
 * activation, other fields copied; noise [2*counts[2], 3] standard normal (torch.randn in the caller so
 * that the RNG stream is the reference's).  sh_width = 3 * max_C^2 floats per row. */
int gs3d_adc_apply(uint32_t N, const uint8_t *cls, int cls_is_keep_mask, const void *plan_scratch,
                   const int64_t *counts_host, const float *mean, const float *qvec, const float *svec_param,
                   const float *sh_coeffs, const float *alpha_param, uint32_t sh_width, int svec_act,
                   float scale_shrink_factor, const float *noise, float *mean_out, float *qvec_out,
                   float *svec_param_out, float *sh_coeffs_out, float *alpha_param_out, void *stream);

/* ---- the loss step between the forward and the backward pass: utils/loss.py:5-24,
 *   loss = ssim_mult * ssim_loss(out, gt, window_size, reduction="mean") + (1 - ssim_mult) * base(out, gt)
 * base_loss 1 = F.l1_loss, 2 = F.mse_loss; ssim_loss is kornia's (un-vendored, unpinned -- requirements.txt:5;
 * 0.6.x semantics: Gaussian window sigma 1.5, reflect border, C1 = 0.01^2, C2 = 0.03^2, eps 1e-12,
 * clamp((1 - ssim)/2, 0, 1)).  out / gt are HWC float32 [H, W, 3].  Writes the scalar loss to *loss (device)
 * and, when grad != NULL, d loss / d out to grad [H, W, 3].  window_size odd, <= 11; H, W > window_size/2.
 * scratch >= gs3d_image_loss_scratch_bytes(H, W).  Deterministic (fixed-order reduction). */
size_t gs3d_image_loss_scratch_bytes(uint32_t H, uint32_t W);
int gs3d_image_loss(const float *out, const float *gt, uint32_t H, uint32_t W, int base_loss, float ssim_mult,
                    uint32_t window_size, float *loss, float *grad, void *scratch, size_t scratch_bytes,
                    void *stream);

#ifdef __cplusplus
}
#endif
#endif /* GS3D_B200_H */
