/* posfeat_b200 -- C ABI of the B200 (sm_100a) post-backbone feature pipeline.
 *
 * Drop-in boundary for the hot path of PoSFeat (reference paths are relative
 * to the reference checkout).  The reference has no FFI: its "plugin API" is
 * name-based dispatch on Python callables (managers/extractor.py:87,
 * evaluations/ETH_local_feature/reconstruction_pipeline.py:99).  Each entry
 * point below is what a ctypes binding behind one of those callables needs;
 * the binding itself lives in posfeat_b200/_lib.py and INTEGRATION.md.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer owned by the caller unless the name
 *    ends in _host; the library never allocates or retains device memory;
 *  - all work is enqueued on `stream` (a cudaStream_t passed as void*); no
 *    entry point synchronises the host unless documented;
 *  - return value 0 = OK, otherwise a POSFEAT_E* code; the message is
 *    available from posfeat_last_error() (thread local);
 *  - no C++ exceptions cross this boundary.
 */
#ifndef POSFEAT_B200_H_
#define POSFEAT_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define POSFEAT_OK 0
#define POSFEAT_EINVAL 1     /* bad argument / unsupported option combination */
#define POSFEAT_ECUDA 2      /* a CUDA runtime call failed                     */
#define POSFEAT_EWORKSPACE 3 /* workspace too small                            */
#define POSFEAT_EUNSUPPORTED 4

/* nms_mode: use_nms=False / True of generate_kpts_single
 * (losses/preprocess_utils.py:225-230). */
#define POSFEAT_NMS_NONE 0
#define POSFEAT_NMS_HARD 1
#define POSFEAT_NMS_SOFT 2          /* use_nms='softnms' (soft_nms, losses/preprocess_utils.py:431-447) */
/* or-ed into nms_mode / mode: NMS, threshold and top-k over the WHOLE map, outputs are the
 * pixel's own grid coordinate and score (generate_kpts_single_noavg, :280-336); idx_out then
 * indexes the H x W map and workspaces are sized with (H+2, W+2). */
#define POSFEAT_DETECT_FULLMAP 0x10
/* thr_mode: thr=False, thr_mod='abs'|'max'|'mean' (:232-240). */
#define POSFEAT_THR_NONE 0
#define POSFEAT_THR_ABS 1
#define POSFEAT_THR_MAX 2
#define POSFEAT_THR_MEAN 3
/* descriptor-map layouts for the sampler are expressed as element strides. */

int posfeat_version(void);
/* Copies the calling thread's last error message (NUL terminated). */
int posfeat_last_error(char* buf, int n);
/* Number of SMs of the current device (grid sizing for callers), <0 on error. */
int posfeat_device_sm_count(void);

/* Device-side alias of a page-locked (pinned, mapped) host buffer.  The sampler entry points accept
 * such a pointer for the descriptor map: the kernel then gathers the four taps of every keypoint
 * straight over the host link (2 KB per keypoint at D=128) instead of the caller copying the whole
 * dense map first (34 MB per 896x1200 image) -- the host-buffer form of sample_feat_by_coord,
 * losses/preprocess_utils.py:40-53.  Fails with POSFEAT_EINVAL for pageable memory. */
int posfeat_host_device_pointer(const void* host, void** dev_out);

/* Accounting used by bench.py: kernels launched by this process so far, and
 * optional CUDA-event timing of the individual kernels (events are recorded on
 * the launching stream around each launch while enabled; read() synchronises
 * on them, returns the summed duration and launch count, and resets the slot). */
int64_t posfeat_launch_count(void);
int posfeat_profile_enable(int on);
int posfeat_profile_slot_count(void);
const char* posfeat_profile_slot_name(int slot);
int posfeat_profile_read(int slot, double* total_ms, int32_t* launches);

/* ---- (1) score-map keypoint selection -------------------------------------
 * Replaces generate_kpts_single(..., stable=True), losses/preprocess_utils.py:215-267
 * (nms :449-464, threshold :232-240, centroid/score :243-247, count clamp
 * :249-261, topk+gather :263-267).
 *
 * score      [B,1,H,W] float32, element strides (stride_b, stride_y, 1)
 * num_pts    requested keypoints; 0 = "False" (all survivors)
 * min_pts    the reference's floor (128, :260-261)
 * cap_pts    capacity of the per-image output rows (>= the n that results)
 * n_fixed    >=0: caller already decided n (two-phase use); -1: decide on
 *            device as max(min(num_pts or inf, min_b count_b), min_pts)
 * counts     [B] int32 out: survivors per image (what :251-259 sums)
 * n_out      [1] int32 out: the n used (same for every image of the batch)
 * idx_out    [B,cap_pts] int64: linear index into the (H-2)x(W-2) interior grid
 * kps_out    [B,cap_pts,2] float32 normalised (x,y); kpscore_out [B,cap_pts]
 * Rows [n, cap_pts) are left untouched.  Order: score descending, index
 * ascending inside an equal-score group.
 */
size_t posfeat_detect_workspace_bytes(int B, int H, int W, int cap_pts);

int posfeat_detect_candidates_f32(const float* score, int B, int H, int W,
                                  int64_t stride_b, int64_t stride_y,
                                  int nms_mode, int radius, int thr_mode, float thr,
                                  int32_t* counts, void* workspace, size_t ws_bytes,
                                  void* stream);

int posfeat_detect_select_f32(const float* score, int B, int H, int W,
                              int64_t stride_b, int64_t stride_y, int mode,
                              int num_pts, int min_pts, int cap_pts, int n_fixed,
                              const int32_t* counts, int32_t* n_out,
                              int64_t* idx_out, float* kps_out, float* kpscore_out,
                              void* workspace, size_t ws_bytes, void* stream);

/* candidates + select in one call (n decided on device, no host sync). */
int posfeat_detect_topk_f32(const float* score, int B, int H, int W,
                            int64_t stride_b, int64_t stride_y,
                            int nms_mode, int radius, int thr_mode, float thr,
                            int num_pts, int min_pts, int cap_pts,
                            int32_t* counts, int32_t* n_out,
                            int64_t* idx_out, float* kps_out, float* kpscore_out,
                            void* workspace, size_t ws_bytes, void* stream);

/* Reads the device-side status of the last select on this workspace
 * (synchronises `stream`): POSFEAT_OK, or POSFEAT_EINVAL when n exceeded
 * cap_pts / the number of interior pixels (torch.topk would raise there). */
int posfeat_detect_status(void* workspace, int B, int H, int W, int cap_pts, void* stream);
/* Same, and also copies the device-side n (n_out of the select call) to *n_host in the same
 * host round trip (n_dev / n_host may be NULL). */
int posfeat_detect_finish(void* workspace, int B, int H, int W, int cap_pts,
                          const int32_t* n_dev, int32_t* n_host, void* stream);

/* ---- (2) bilinear descriptor sampling + L2 normalisation -------------------
 * Replaces sample_feat_by_coord, losses/preprocess_utils.py:40-53
 * (grid_sample bilinear/zeros/align_corners=False, then F.normalize, eps 1e-12).
 * fmap element strides (sb, sc, sy, sx) describe NCHW (sx=1) or NHWC (sc=1).
 * coord_n [B,n,2] normalised (x,y), contiguous; out [B,n,D] contiguous.
 * n_valid: optional device int32 (may be NULL) -- only the first *n_valid of
 * the n points are processed (lets the detector's n stay on the device).
 * out_bf16: optional [B,n,D] bf16 copy of the result (may be NULL), the
 * operand format of the tensor-core matcher.
 */
int posfeat_sample_l2norm_f32(const float* fmap, int B, int D, int h, int w,
                              int64_t sb, int64_t sc, int64_t sy, int64_t sx,
                              const float* coord_n, int n, const int32_t* n_valid,
                              int do_norm, float* out, void* out_bf16, void* stream);

/* Pair pipeline: sampling as above for 2P images (pair p = images 2p, 2p+1; D == 128, channels-last
 * map) that also leaves the tensor-core matcher's bf16 operands, row norms and maxima in the workspace
 * of posfeat_mnn_batched_f32(P, N = M = n): one pass over the descriptors instead of two. */
int posfeat_sample_pairs_f32(const float* fmap, int B, int D, int h, int w,
                             int64_t sb, int64_t sc, int64_t sy, int64_t sx,
                             const float* coord_n, int n, int do_norm, float* out,
                             void* mnn_workspace, size_t mnn_ws_bytes, void* stream);

/* Host-buffer callers: sparse host->device staging of exactly the map pixels the sampler will read.
 * fmap_host: device-side alias (posfeat_host_device_pointer) of the pinned dense descriptor map,
 * fmap_dev: device map with the SAME strides; both channels-last (sc == 1), D % 4 == 0.  One bit per
 * pixel covered by some keypoint's 2x2 tap block is set (same tap arithmetic as the sampler), then every
 * marked pixel crosses the host link once (D floats, coalesced) into its place in fmap_dev; unmarked
 * pixels of fmap_dev are left untouched.  posfeat_sample_*_f32 on fmap_dev with the same coord_n then
 * returns what it would return on the whole map.  workspace: bitmap + one 64-bit counter;
 * posfeat_fetch_taps_count reads the number of pixels moved (synchronises the stream). */
size_t posfeat_fetch_taps_workspace_bytes(int B, int h, int w);
int posfeat_fetch_taps_f32(const float* fmap_host, float* fmap_dev, int B, int D, int h, int w,
                           int64_t sb, int64_t sc, int64_t sy, int64_t sx,
                           const float* coord_n, int n, void* workspace, size_t ws_bytes, void* stream);
int posfeat_fetch_taps_count(const void* workspace, int B, int h, int w, unsigned long long* pixels_out,
                             void* stream);

/* Backward of the gather (training path, losses/preprocess.py:56-57): accumulates
 * w_tap * g_out [B,n,D] into g_fmap (same strides as fmap, zero-initialised by
 * the caller; float atomics).  The L2 normalisation is differentiated on the
 * host side (plain tensor ops). */
int posfeat_sample_bwd_f32(const float* g_out, int B, int D, int h, int w,
                           int64_t sb, int64_t sc, int64_t sy, int64_t sx,
                           const float* coord_n, int n, float* g_fmap, void* stream);

/* ---- (3) mutual nearest neighbour matching ---------------------------------
 * Replaces mnn_matcher / mutual_nn_matcher: evaluations/hpatches/evaluation.py:27-38,
 * evaluations/aachen/matchers.py:5-13, evaluations/ETH_local_feature/custom_matcher.py:5-13,
 * losses/preprocess_utils.py:795-803.
 * A [N,D], Bm [M,D] float32 row-major (row stride lda/ldb elements).
 * nn12 [N] int32 = argmax_j <A_i,B_j>, nn21 [M] int32 = argmax_i (first index
 * on exact ties); matches [N,2] int64 holds the K mutual pairs (i, nn12[i]) in
 * ascending i; n_matches [1] int32.  The N x M similarity is never written to
 * memory.  algo: 0 = auto, 1 = exact SIMT kernel (fp64 accumulation),
 * 2 = tcgen05 tensor-core kernel (bf16 operands, candidate rescoring in fp64;
 * requires D == 128).  Both give the same nn12/nn21.
 * nn21 may be NULL: only the match list is produced then (the tensor-core path
 * contracts one direction and verifies mutuality per column chunk; M <= 65536).
 */
#define POSFEAT_MNN_AUTO 0
#define POSFEAT_MNN_SIMT 1
#define POSFEAT_MNN_TC 2
/* or-ed into algo of posfeat_mnn_batched_f32: the workspace already holds the operands
 * posfeat_sample_pairs_f32 prepared for the same (P, N = M) */
#define POSFEAT_MNN_PREPARED 0x100
size_t posfeat_mnn_workspace_bytes(int N, int M, int D, int algo);

int posfeat_mnn_f32(const float* A, int N, int64_t lda, const float* Bm, int M, int64_t ldb,
                    int D, int algo, int32_t* nn12, int32_t* nn21, int64_t* matches,
                    int32_t* n_matches, void* workspace, size_t ws_bytes, void* stream);

/* Batched over P independent pairs of equal shape (the pair pipeline's call):
 * pair p reads A + p*stride_a and Bm + p*stride_b (strides in elements) and
 * writes nn12[p*N..], nn21[p*M..], matches[p*N*2..], n_matches[p].  One launch
 * chain serves all pairs, so launch and prologue costs are paid once. */
size_t posfeat_mnn_batched_workspace_bytes(int P, int N, int M, int D, int algo);
int posfeat_mnn_batched_f32(const float* A, int64_t stride_a, int N, int64_t lda,
                            const float* Bm, int64_t stride_b, int M, int64_t ldb,
                            int D, int P, int algo, int32_t* nn12, int32_t* nn21,
                            int64_t* matches, int32_t* n_matches,
                            void* workspace, size_t ws_bytes, void* stream);

/* Ratio-test matchers: ratio_matcher / mutual_nn_ratio_matcher,
 * evaluations/aachen/matchers.py:17-75 (ETH copy custom_matcher.py:16-73).  Top-2
 * similarities per row and per column (exact kernel), dist = sqrt(2 - 2 sim),
 * ratio = d0 / (d1 + 1e-8); row i is kept iff ratio12[i] <= ratio and
 * ratio21[nn12[i]] <= ratio, and for mutual != 0 also nn21[nn12[i]] == i.
 * N, M >= 2 (torch.topk(2) raises otherwise). */
size_t posfeat_ratio_match_workspace_bytes(int N, int M, int D);
int posfeat_ratio_match_f32(const float* A, int N, int64_t lda, const float* Bm, int M, int64_t ldb,
                            int D, float ratio, int mutual, int32_t* nn12, int64_t* matches,
                            int32_t* n_matches, void* workspace, size_t ws_bytes, void* stream);

/* Same, operands and results in HOST memory (pageable or pinned): copies in,
 * runs, copies back and synchronises `stream`.  This is the call a reader of
 * .npz descriptor files makes (evaluations/hpatches/evaluation.py:64-67).
 * Device scratch [dev_scratch, +scratch_bytes) is supplied by the caller;
 * query the size with posfeat_mnn_host_scratch_bytes. */
size_t posfeat_mnn_host_scratch_bytes(int N, int M, int D, int algo);
int posfeat_mnn_host_f32(const float* A_host, int N, const float* B_host, int M, int D, int algo,
                         int64_t* matches_host, int32_t* n_matches_host,
                         void* dev_scratch, size_t scratch_bytes, void* stream);

/* ---- (4) training-side correlation + softmax expectation -------------------
 * Dense variant: get_expected_correspondence_locs, losses/preprocess_utils.py:55-82,
 * and the grid<->grid stage of Preprocess_Line2Window.forward,
 * losses/preprocess.py:59-81 (both are softmax(scale*<q,k>) expectations of a
 * coordinate table).
 *   q [B,n,D], k [B,m,D] (row strides D), v [B or 1, m, C] value table
 *   (C <= 4: e.g. x, y, x^2, y^2); out [B,n,C] = sum_j softmax_j(scale*q.k_j) v_j;
 *   lse [B,n] = log-sum-exp of the scaled logits (saved for backward).
 * Backward: given g_out [B,n,C] produces g_q [B,n,D] and/or g_k [B,m,D] (either
 * may be NULL); the logits are recomputed from (q, k, lse); no atomics.
 * D <= 128, C <= 4.
 * Forward: problems with m >= 128 and B*n*m >= 2^18 run on the tensor cores (tcgen05
 * kind::tf32 with hi/lo operand splitting, float32-accurate logits) and need the
 * workspace posfeat_corr_expect_workspace_bytes reports (256-byte aligned); smaller ones
 * use a SIMT kernel and the workspace may be NULL (the size query returns 0).
 */
size_t posfeat_corr_expect_workspace_bytes(int B, int n, int m, int D, int C);
int posfeat_corr_expect_fwd_f32(const float* q, const float* k, const float* v, int v_batched,
                                int B, int n, int m, int D, int C, float scale,
                                float* out, float* lse, void* workspace, size_t ws_bytes,
                                void* stream);
/* Backward on the same size rule: tensor cores (W = P o (g.v - g.out) materialised as tf32 hi/lo, then two
 * split-K 3xTF32 GEMMs) with the workspace posfeat_corr_expect_bwd_workspace_bytes reports, else SIMT. */
size_t posfeat_corr_expect_bwd_workspace_bytes(int B, int n, int m, int D, int C);
int posfeat_corr_expect_bwd_f32(const float* q, const float* k, const float* v, int v_batched,
                                int B, int n, int m, int D, int C, float scale,
                                const float* out, const float* lse, const float* g_out,
                                float* g_q, float* g_k, void* workspace, size_t ws_bytes, void* stream);

/* compute_prob, losses/preprocess_utils.py:89-115, for callers that want the [B,m,n] probability tensor itself
 * (the three expectations above fuse it away and never write it).  f1 [B,m,D], f2 [B,n,D] contiguous.
 * mode 0 ('cos'): prob = softmax_j(scale * <f1_i,f2_j>), scale = sqrt(n) for with_scale else 1; sim (optional,
 * may be NULL) receives the raw similarities (return_sim).  mode 1 ('euc'): prob = softmax_j(-|f1_i - f2_j|^2)
 * evaluated as |f1|^2 + |f2|^2 - 2<f1,f2> like the reference; sim must be NULL. */
int posfeat_compute_prob_f32(const float* f1, const float* f2, int B, int m, int n, int D, int mode,
                             float scale, float* prob, float* sim, void* stream);

/* scale * F.normalize(x, p=2, dim=1, eps) of a descriptor map, written channels-last in one pass: the operation
 * Preprocess_Line2Window applies to both fine maps before the line search and the window expectation
 * (losses/preprocess.py:56-57 via preprocess_utils.py:40-53 / :118-160, which normalise inside).  x: B images of D
 * channels at stride sc (floats), the H*W pixels of a channel contiguous (NCHW), image stride sb; out: [B][HW][D];
 * norm (optional for the forward call): [B][HW], the L2 norms -- the backward call needs them.  Backward: g is the
 * gradient of out ([B][HW][D]), gx receives the gradient of x at the strides of x. */
int posfeat_normalize_scale_fwd_f32(const float* x, int B, int D, int HW, int64_t sb, int64_t sc, float scale, float eps,
                                    float* out, float* norm, void* stream);
int posfeat_normalize_scale_bwd_f32(const float* g, const float* x, const float* norm, int B, int D, int HW, int64_t sb,
                                    int64_t sc, float scale, float eps, float* gx, void* stream);

/* DiskLoss dense affinity, losses/kploss.py:158-182 (second training stage): with A = T*<q_i,k_j> - T,
 * p_ij = softmax_j(A)_ij * softmax_i(A)_ij and log p_ij the sum of the two log-softmaxes, computes per row i
 *   rows_out[b,i] = { sum_j acc r p (log p + logp_i + logp_j),  sum_j acc r p,  sum_j p,  max_j p }
 * without materialising any [B,n,m] tensor (tcgen05 kind::tf32, hi/lo operand split).  rowtab [B,n,8] /
 * coltab [B,m,8] hold per point {lse, la, lb, lc, x, y, logp, accept}: lse = logsumexp of the point's row
 * (column) of T*<q,k> (from posfeat_corr_expect_fwd_f32 with scale T), (la,lb,lc) its normalised epipolar
 * line in the other image, (x,y) its pixel coordinates, logp / accept from the keypoint sampler.
 * reward r: constant (good/bad by |line.point| < thr on both sides, kploss.py:52-88) or dynamic (:90-129). */
size_t posfeat_dual_softmax_reward_workspace_bytes(int B, int n, int m, int D);
int posfeat_dual_softmax_reward_f32(const float* q, const float* k, const float* rowtab, const float* coltab,
                                    int B, int n, int m, int D, float temperature, float thr_own,
                                    float thr_other, float good_reward, float bad_reward, int dynamic_reward,
                                    float* rows_out, void* workspace, size_t ws_bytes, void* stream);

/* Epipolar line search in one launch: epipolar_line_search, losses/preprocess_utils.py:662-694, with
 * get_endpoints (:697-719) folded in.  coord_px [B,n,2] pixel coordinates in an img_h x img_w image,
 * Fmat [B,3,3].  Per query: the epipolar line is clipped to the image rectangle (ends [B,n,4] =
 * normalised x1,y1,x2,y2; valid [B,n] = exactly two border intersections inside), line_step positions
 * between the endpoints are sampled from fmap (bilinear, border padding), dotted with q and soft-maxed.
 * exp_soft [B,n,2] soft expectation; nn_xy [B,n,2] sum of the positions attaining the largest
 * probability; m2 [B,n,2] = sum_p prob_p * pos_p^2; prob [B,n,line_step] optional (may be NULL). */
int posfeat_line_search_f32(const float* fmap, int B, int D, int h, int w,
                            int64_t sb, int64_t sc, int64_t sy, int64_t sx,
                            const float* q, const float* coord_px, const float* Fmat, int n,
                            int img_h, int img_w, int line_step, float* ends, unsigned char* valid,
                            float* exp_soft, float* nn_xy, float* m2, float* prob, void* stream);

/* Window / line variant.  mode 0: get_expected_correspondence_within_window,
 * losses/preprocess_utils.py:721-758 -- positions = centre [B,n,2] + offsets [m,2]
 * (the gen_grid(-ws,ws,...) table), grid_sample padding 'zeros'.  mode 1: the
 * sampling half of epipolar_line_search, :668-675 -- centre holds two endpoints
 * [B,n,4] = (x1,y1,x2,y2), positions = e1 + (e2-e1)*linspace(0,1,m), padding
 * 'border'.  fmap [B,D,h,w] with element strides (sb,sc,sy,sx); q [B,n,D].  The
 * [B,n,m,D] gathered tensor is never materialised.  Outputs: exp_xy [B,n,2] =
 * sum_p prob_p pos_p, std [B,n] = sum_xy sqrt(clamp(var, 1e-10)), prob [B,n,m]
 * (may be NULL), lse [B,n].
 * Backward (mode 0): given g_exp [B,n,2] and g_std [B,n] produces g_q [B,n,D]
 * and accumulates into g_fmap (same strides as fmap, zero-initialised by the
 * caller, float atomics -- as ATen's grid_sampler backward does).
 */
int posfeat_window_expect_fwd_f32(const float* fmap, int B, int D, int h, int w,
                                  int64_t sb, int64_t sc, int64_t sy, int64_t sx,
                                  const float* q, const float* centre, int n,
                                  const float* offsets, int m, int mode,
                                  float* exp_xy, float* std_out, float* prob, float* lse,
                                  void* stream);
int posfeat_window_expect_bwd_f32(const float* fmap, int B, int D, int h, int w,
                                  int64_t sb, int64_t sc, int64_t sy, int64_t sx,
                                  const float* q, const float* centre, int n,
                                  const float* offsets, int m,
                                  const float* exp_xy, const float* prob,
                                  const float* g_exp, const float* g_std,
                                  float* g_q, float* g_fmap, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* POSFEAT_B200_H_ */
