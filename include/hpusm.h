/*
 * hpusm.h -- C ABI of libhpusm.so, the B200 (sm_100a) matching hot path.
 *
 * The reference (gilbeckers/HumanPose-and-UrbanScene-Matching) has NO native/FFI layer: its hot path
 * calls OpenCV / NumPy binaries from Python.  Each entry point below therefore replaces one of those
 * third-party call sites; the reference file:line of the call site is cited per function.  The Python
 * modules under humanpose-and-urbanscene-matching_b200/dropin/ keep the reference's own signatures and
 * bind these symbols with ctypes (see INTEGRATION.md for the stub a maintainer of the reference adds).
 *
 * Conventions (all functions):
 *   - every data pointer is a DEVICE pointer owned by the caller (the *_host conveniences say so);
 *   - no allocation escapes the library: scratch is passed in (`ws`, `ws_bytes`), sized by the
 *     matching hpusm_*_workspace_bytes() query.  `ws` must be 256-byte aligned;
 *   - work is enqueued asynchronously on `stream` (a cudaStream_t passed as void*; NULL = legacy
 *     default stream); outputs are valid after the stream is synchronised;
 *   - return value: HPUSM_OK (0) or a negative HPUSM_ERR_* code; hpusm_last_error() returns a
 *     thread-local message for the last failure on the calling thread;
 *   - thread-safe for distinct (stream, workspace) pairs;
 *   - there is no CPU fallback: without a CUDA device every compute call returns HPUSM_ERR_CUDA.
 */
#ifndef HPUSM_H_
#define HPUSM_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HPUSM_ABI_VERSION 202

#define HPUSM_OK 0
#define HPUSM_ERR_INVALID (-1)     /* bad argument (null pointer, negative size, unsupported width) */
#define HPUSM_ERR_CUDA (-2)        /* CUDA runtime/driver error, see hpusm_last_error()             */
#define HPUSM_ERR_WORKSPACE (-3)   /* workspace missing, misaligned or too small                    */
#define HPUSM_ERR_UNSUPPORTED (-4) /* shape outside what the sm_100a kernels were built for         */

/* L2 kNN engine selection (hpusm_knn2_l2 `mode`). */
#define HPUSM_L2_AUTO 0   /* tensor-core path; engine chosen from the data (u8 integer MMA, else bf16 exact-int, else
                             the fixed-point integer engine HPUSM_L2_TC_Q16); reads a 4-byte flag back per decision, i.e. synchronises the stream      */
#define HPUSM_L2_SIMT 1   /* exact FP32 CUDA-core brute force (cross-check / tiny inputs)             */
#define HPUSM_L2_TC_INT 2 /* tcgen05, bf16 operands, integer values in [0,256] (checked on the device: a violation leaves
                             every idx at -1 and is reported by hpusm_async_status())                                   */
#define HPUSM_L2_TC_SPLIT 3 /* tcgen05, hi/lo bf16 split (general float), top-4 + verified FP32 re-rank; synchronises
                               the stream once (count of rows sent to the exact fallback)                          */
#define HPUSM_L2_TC_I8 4  /* tcgen05 kind::i8, u8 operands, s32 accumulation: integer values in [0,255] and squared
                             norms below 4,000,000.  Fully asynchronous: the range check runs on the device and gates
                             the sweep there; a violation leaves every idx at -1 and is reported by hpusm_async_status() */
#define HPUSM_L2_TC_Q16 5 /* tcgen05 kind::i8 for GENERAL floats: 15-bit fixed point split into two signed bytes, two s32
                             accumulators per tile (13 integer MMAs of K = 32 where the bf16 split needs 25 of K = 16),
                             top-6 per list + verified FP32 re-rank with error bounds measured on the data; synchronises
                             the stream once (count of rows sent to the exact fallback)                             */

/* Engines of the segmented Hamming matcher (hpusm_knn2_hamming_segmented). */
#define HPUSM_HAMMING_AUTO 0 /* tensor-core engine where its shape limits hold, else the __popc engine */
#define HPUSM_HAMMING_POPC 1 /* XOR + POPC on the integer pipes */
#define HPUSM_HAMMING_TC 2   /* tcgen05 kind::i8 on signed expansions of the descriptor bits */

int hpusm_version(void);
const char* hpusm_last_error(void);
/* Number of CUDA devices visible, or a negative error code. */
int hpusm_device_count(void);
/* Running count of kernels this library launched from the calling process (for bench `gpu_launches`). */
int64_t hpusm_launch_count(void);
/* Entry points that stay asynchronous where a data-dependent precondition can only be checked on the device (explicit
 * HPUSM_L2_TC_I8) record a violation in a pinned host word instead of reading a flag back.  Call after synchronising the
 * stream: returns HPUSM_OK or the recorded HPUSM_ERR_* code (and clears it); the affected call's outputs are all idx -1. */
int hpusm_async_status(void);

/* Event bracketing of the dominant kernel of each path (the kNN sweep, the RANSAC scoring kernel, the pose
 * kernel), for roofline accounting: while enabled every such launch is bracketed by two CUDA events on its
 * own stream.  hpusm_profile_collect() synchronises on them, returns the summed duration and the number of
 * bracketed launches since the previous collect (HOST pointers), and clears the list. */
int hpusm_profile_enable(int on);
int hpusm_profile_collect(double* total_ms, int64_t* launches);

/* --------------------------------------------------------------------------------------------------
 * K2  Binary-descriptor kNN, k=2, Hamming norm.
 * Replaces cv2.BFMatcher(cv2.NORM_HAMMING).knnMatch(desc_model, trainDescriptors=desc_input, k=2)
 * (reference urbanscene/features.py:42 and :59).
 *   q [nq][nbytes], t [nt][nbytes] uint8, C-contiguous, 16-byte aligned base, nbytes in {32, 64}.
 *   idx  [nq][2] int32: train index of the nearest / second nearest; -1 where nt < 2 leaves a slot empty.
 *   dist [nq][2] int32: popcount(q ^ t).  Ties resolve to the LOWER train index (OpenCV behaviour).
 * ------------------------------------------------------------------------------------------------ */
size_t hpusm_knn2_hamming_workspace_bytes(int64_t nq, int64_t nt, int nbytes);
int hpusm_knn2_hamming(const uint8_t* q, int64_t nq, const uint8_t* t, int64_t nt, int nbytes,
                       int32_t* idx, int32_t* dist, void* ws, size_t ws_bytes, void* stream);

/* Segmented variant for query-vs-database retrieval (dataset/dataset_urbanscene.py:46-84 loops
 * match_whole over image pairs; every (query image, database image) pair is an independent kNN+ratio).
 *   t holds the descriptors of nseg database images back to back; seg_offsets [nseg+1] int64 (device)
 *   gives each image's row range.  For every segment s and query row i the kernel finds the two nearest
 *   rows INSIDE segment s, applies Lowe's ratio test (features.py:48) in double precision and counts
 *   survivors:  score[s] = #{ i : seg has >= 2 rows and (double)d0 < (double)d1 * ratio }.
 *   score [nseg] int32 is overwritten.
 *   best_idx (optional, may be NULL) [nseg][nq] int32: global row index of the nearest neighbour where
 *   the ratio test passed, else -1 (feeds the RANSAC stage without a second pass).
 *   `engine` is one of HPUSM_HAMMING_*: the __popc engine (XOR + POPC on the integer pipes, any shape), or the tensor-core
 *   engine (tcgen05 kind::i8 on +-1 / -+64 expansions of the bits, s32 accumulators that hold 128 * Hamming + column; built
 *   for 32-byte descriptors, >= 128 queries, >= 16384 database rows and images of >= 256 rows on average -- other shapes
 *   return HPUSM_ERR_UNSUPPORTED); HPUSM_HAMMING_AUTO takes the tensor-core engine where it applies.  Both give the same
 *   score and best_idx, bit for bit.  The workspace (256-byte aligned) holds the expanded operands of the tensor engine:
 *   288 bytes per database row, every image padded to a multiple of 128 rows. */
size_t hpusm_knn2_hamming_segmented_workspace_bytes(int64_t nq, int64_t nt, int64_t nseg, int nbytes, int engine);
int hpusm_knn2_hamming_segmented(const uint8_t* q, int64_t nq, const uint8_t* t, int64_t nt,
                                 const int64_t* seg_offsets, int64_t nseg, int nbytes, double ratio, int engine,
                                 int32_t* score, int32_t* best_idx, void* ws, size_t ws_bytes,
                                 void* stream);

/* --------------------------------------------------------------------------------------------------
 * K1  Float-descriptor kNN, k=2, L2 norm.
 * Replaces cv2.BFMatcher(cv2.NORM_L2).knnMatch(..., k=2) (features.py:42,:59).
 *   q [nq][d], t [nt][d] float32, C-contiguous, 16-byte aligned, d a multiple of 4, d <= 512.
 *   The tensor-core engines are built for d == 128 (SIFT); HPUSM_L2_AUTO zero-pads narrower rows (SURF: d = 64) to 128
 *   columns inside the workspace when nq * nt >= 2^24 (same answer bit for bit) and takes the exact FP32 engine otherwise;
 *   HPUSM_L2_TC_INT / HPUSM_L2_TC_SPLIT / HPUSM_L2_TC_I8 / HPUSM_L2_TC_Q16 return HPUSM_ERR_UNSUPPORTED for d != 128.
 *   HPUSM_L2_TC_INT / HPUSM_L2_TC_I8 never synchronise: their value-range check runs on the device (see hpusm_async_status);
 *   HPUSM_L2_AUTO checks the value range on the device: integers in [0,255] (what OpenCV SIFT emits) run on the u8
 *   integer tensor pipe, integers in [0,256] on the bf16 engine, anything else (mixed signs included) on the fixed-point engine HPUSM_L2_TC_Q16, whose
 *   answer is the exact FP32 one as well (candidate lists are proven complete or the row is redone by the FP32 engine).
 *   idx  [nq][2] int32 as above.
 *   dist [nq][2] float32 = sqrtf( sum_k (q_k - t_k)^2 ), the sum taken in FP32 in increasing k with
 *   fused multiply-add (the "exact FP32 re-rank"); this is what DMatch.distance holds in the reference.
 *   Ties (equal FP32 distance) resolve to the lower train index.
 *   `mode` is one of HPUSM_L2_*.
 * ------------------------------------------------------------------------------------------------ */
size_t hpusm_knn2_l2_workspace_bytes(int64_t nq, int64_t nt, int d, int mode);
int hpusm_knn2_l2(const float* q, int64_t nq, const float* t, int64_t nt, int d, int mode,
                  int32_t* idx, float* dist, void* ws, size_t ws_bytes, void* stream);
/* Which engine the last hpusm_knn2_l2 call on this thread resolved HPUSM_L2_AUTO to (HPUSM_L2_*),
 * and how many query rows needed the exact brute-force fallback after candidate verification. */
int hpusm_knn2_l2_last_mode(void);
int64_t hpusm_knn2_l2_last_fallback_rows(void);

/* Lexicographic (distance, index) merge of `nparts` partial top-2 lists (DB shards or DB splits):
 *   part_idx/part_dist [nparts][nq][2]; empty slots carry idx -1.  Used after the NCCL all-gather of
 *   per-rank results; indices must already be global.  In-place safe when nparts == 1. */
int hpusm_knn2_merge_f32(const int32_t* part_idx, const float* part_dist, int nparts, int64_t nq,
                         int32_t* idx, float* dist, void* stream);
int hpusm_knn2_merge_i32(const int32_t* part_idx, const int32_t* part_dist, int nparts, int64_t nq,
                         int32_t* idx, int32_t* dist, void* stream);

/* Lowe ratio test, replaces the loop of features.filter_matches (features.py:45-55):
 *   keep[i] = idx[i][1] >= 0 && (double)dist[i][0] < (double)dist[i][1] * ratio      (features.py:48)
 *   *count (device int32) receives the number of survivors.  dist_is_int selects int32 vs float32. */
int hpusm_ratio_filter(const void* dist, const int32_t* idx, int64_t nq, int dist_is_int, double ratio,
                       uint8_t* keep, int32_t* count, void* stream);

/* Order-preserving compaction of the ratio survivors + keypoint gather (features.py:49-53):
 *   kp_q [nq][2], kp_t [nt][2] float32 keypoint coordinates; for the j-th survivor (increasing query
 *   row) p_q[j] = kp_q[i], p_t[j] = kp_t[idx[i][0]], sel[j] = i.  *count as above.
 *   ws: hpusm_gather_matches_workspace_bytes(nq). */
size_t hpusm_gather_matches_workspace_bytes(int64_t nq);
int hpusm_gather_matches(const float* kp_q, const float* kp_t, const int32_t* idx, const uint8_t* keep,
                         int64_t nq, float* p_q, float* p_t, int32_t* sel, int32_t* count, void* ws,
                         size_t ws_bytes, void* stream);

/* Per-image match lists for batched RANSAC after a segmented retrieval pass (the loop of
 * dataset/dataset_urbanscene.py:46-84 over features.match_and_draw, features.py:59-64):
 *   best_idx [nseg][nq] and count [nseg] as written by hpusm_knn2_hamming_segmented; kp_q [nq][2], kp_t [nt][2] keypoints.
 *   Images with fewer than min_count survivors (thresholds.MIN_MATCH_COUNT = 16, features.py:64) contribute no matches.
 *   pair_offsets [nseg+1] int64 (device) receives the scan; src/dst must hold sum(count) rows (<= nseg*nq);
 *   src = query keypoint, dst = database keypoint of each surviving match in increasing query order; sel (optional)
 *   the query row.  Feeds hpusm_ransac_homography_batch directly. */
int hpusm_segment_matches(const int32_t* best_idx, const int32_t* count, int64_t nseg, int64_t nq, const float* kp_q,
                          const float* kp_t, int min_count, int64_t* pair_offsets, float* src, float* dst, int32_t* sel,
                          void* stream);

/* features.validate_homography (features.py:177-211) on a batch: H [n][9] float64 -> valid [n] uint8 (NaN = no model = 0). */
int hpusm_validate_homography_batch(const double* H, int64_t n, uint8_t* valid, void* stream);

/* --------------------------------------------------------------------------------------------------
 * K3  Batched RANSAC homography.
 * Replaces cv2.findHomography(src, dst, cv2.RANSAC, thresh) (features.py:66 and :69), batched over
 * image pairs.  Pair p owns matches [pair_offsets[p], pair_offsets[p+1]).
 *   src, dst [total][2] float32; pair_offsets [npairs+1] int64 (device).
 *   nhyp      RANSAC iterations per pair (cv2 maxIters; reference default 2000).  As in OpenCV every iteration redraws
 *             its 4-point sample until HomographyEstimatorCallback::checkSubset accepts it, so every iteration scores a model.
 *   confidence  > 0: reproduce the sequential early-exit of RANSAC exactly (iteration h is considered
 *               only while h < the adaptive iteration bound, cv2 default 0.995); <= 0: use all nhyp.
 *   sampler   HPUSM_RANSAC_SAMPLER_CV2: OpenCV's own fixed-seed sample sequence (its RNG cannot be seeded from Python and
 *               ignores cv2.setRNGSeed) replayed by a sequential per-pair pre-pass, and OpenCV's float32 computeError as
 *               the inlier test: mask and H are those of cv2.findHomography (identical masks, H to 1e-4 and better; pinned
 *               on 13 real image pairs and 1000 synthetic ones, tests/golden/ransac_*.npz).  `seed`/`pair_index0` are unused.
 *             HPUSM_RANSAC_SAMPLER_COUNTER: counter-based sample table keyed on (seed, pair_index0 + p, iteration, attempt),
 *               drawn by all lanes in parallel, division-free inlier test; shared bit for bit with oracle/ransac_oracle.c.
 *               A batch cut into chunks or sharded over GPUs passes its first pair's global index as pair_index0 and
 *               draws exactly the samples of the one-call run.
 *   H    [npairs][9] float64 row-major, H[8] == 1; all NaN when no model with >= 4 inliers exists.
 *   mask [total] uint8 inlier flags: with refine != 0 the inliers of the REFINED homography under cv2's float32
 *               reprojection error (what cv2 4.x returns); otherwise those of the winning minimal-sample model.
 *   ninl [npairs] int32 = number of set mask entries; best_hyp [npairs] int32 (optional, may be NULL) winning iteration.
 *   refine != 0: findHomography's post-processing -- runKernel's normalised DLT on the inliers of the winning model, then
 *               LMSolver's <= 10 Levenberg-Marquardt iterations on the reprojection error; refine == 0 returns the winning
 *               minimal-sample model.
 * ------------------------------------------------------------------------------------------------ */
#define HPUSM_RANSAC_SAMPLER_COUNTER 0
#define HPUSM_RANSAC_SAMPLER_CV2 1
size_t hpusm_ransac_homography_workspace_bytes(int64_t total_matches, int64_t npairs, int nhyp, int sampler);
int hpusm_ransac_homography_batch(const float* src, const float* dst, const int64_t* pair_offsets,
                                  int64_t npairs, int64_t total_matches, float thresh, int nhyp,
                                  double confidence, int sampler, uint64_t seed, int64_t pair_index0, int refine,
                                  double* H, uint8_t* mask, int32_t* ninl, int32_t* best_hyp, void* ws,
                                  size_t ws_bytes, void* stream);
/* Scoring only (the C4 benchmark kernel): counts [npairs][nhyp] int32 inlier count of every iteration's model,
 * -1 where an iteration found no acceptable sample (degenerate input).  ws is only needed by the cv2 sampler. */
size_t hpusm_ransac_score_workspace_bytes(int64_t total_matches, int64_t npairs, int nhyp, int sampler);
int hpusm_ransac_score_batch(const float* src, const float* dst, const int64_t* pair_offsets,
                             int64_t npairs, int64_t total_matches, float thresh, int nhyp, int sampler,
                             uint64_t seed, int64_t pair_index0, int32_t* counts, void* ws, size_t ws_bytes,
                             void* stream);

/* Projective transform of points, replaces cv2.perspectiveTransform (features.py:77,
 * transformation.py:38): pts [n][2] float32, H [9] float64 (device) -> out [n][2] float32. */
int hpusm_perspective_transform(const float* pts, int64_t n, const double* H, float* out, void* stream);

/* --------------------------------------------------------------------------------------------------
 * K4  Pose scoring.
 * ------------------------------------------------------------------------------------------------ */

/* Batched affine least squares, replaces common.find_transformation (common.py:37-101, np.linalg.lstsq
 * at :84).  Problem b: model/input [n][npts][2] float64.  Rows whose INPUT point is (0,0) are left out
 * of the fit and come back as (0,0); a problem with no rows left returns the input unchanged and
 * has_A[b] = 0 (the reference returns an empty list).  Rank-deficient fits get the minimum-norm
 * solution (what SVD-based lstsq returns).
 *   out_transform [n][npts][2] float64; A [n][9] float64 row-major, row-vector convention
 *   ([x y 1] * A), entries with |a| < 1e-10 zeroed AFTER the transform was applied (common.py:99). */
int hpusm_affine_lsq_batch(const double* model, const double* input, int64_t n, int npts,
                           double* out_transform, double* A, uint8_t* has_A, void* stream);

/* Batched Procrustes superimposition, replaces posematching/procrustes.py:150-259 (np.linalg.svd :218).
 *   X, Y [n][npts][2] float64.  scaling: 0/1.  reflection: 0 = 'best', 1 = force reflection,
 *   2 = force no reflection.
 *   d [n]; Z [n][npts][2]; tform [n][7] = rotation T row-major (4), scale b, translation c (2). */
int hpusm_procrustes_batch(const double* X, const double* Y, int64_t n, int npts, int scaling,
                           int reflection, double* d, double* Z, double* tform, void* stream);

/* Batched single-person pose decision, replaces posematching/single_person.match_single
 * (single_person.py:80-189) = handle_undetected_points (common.py:281) -> feature_scaling x2
 * (common.py:688) -> split_in_face_legs_torso_v2 (common.py:269) -> find_transformation x3 ->
 * pose_comparison.max_euclidean_distance / decide_* (pose_comparison.py:8-81).
 *   query [18][2] float32, db [n][18][2] float32 (OpenPose COCO order, undetected joints = (0,0)).
 *   query_is_model != 0: the query plays the reference's `model_features`, db rows are `input_features`
 *   (dataset/Multipose_dataset_actions.py:38-49); 0 swaps the roles.
 *   thresholds [5] float64 = {SP_DISTANCE_TORSO, SP_ROTATION_TORSO, SP_DISTANCE_LEGS, SP_ROTATION_LEGS,
 *   SP_DISTANCE_SHOULDER} (thresholds.py:7-13), host pointer.
 *   match [n] uint8, score [n] float32 (= error_score, single_person.py:184).
 *   detail (optional, NULL to skip) [n][DETAIL] float64: torso/legs/face max distance, shoulder
 *   distance, rot torso, rot legs, then the three 3x3 A matrices and the (22,2) transformed input,
 *   HPUSM_POSE_DETAIL doubles per pose.
 *   ws (optional): hpusm_pose_match_workspace_bytes(n).  With a workspace (and detail == NULL, rotation thresholds
 *   below 86 degrees) the database is swept by an FP32 fast path that certifies its own decisions: a pose whose
 *   comparisons fall inside the FP32 error bands, whose part fits are ill conditioned or which needs one of the
 *   reference's corner cases is redone by the exact FP64 kernel, so `match` is the FP64 answer for every pose and
 *   `score` agrees with it to 2e-6.  ws == NULL: the exact FP64 kernel for every pose. */
#define HPUSM_POSE_JOINTS 18
#define HPUSM_POSE_DETAIL (8 + 27 + 44)
size_t hpusm_pose_match_workspace_bytes(int64_t n);
int hpusm_pose_match_batch(const float* query, const float* db, int64_t n, int query_is_model,
                           const double* thresholds, uint8_t* match, float* score, double* detail,
                           void* ws, size_t ws_bytes, void* stream);

/* Same with float64 poses (the reference's own dtype); used by the drop-in single_person.match_single. */
int hpusm_pose_match_batch_f64(const double* query, const double* db, int64_t n, int query_is_model,
                               const double* thresholds, uint8_t* match, float* score, double* detail,
                               void* stream);

/* Arbitrary pose pairs in one launch: pair i = match_single(model[i], input[i]).  This is the candidate stage of
 * posematching/multi_person.py:169-207 (every model person x every input person).  model, input [n][18][2] float32
 * (is_f64 == 0) or float64; outputs as hpusm_pose_match_batch. */
int hpusm_pose_match_pairs(const void* model, const void* input, int64_t n, int is_f64, const double* thresholds,
                           uint8_t* match, float* score, double* detail, void* stream);

/* Scene score of image pairs after RANSAC: urban_scene.match_scene_multi steps 3.1-5 (urban_scene.py:80-145) =
 * transformation.perspective_correction (transformation.py:22-53) + transformation.affine_multi (:101-173,:413).
 *   src/dst/pair_offsets/mask as produced by hpusm_segment_matches + hpusm_ransac_homography_batch (src = model side);
 *   H2 [npairs][9] float64 maps input -> model (NaN = no homography -> score +inf, as urban_scene.py:45-47);
 *   model_pose [18][2] float64 shared by all pairs (model_pose_shared != 0) or [npairs][njoints][2];
 *   input_pose [npairs][njoints][2] float64 (undetected joints (0,0)).
 *   The 5 background matches (thresholds.AMOUNT_BACKGROUND_FEATURES) are ordinals among each pair's inliers: taken from
 *   picks [npairs][5] if given, else from a counter-based table keyed on (seed, pair) -- the reference uses Python's
 *   unseeded random.sample (transformation.py:111).  picks_out (optional) receives the ordinals used.
 *   score [npairs] float64: max normalised distance; 1 when a pair has fewer than 5 inliers (transformation.py:107). */
int hpusm_scene_score_batch(const float* src, const float* dst, const int64_t* pair_offsets, const uint8_t* mask,
                            const double* H2, const double* model_pose, int model_pose_shared, const double* input_pose,
                            int64_t npairs, int njoints, double model_w, double model_h, uint64_t seed,
                            const int32_t* picks, double* score, int32_t* picks_out, void* stream);

/* Batched min/max normalisation, replaces common.feature_scaling (common.py:688-708) including its quirk that
 * each minimum runs over BOTH coordinates of the rows whose x (resp. y) is non-zero; negatives clamp to 0.
 *   pts, out [n][npts][2] float64. */
int hpusm_feature_scaling_batch(const double* pts, int64_t n, int npts, double* out, void* stream);

/* Per-rank top-k smallest scores among matching poses (retrieval answer that is all-gathered over
 * NCCL): writes k (score, index+index_offset) pairs, score = +inf padding. ws from
 * hpusm_pose_topk_workspace_bytes(n, k). */
size_t hpusm_pose_topk_workspace_bytes(int64_t n, int k);
int hpusm_pose_topk(const uint8_t* match, const float* score, int64_t n, int64_t index_offset, int k,
                    float* top_score, int64_t* top_index, int32_t* match_count, void* ws,
                    size_t ws_bytes, void* stream);
/* Merge of candidate lists (the all-gathered per-rank top-k lists): the k lexicographically smallest (score, index)
 * among `count` <= 4096 candidates; entries with index < 0 are padding.  One launch. */
int hpusm_pose_topk_merge(const float* cand_score, const int64_t* cand_index, int64_t count, int k, float* top_score,
                          int64_t* top_index, void* stream);

/* --------------------------------------------------------------------------------------------------
 * ORB description (the step in front of K2): cv2.ORB.compute for given keypoints, bit for bit.
 * Replaces the description half of detector.detectAndCompute (urbanscene/urban_scene.py:31-32) for the reference's ORB
 * (features.py:19-20: scaleFactor 1.2, 8 levels, WTA_K 2, patchSize 31); detection stays with OpenCV.
 *   img [height][pitch] uint8 grayscale (device).  The smoothed pyramid is built in `ws`
 *   (hpusm_orb_workspace_bytes): level L = INTER_LINEAR_EXACT resize of level L-1, each level through ORB's 7x7 Gaussian.
 *   kp_lxy [n][3] int32: octave, and the keypoint centre in that level, (cvRound(pt.x * s), cvRound(pt.y * s)) with
 *   s = 1.f / (float)pow(scale_factor, octave); kp_cs [n][2] float32: cos and sin of angle * pi/180 (computed on the host,
 *   where the keypoints come from).  desc [n][32] uint8 stays on the device and feeds hpusm_knn2_hamming directly. */
size_t hpusm_orb_workspace_bytes(int width, int height, int nlevels, double scale_factor);
int hpusm_orb_compute(const uint8_t* img, int width, int height, int64_t pitch, int nlevels, double scale_factor,
                      const int32_t* kp_lxy, const float* kp_cs, int64_t n, uint8_t* desc, void* ws, size_t ws_bytes,
                      void* stream);

/* Detection + description: cv2.ORB.detectAndCompute(img, None) for the reference's configuration (FAST threshold 20,
 * ORB_FAST_SCORE, edgeThreshold 31, patchSize 31; features.py:19-20), restated in oracle/orb_oracle.py and pinned on cv2's own
 * keypoint lists: same keypoints in the same order (level by level, row-major inside a level) with identical pt, octave,
 * size, response and angle, and their descriptors.  Everything stays on the device and is asynchronous: the number of
 * keypoints is only known there.
 *   capacity      rows available in every output; keypoints beyond it are dropped (compare level_counts[nlevels] with it).
 *   kp_xy [cap][2] float32 KeyPoint.pt; kp_octave [cap] int32; kp_angle, kp_response, kp_size [cap] float32;
 *   kp_lxy [cap][3] int32 and kp_cs [cap][2] float32: the keypoints in the form hpusm_orb_compute takes;
 *   desc [cap][32] uint8 (may be NULL: detection only);
 *   level_counts [nlevels + 1] int32: keypoints per level, then their total.  OpenCV caps every level at a quota
 *   (21717 on level 0 for nfeatures = 100000) by response; that cull is NOT done here: a caller that sees a count above
 *   the quota must fall back (the drop-in does) -- it does not happen on images below ~2 Mpixel of dense texture. */
size_t hpusm_orb_detect_workspace_bytes(int width, int height, int nlevels, double scale_factor);
int hpusm_orb_detect_and_compute(const uint8_t* img, int width, int height, int64_t pitch, int nlevels, double scale_factor,
                                 int fast_threshold, int edge_threshold, int patch_size, int64_t capacity, float* kp_xy,
                                 int32_t* kp_octave, float* kp_angle, float* kp_response, float* kp_size, int32_t* kp_lxy,
                                 float* kp_cs, uint8_t* desc, int32_t* level_counts, void* ws, size_t ws_bytes, void* stream);

/* Integer-pipe micro-benchmarks: every thread issues `iters` POPC.32 (resp. LOP3) in 8 independent chains; returns
 * nothing useful in `sink`.  bench.py uses them as the measured denominators of the Hamming roofline. */
int hpusm_microbench_popc(int64_t blocks, int threads, int64_t iters, uint32_t* sink, void* stream);
/* Same for LOP3 on the ALU pipe (the Hamming kernel issues 16 LOP3 + 4 POPC per 256-bit pair). */
int hpusm_microbench_lop3(int64_t blocks, int threads, int64_t iters, uint32_t* sink, void* stream);
/* FP32 pipe: every thread issues `iters` fused multiply-adds in 8 independent chains; packed != 0: fma.rn.f32x2 (two FMAs
 * per instruction).  bench.py uses it as the measured denominator of the RANSAC scoring roofline. */
int hpusm_microbench_ffma(int packed, int64_t blocks, int threads, int64_t iters, uint32_t* sink, void* stream);
/* Issue rate of the max-scan instructions the fused top-2 epilogues of K1 are made of.
 * op 0: 3-input s32 max, 1: 3-input f32 max, 2: 2-input s32 max, 3: 2-input f32 max. */
int hpusm_microbench_minmax(int op, int64_t blocks, int threads, int64_t iters, uint32_t* sink, void* stream);
/* Tensor-pipe issue rate: every SM issues rounds x 64 back-to-back tcgen05.mma of 128 x 256 x (16 bf16 | 32 u8) and nothing
 * else.  kind 0 = bf16 operands (kind::f16), 1 = u8 operands (kind::i8).  The measured denominator of the K1 roofline.
 * kinds 2..5: u8 at the kNN engines' own MMA shapes -- N = 64, N = 64 with A in tensor memory, N = 128, N = 128 with A in
 * tensor memory (rounds x 64 MMAs of 128 x N x 32). */
int hpusm_microbench_mma(int kind, int rounds, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* HPUSM_H_ */
