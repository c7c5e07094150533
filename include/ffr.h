/*
 * ffr.h -- C ABI of the B200-native similar-face-filtering hot path ("ffr" = face filter by reference).
 *
 * The reference (SamSamhuns/face_detection_and_recognition) is pure Python and has NO FFI / plugin
 * interface for this path: the arithmetic is inline NumPy in
 *     similar_face_filtering/filter_faces_using_reference.py:85-99   (mean vector + max-dist threshold)
 *     similar_face_filtering/filter_faces_using_reference.py:186-189 (per-row Euclid keep test)
 *     face_detection_and_extraction/face_extraction/extract_and_label_faces_from_dataset.py:101-116
 *                                                                    (per-pair cosine/Euclid + threshold scan)
 *     face_detection_and_extraction/modules/mobile_facenet/mobile_facenet.py:30-33  (l2_norm)
 *     face_detection_and_extraction/face_extraction/extract_and_clean_imdb_wiki_faces.py:146 (NumPy L2-normalise)
 * so the entry points below are what a ctypes binding added to those call sites would bind
 * (INTEGRATION.md shows the stub).  Every entry point cites the reference expression it replaces.
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / CUDA types in the signatures (ffr_stream_t is a
 *     cudaStream_t passed as void*; NULL = the legacy default stream).
 *   - "device" pointers are caller-owned CUDA device memory, row-major, contiguous, 16-byte aligned.
 *   - every device-pointer call is asynchronous on the given stream and allocates nothing: scratch
 *     memory is an explicit workspace sized by the matching *_workspace_bytes query.
 *   - return 0 on success, a negative FFR_ERR_* code otherwise; ffr_last_error() gives the
 *     thread-local message.  There is no CPU fallback: without a CUDA device every compute entry
 *     point fails with FFR_ERR_CUDA.
 */
#ifndef FFR_H_
#define FFR_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FFR_ABI_VERSION 1

#define FFR_OK               0
#define FFR_ERR_INVALID     -1   /* bad argument (null pointer, non-positive size, misalignment) */
#define FFR_ERR_CUDA        -2   /* CUDA runtime / driver error, or no device */
#define FFR_ERR_WORKSPACE   -3   /* workspace missing or too small */
#define FFR_ERR_UNSUPPORTED -4   /* shape / dtype / metric combination not implemented */
#define FFR_ERR_NCCL        -5   /* NCCL could not be loaded or returned an error */

#define FFR_METRIC_COSINE 0      /* best = max_i (r_i.c)/(|r_i||c|), keep = best >= thr  (extract_and_label...:106) */
#define FFR_METRIC_EUCLID 1      /* best = min_i |c - r_i|,          keep = best <= thr  (filter_faces...:189, extract_and_label...:104) */

#define FFR_DTYPE_F32 0          /* raw fp32 embeddings; the op normalises internally */
#define FFR_DTYPE_F16 1          /* rows already L2-normalised + converted by ffr_l2norm_rows_f32 (leading dim = ffr_padded_dim).
                                    There are no fp32 rows to re-check against: keep / best_idx are then exact in fp16-OPERAND space
                                    (the index carries the row's largest fp16-operand score, ties -> first), i.e. they can differ from
                                    the fp32 decision for pairs closer than ~1.5e-4; best_val is within 1e-3 as always */

/* flags for ffr_filter_ex */
#define FFR_FLAG_FORCE_FP32   1  /* use the exact fp32 CUDA-core kernel whatever the shape */
#define FFR_FLAG_FORCE_MMA    2  /* use the tcgen05 kernel (cosine, dim <= 512) whatever n_ref */
#define FFR_FLAG_NO_RECHECK   4  /* skip the fp32 re-check of near-tie / near-threshold rows (benchmarking only): decisions as for FFR_DTYPE_F16 */

typedef void* ffr_stream_t;      /* cudaStream_t */
typedef struct ffr_ctx  ffr_ctx; /* host-buffer pipeline context (streams, staging, workspace) */
typedef struct ffr_comm ffr_comm;/* one NCCL communicator + scratch */

int         ffr_abi_version(void);
const char* ffr_last_error(void);
const char* ffr_build_info(void);            /* "sm_100a; nvcc x.y; ..." */
int         ffr_device_count(void);           /* 0 when no CUDA device is visible */

/* leading dimension (in elements) of the fp16 rows the tensor-core kernel consumes: dim rounded up to 64 */
int32_t ffr_padded_dim(int32_t dim);

/* ---- K1: row L2-normalisation ---------------------------------------------------------------
 * y[i,:] = x[i,:] / |x[i,:]|_2   (no epsilon)      replaces  mobile_facenet.py:30-33  l2_norm
 *                                                            extract_and_clean_imdb_wiki_faces.py:146
 * and supplies the denominators of extract_and_label_faces_from_dataset.py:106.
 * Any of the three outputs may be NULL.  y_f16 has leading dimension y_f16_ld >= dim (columns
 * dim..y_f16_ld-1 are written as zeros); y_f32 and x may alias.  norms[i] = |x[i,:]|_2. */
int ffr_l2norm_rows_f32(const float* x, int64_t rows, int32_t dim,
                        void* y_f16, int32_t y_f16_ld, float* y_f32, float* norms,
                        ffr_stream_t stream);

/* ---- K2/K2s/K3: fused reference x candidate filter -------------------------------------------
 * For every candidate row c of cand[n_cand, dim]:
 *   cosine:  best = max_i (r_i.c)/(|r_i||c|)   idx = first argmax   keep = best >= thr
 *   euclid:  best = min_i |c - r_i|_2          idx = first argmin   keep = best <= thr
 * replaces the per-row test  filter_faces_using_reference.py:186-189 (euclid, n_ref = 1, ref = mean
 * vector) and the per-pair scan extract_and_label_faces_from_dataset.py:101-116; the n_ref x n_cand
 * similarity matrix is never materialised.
 *   ref, cand      device, dtype FFR_DTYPE_F32 (dim floats per row) or FFR_DTYPE_F16 (see above)
 *   ref_norm,cand_norm  device, only read with FFR_DTYPE_F16 (may be NULL for cosine)
 *   ref_index_base added to every best_idx (global index of ref row 0)
 *   keep u8[n_cand], best_idx i32[n_cand], best_val f32[n_cand]   device outputs (best_val may be NULL)
 * ffr_filter_ex additionally lists the tolerance band: rows with |best - thr| <= band_tol are appended
 * (unordered) to band_rows[0..band_cap) and counted in *band_count (device int32; count may exceed cap); `best` is the
 * fp32 re-checked score, or the fp16-operand score where no re-check runs (FFR_DTYPE_F16, FFR_FLAG_NO_RECHECK).
 * Bit-identical reference rows are folded internally (a later copy can never be the first arg-best); best_idx always
 * refers to the caller's row numbering. */
size_t ffr_filter_workspace_bytes(int64_t n_ref, int64_t n_cand, int32_t dim, int dtype, int metric);

int ffr_filter(const void* ref, int64_t n_ref, const void* cand, int64_t n_cand, int32_t dim, int dtype,
               const float* ref_norm, const float* cand_norm, int metric, float thr, int64_t ref_index_base,
               uint8_t* keep, int32_t* best_idx, float* best_val,
               void* workspace, size_t ws_bytes, ffr_stream_t stream);

int ffr_filter_ex(const void* ref, int64_t n_ref, const void* cand, int64_t n_cand, int32_t dim, int dtype,
                  const float* ref_norm, const float* cand_norm, int metric, float thr, int64_t ref_index_base,
                  uint8_t* keep, int32_t* best_idx, float* best_val,
                  float band_tol, int32_t* band_count, int64_t* band_rows, int64_t band_cap,
                  int flags, void* workspace, size_t ws_bytes, ffr_stream_t stream);

/* statistics of the last ffr_filter* call that used this workspace (device->host copy, synchronises
 * the stream): out[0] = rows re-checked in fp32 against their two or three leading references (near-tie or
 * near-threshold), out[1] = rows that needed the full fp32 rescan, out[2] = path taken (0 fp32 CUDA-core,
 * 1 tcgen05), out[3] = kernels launched, out[4] = rows re-checked against one 128-reference part, out[5] = reference
 * rows the tensor-core kernel scanned (< n_ref when bit-identical reference rows were folded), out[6..7] = 0. */
int ffr_filter_stats(const void* workspace, int64_t out[8], ffr_stream_t stream);

/* ---- K5: reference statistics -----------------------------------------------------------------
 * mean[d] = mean_i ref_feat[i,d];  *thres = max_i |mean - ref_feat[i,:]|_2
 * replaces filter_faces_using_reference.py:85-99.  n_ref <= 4096.  mean f32[dim], thres f32[1]: device. */
int ffr_ref_mean_and_thres(const float* ref_feat, int32_t n_ref, int32_t dim,
                           float* mean, float* thres, ffr_stream_t stream);

/* all classes in one launch: class c owns rows offsets[c] .. offsets[c+1]-1 of ref_feat (offsets: device int32
 * [n_classes + 1]); mean f32[n_classes, dim], thres f32[n_classes].  Same arithmetic per class as above. */
int ffr_ref_mean_and_thres_batched(const float* ref_feat, const int32_t* offsets, int32_t n_classes, int32_t dim,
                                   float* mean, float* thres, ffr_stream_t stream);

/* ---- streaming first-match gallery scan ("next" row: the reference's face tracker) -------------------
 * replaces Net.check_if_face_exists + add_face, face_extraction/extract_and_label_faces_from_dataset.py:101-121.
 * Queries are processed IN ORDER against a device-resident gallery (feat f32[capacity, dim], bbox f32[capacity, 4],
 * *gallery_count entries in use): the first entry i (ascending) with
 *     (dist < normal_thres and iou(bbox_i, query_bbox) > 0.1) or dist < harsh_thres
 * is overwritten by the query and match_idx[q] = i; otherwise the query is appended at position p = old count and
 * match_idx[q] = -1 - p (INT32_MIN if the gallery is full).  dist = 1 - cos (FFR_METRIC_COSINE, :106) or the Euclid
 * distance (FFR_METRIC_EUCLID, :104).  query_bbox may be NULL (iou = 0: only the harsh threshold can match). */
int ffr_first_match_stream(float* gallery_feat, float* gallery_bbox, int32_t* gallery_count, int32_t capacity,
                           const float* queries, const float* query_bbox, int32_t n_queries, int32_t dim, int metric,
                           float normal_thres, float harsh_thres, int32_t* match_idx, ffr_stream_t stream);

/* ---- host-buffer ("plugin") entry points -----------------------------------------------------
 * Same semantics as ffr_filter with HOST pointers: candidates are streamed host->device in chunks on
 * one stream while the previous chunk is filtered on another; results are copied back.  Pinned host
 * buffers overlap fully; pageable ones work but serialise.  Synchronous (returns when results are in
 * the host arrays). */
int  ffr_ctx_create(int device, int64_t max_ref, int64_t chunk_cand, int32_t max_dim, ffr_ctx** out);
void ffr_ctx_destroy(ffr_ctx* ctx);
int  ffr_ctx_filter_host(ffr_ctx* ctx, const float* ref, int64_t n_ref, const float* cand, int64_t n_cand,
                         int32_t dim, int metric, float thr, int64_t ref_index_base,
                         uint8_t* keep, int32_t* best_idx, float* best_val, int flags);
/* number of this library's kernel launches issued through ctx so far (for bench accounting) */
int64_t ffr_ctx_launch_count(const ffr_ctx* ctx);
/* process-wide count of this library's kernel launches */
int64_t ffr_launch_count(void);

/* ---- K4: multi-GPU gather of the per-candidate result (one rank per GPU) -----------------------
 * Candidates are sharded contiguously, m_local rows per rank (equal on every rank; pad the last).
 * keep_all / idx_all receive nranks * m_local entries in rank order.  One ncclAllGather of the
 * packed {idx i32, keep u8} records over NVLink; NCCL is dlopen'ed (libnccl.so.2) on first use. */
int  ffr_nccl_unique_id(void* id128);                        /* 128 bytes, call on rank 0 and broadcast */
int  ffr_comm_create(const void* id128, int nranks, int rank, int device, ffr_comm** out);
void ffr_comm_destroy(ffr_comm* comm);
size_t ffr_allgather_workspace_bytes(int nranks, int64_t m_local);
int  ffr_allgather_results(ffr_comm* comm, const uint8_t* keep_local, const int32_t* idx_local, int64_t m_local,
                           uint8_t* keep_all, int32_t* idx_all, void* workspace, size_t ws_bytes,
                           ffr_stream_t stream);
/* In-place form (no pack / unpack kernels, no workspace): this rank's ffr_filter already wrote its m_local results at
 * keep_all + rank * m_local and idx_all + rank * m_local (pass those as the filter's keep / best_idx outputs); the two
 * in-place ncclAllGather calls are grouped into ONE NCCL launch that fills in everybody else's slices. */
int  ffr_allgather_results_inplace(ffr_comm* comm, uint8_t* keep_all, int32_t* idx_all, int64_t m_local,
                                   ffr_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* FFR_H_ */
