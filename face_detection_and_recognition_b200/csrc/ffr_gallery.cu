// "Next" rows of the hot-path scope: batched reference statistics and the streaming first-match
// gallery scan of the reference's face tracker.
//
//   ref_stats_batched_kernel   mean vector + max distance per class, all classes in one launch
//                              (similar_face_filtering/filter_faces_using_reference.py:85-99, once per class at :161-164)
//   first_match_stream_kernel  Net.check_if_face_exists + add_face
//                              (face_detection_and_extraction/face_extraction/extract_and_label_faces_from_dataset.py:101-121):
//                              for every query, in order: scan the gallery in ascending order, the FIRST entry with
//                              (dist < normal_thres and iou > 0.1) or dist < harsh_thres wins and is overwritten by the
//                              query (feature + bbox); otherwise the query is appended.  Queries depend on each other
//                              through the gallery, so one CTA walks them sequentially; each query's scan is parallel
//                              (one warp per gallery entry, 32 entries per round, early exit after the first hit).
#include "ffr_common.cuh"

namespace ffr {

namespace {

__global__ void __launch_bounds__(256)
ref_stats_batched_kernel(const float* __restrict__ x, const int32_t* __restrict__ offsets, int32_t dim,
                         float* __restrict__ mean, float* __restrict__ thres, int mode) {
    extern __shared__ __align__(16) float s_mean[];   // dim floats + 8 warp maxima
    float* s_max = s_mean + dim;
    const int c = blockIdx.x;
    const int32_t lo = offsets[c], hi = offsets[c + 1];
    const int32_t n = hi - lo;
    for (int d = threadIdx.x; d < dim; d += blockDim.x) {
        float acc = 0.f;
        for (int i = lo; i < hi; ++i) acc += x[static_cast<int64_t>(i) * dim + d];     // np.mean(axis=0): row order
        const float m = __fdiv_rn(acc, static_cast<float>(n));
        s_mean[d] = m;
        mean[static_cast<int64_t>(c) * dim + d] = m;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    float wmax = 0.f;
    for (int i = lo + warp; i < hi; i += nw)          // the distance K2s will compute for this row as a candidate, bit for bit
        wmax = fmaxf(wmax, k2s_euclid_dist(x + static_cast<int64_t>(i) * dim, s_mean, dim, lane, mode));
    if (lane == 0) s_max[warp] = wmax;
    __syncthreads();
    if (threadIdx.x == 0) {
        float m = 0.f;
        for (int w = 0; w < nw; ++w) m = fmaxf(m, s_max[w]);
        thres[c] = m;
    }
}

// modules/utils/image.py:124-143
__device__ __forceinline__ float bbox_iou(const float* a, const float* b) {
    const float xd = fminf(a[2], b[2]) - fmaxf(a[0], b[0]);
    const float yd = fminf(a[3], b[3]) - fmaxf(a[1], b[1]);
    if (xd < 0.f || yd < 0.f) return 0.f;
    const float inter = xd * yd;
    return inter / (((a[2] - a[0]) * (a[3] - a[1])) + ((b[2] - b[0]) * (b[3] - b[1])) - inter);
}

constexpr int kScanThreads = 1024;

__global__ void __launch_bounds__(kScanThreads)
first_match_stream_kernel(float* __restrict__ g_feat, float* __restrict__ g_bbox, int32_t* __restrict__ g_count, int32_t cap,
                          const float* __restrict__ queries, const float* __restrict__ qboxes, int32_t n_queries,
                          int32_t dim, int metric, float normal_thres, float harsh_thres, int32_t* __restrict__ match_idx) {
    extern __shared__ float s_q[];                    // the current query
    __shared__ int s_first;
    __shared__ float s_qnorm;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = kScanThreads / 32;
    int count = *g_count;
    for (int q = 0; q < n_queries; ++q) {
        const float* qf = queries + static_cast<int64_t>(q) * dim;
        __syncthreads();
        for (int d = threadIdx.x; d < dim; d += kScanThreads) s_q[d] = qf[d];
        if (threadIdx.x == 0) s_first = 0x7FFFFFFF;
        __syncthreads();
        if (warp == 0) {
            float a = 0.f;
            for (int d = lane; d < dim; d += 32) a = fmaf(s_q[d], s_q[d], a);
            a = warp_sum(a);
            if (lane == 0) s_qnorm = __fsqrt_rn(a);
        }
        __syncthreads();
        for (int base = 0; base < count; base += nw) {        // ascending rounds of one entry per warp
            const int i = base + warp;
            if (i < count) {
                const float* gf = g_feat + static_cast<int64_t>(i) * dim;
                float a = 0.f, b = 0.f;
                if (metric == FFR_METRIC_EUCLID) {
                    for (int d = lane; d < dim; d += 32) { const float t = gf[d] - s_q[d]; a = fmaf(t, t, a); }
                    a = warp_sum(a);
                } else {
                    for (int d = lane; d < dim; d += 32) { const float gv = gf[d]; a = fmaf(gv, s_q[d], a); b = fmaf(gv, gv, b); }
                    a = warp_sum(a);
                    b = warp_sum(b);
                }
                if (lane == 0) {
                    const float dist = metric == FFR_METRIC_EUCLID
                                           ? __fsqrt_rn(a)                                                  // :104
                                           : 1.f - __fdiv_rn(a, __fmul_rn(__fsqrt_rn(b), s_qnorm));         // :106
                    const float iou = qboxes != nullptr ? bbox_iou(g_bbox + static_cast<int64_t>(i) * 4,
                                                                  qboxes + static_cast<int64_t>(q) * 4) : 0.f;
                    if ((dist < normal_thres && iou > 0.1f) || dist < harsh_thres) atomicMin(&s_first, i);   // :110
                }
            }
            __syncthreads();
            if (s_first != 0x7FFFFFFF) break;                  // every entry below base + nw has been examined
            __syncthreads();
        }
        __syncthreads();
        const int first = s_first;
        int dst;
        if (first != 0x7FFFFFFF) {                             // :113-114 overwrite the matched entry
            dst = first;
            if (threadIdx.x == 0) match_idx[q] = first;
        } else if (count < cap) {                              // add_face :118-121
            dst = count;
            if (threadIdx.x == 0) match_idx[q] = -1 - count;
            ++count;
        } else {
            dst = -1;
            if (threadIdx.x == 0) match_idx[q] = INT32_MIN;    // gallery full: query dropped
        }
        if (dst >= 0) {
            for (int d = threadIdx.x; d < dim; d += kScanThreads) g_feat[static_cast<int64_t>(dst) * dim + d] = s_q[d];
            if (threadIdx.x < 4 && qboxes != nullptr)
                g_bbox[static_cast<int64_t>(dst) * 4 + threadIdx.x] = qboxes[static_cast<int64_t>(q) * 4 + threadIdx.x];
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) *g_count = count;
}

}  // namespace

int launch_ref_stats_batched(const float* ref_feat, const int32_t* offsets, int32_t n_classes, int32_t dim, float* mean,
                             float* thres, cudaStream_t s) {
    if (n_classes == 0) return FFR_OK;
    const size_t smem = (static_cast<size_t>(dim) + 8) * sizeof(float);
    ref_stats_batched_kernel<<<static_cast<unsigned>(n_classes), 256, smem, s>>>(ref_feat, offsets, dim, mean, thres,
                                                                                    k2s_dist_mode(ref_feat, dim));
    FFR_LAUNCH_CHECK("ref_stats_batched");
    return FFR_OK;
}

int launch_first_match_stream(float* g_feat, float* g_bbox, int32_t* g_count, int32_t cap, const float* queries,
                              const float* qboxes, int32_t n_queries, int32_t dim, int metric, float normal_thres,
                              float harsh_thres, int32_t* match_idx, cudaStream_t s) {
    if (n_queries == 0) return FFR_OK;
    const size_t smem = static_cast<size_t>(dim) * sizeof(float);
    first_match_stream_kernel<<<1, kScanThreads, smem, s>>>(g_feat, g_bbox, g_count, cap, queries, qboxes, n_queries, dim,
                                                            metric, normal_thres, harsh_thres, match_idx);
    FFR_LAUNCH_CHECK("first_match_stream");
    return FFR_OK;
}

}  // namespace ffr
