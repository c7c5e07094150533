// K3 -- fp32 re-check of the rows K2 flagged, and K4's pack/unpack helpers.
//
// K2 scores with fp16-rounded operands.  The parity contract wants keep masks and best-match
// indices equal to the reference's fp32 cosine path (extract_and_label_faces_from_dataset.py:106 evaluated
// in fp32), so every row whose decision could be changed by the rounding -- top-2 gap <= delta, or best
// within the band of the threshold -- is recomputed here with the reference's own formula
//     inner(f, g) / (|f| * |g|)
// on the ORIGINAL fp32 embeddings: only the two leading references when the third-best score was clearly
// lower, the whole reference set otherwise (first-occurrence argmax, strict '>' in ascending order).
// One warp per flagged row; the candidate row is parked in shared memory.
#include "ffr_common.cuh"

namespace ffr {

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;

__device__ __forceinline__ float cos_fp32(const float* __restrict__ c_smem, const float* __restrict__ r, int32_t dim,
                                          float cc_sqrt, int lane, bool vec) {
    float a = 0.f, b = 0.f;
    if (vec) {
        const float4* r4 = reinterpret_cast<const float4*>(r);
        const float4* c4 = reinterpret_cast<const float4*>(c_smem);
        for (int k = lane; k < (dim >> 2); k += 32) {
            const float4 rv = __ldg(r4 + k);
            const float4 cv = c4[k];
            a = fmaf(cv.x, rv.x, a); a = fmaf(cv.y, rv.y, a); a = fmaf(cv.z, rv.z, a); a = fmaf(cv.w, rv.w, a);
            b = fmaf(rv.x, rv.x, b); b = fmaf(rv.y, rv.y, b); b = fmaf(rv.z, rv.z, b); b = fmaf(rv.w, rv.w, b);
        }
    } else {
        for (int k = lane; k < dim; k += 32) {
            const float rv = __ldg(r + k), cv = c_smem[k];
            a = fmaf(cv, rv, a);
            b = fmaf(rv, rv, b);
        }
    }
    a = warp_sum(a);
    b = warp_sum(b);
    return __fdiv_rn(a, __fmul_rn(__fsqrt_rn(b), cc_sqrt));
}

__global__ void __launch_bounds__(kThreads)
recheck_kernel(const float* __restrict__ ref, int64_t n_ref, const float* __restrict__ cand, int32_t dim, float thr,
               int64_t ref_index_base, uint8_t* __restrict__ keep, int32_t* __restrict__ best_idx,
               float* __restrict__ best_val, const WsHeader* __restrict__ hdr, const RecheckRec* __restrict__ recs,
               int64_t rec_cap, float band_tol, int32_t* band_count, int64_t* band_rows, int64_t band_cap, int vec) {
    extern __shared__ __align__(16) float s_rows[];            // kWarps x dim
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    float* c_smem = s_rows + static_cast<size_t>(w) * dim;
    int64_t count = hdr->recheck_count;
    if (count > rec_cap) count = rec_cap;
    const int64_t warp = static_cast<int64_t>(blockIdx.x) * kWarps + w;
    const int64_t nwarps = static_cast<int64_t>(gridDim.x) * kWarps;
    for (int64_t k = warp; k < count; k += nwarps) {
        const RecheckRec rec = recs[k];
        const float* c = cand + static_cast<int64_t>(rec.row) * dim;
        float cc = 0.f;
        __syncwarp();
        for (int d = lane; d < dim; d += 32) { const float t = __ldg(c + d); c_smem[d] = t; cc = fmaf(t, t, cc); }
        __syncwarp();
        const float cc_sqrt = __fsqrt_rn(warp_sum(cc));
        float best;
        int32_t bi;
        if (rec.full) {
            best = -INFINITY;
            bi = 0;
            for (int64_t i = 0; i < n_ref; ++i) {
                const float s = cos_fp32(c_smem, ref + i * dim, dim, cc_sqrt, lane, vec != 0);
                if (s > best || i == 0) { best = s; bi = static_cast<int32_t>(i); }
            }
        } else {
            best = cos_fp32(c_smem, ref + static_cast<int64_t>(rec.idx1) * dim, dim, cc_sqrt, lane, vec != 0);
            bi = rec.idx1;
            if (rec.idx2 >= 0) {
                const float s2 = cos_fp32(c_smem, ref + static_cast<int64_t>(rec.idx2) * dim, dim, cc_sqrt, lane, vec != 0);
                if (s2 > best || (s2 == best && rec.idx2 < bi)) { best = s2; bi = rec.idx2; }
            }
        }
        if (lane == 0) {
            keep[rec.row] = (best >= thr) ? 1 : 0;
            best_idx[rec.row] = static_cast<int32_t>(bi + ref_index_base);
            if (best_val != nullptr) best_val[rec.row] = best;
            if (band_count != nullptr && fabsf(best - thr) <= band_tol) {
                const int32_t slot = atomicAdd(band_count, 1);
                if (band_rows != nullptr && slot < band_cap) band_rows[slot] = rec.row;
            }
        }
    }
}

// packed per-rank record block for the allgather: [idx i32 x m_pad][keep u8 x m_pad], m_pad = m rounded up to 16
__global__ void pack_results_kernel(const uint8_t* __restrict__ keep, const int32_t* __restrict__ idx, int64_t m,
                                    int64_t m_pad, uint8_t* __restrict__ packed) {
    int32_t* pidx = reinterpret_cast<int32_t*>(packed);
    uint8_t* pkeep = packed + m_pad * 4;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < m;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        pidx[i] = idx[i];
        pkeep[i] = keep[i];
    }
}

__global__ void unpack_results_kernel(const uint8_t* __restrict__ packed, int64_t m, int64_t m_pad, int nranks,
                                      uint8_t* __restrict__ keep, int32_t* __restrict__ idx) {
    const int64_t total = m * nranks;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int64_t r = i / m, j = i - r * m;
        const uint8_t* blk = packed + r * m_pad * 5;
        idx[i] = reinterpret_cast<const int32_t*>(blk)[j];
        keep[i] = blk[m_pad * 4 + j];
    }
}

}  // namespace

int launch_recheck(const float* ref, int64_t n_ref, const float* cand, int64_t n_cand, int32_t dim,
                   const float* /*ref_norm*/, const float* /*cand_norm*/, float thr, int64_t ref_index_base,
                   uint8_t* keep, int32_t* idx, float* val, const WsHeader* hdr, const RecheckRec* recs, int64_t rec_cap,
                   float band_tol, int32_t* band_count, int64_t* band_rows, int64_t band_cap, cudaStream_t s) {
    if (n_cand == 0) return FFR_OK;
    const int sms = num_sms();
    const size_t smem = static_cast<size_t>(kWarps) * dim * sizeof(float);
    if (smem > 48 * 1024) { set_error("recheck: dim %d too large", dim); return FFR_ERR_UNSUPPORTED; }
    const int vec = (dim % 4 == 0) && ((reinterpret_cast<uintptr_t>(ref) & 15) == 0);
    // the flagged-row count lives on the device; a fixed grid strides over it
    int64_t grid = (n_cand + kWarps - 1) / kWarps;
    if (grid > static_cast<int64_t>(sms) * 4) grid = static_cast<int64_t>(sms) * 4;
    recheck_kernel<<<static_cast<unsigned>(grid), kThreads, smem, s>>>(ref, n_ref, cand, dim, thr, ref_index_base, keep,
                                                                       idx, val, hdr, recs, rec_cap, band_tol,
                                                                       band_count, band_rows, band_cap, vec);
    FFR_LAUNCH_CHECK("recheck");
    return FFR_OK;
}

int launch_pack_results(const uint8_t* keep, const int32_t* idx, int64_t m, int64_t m_pad, uint8_t* packed,
                        cudaStream_t s) {
    if (m == 0) return FFR_OK;
    int64_t grid = (m + 255) / 256;
    if (grid > num_sms() * 8) grid = num_sms() * 8;
    pack_results_kernel<<<static_cast<unsigned>(grid), 256, 0, s>>>(keep, idx, m, m_pad, packed);
    FFR_LAUNCH_CHECK("pack_results");
    return FFR_OK;
}

int launch_unpack_results(const uint8_t* packed, int64_t m, int64_t m_pad, int nranks, uint8_t* keep, int32_t* idx,
                          cudaStream_t s) {
    if (m == 0) return FFR_OK;
    int64_t grid = (m * nranks + 255) / 256;
    if (grid > num_sms() * 8) grid = num_sms() * 8;
    unpack_results_kernel<<<static_cast<unsigned>(grid), 256, 0, s>>>(packed, m, m_pad, nranks, keep, idx);
    FFR_LAUNCH_CHECK("unpack_results");
    return FFR_OK;
}

}  // namespace ffr
