// K1 -- row L2-normalisation of an embedding matrix, sm_100a.
//
// Replaces  face_detection_and_extraction/modules/mobile_facenet/mobile_facenet.py:30-33 (l2_norm:
// norm = torch.norm(x, 2, axis, True); x / norm) and
// face_detection_and_extraction/face_extraction/extract_and_clean_imdb_wiki_faces.py:146
// (face_feat / np.linalg.norm(face_feat)).  No epsilon, true IEEE division, like the reference.
//
// HBM-bound: reads 4*dim bytes per row, writes 2*ld16 (fp16 copy for the tensor-core filter) and/or
// 4*dim (fp32 copy) and 4 (norm).  One warp owns one row at a time: lane l loads float4 #(l + 32 j) of
// the row (fully coalesced 512 B per warp instruction, streaming / no L1 allocation), the sum of squares
// is reduced with warp shuffles, and each lane normalises and stores what it loaded.  Two rows are in
// flight per warp iteration to keep enough loads outstanding; the grid is a multiple of the SM count and
// warps stride over rows.
#include "ffr_common.cuh"

namespace ffr {

namespace {

constexpr int kThreads = 256;           // 8 warps per CTA
constexpr int kRowsPerIter = 2;

template <int NV>                        // float4 per lane per row: dim == 128 * NV
__global__ void __launch_bounds__(kThreads)
l2norm_rows_vec_kernel(const float* __restrict__ x, int64_t rows, int32_t dim,
                       __half* __restrict__ y16, int32_t ld16, float* __restrict__ y32,
                       float* __restrict__ norms) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (static_cast<int64_t>(blockIdx.x) * kThreads + threadIdx.x) >> 5;
    const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * kThreads) >> 5;

    for (int64_t r0 = warp * kRowsPerIter; r0 < rows; r0 += nwarps * kRowsPerIter) {
        float4 v[kRowsPerIter][NV];
#pragma unroll
        for (int i = 0; i < kRowsPerIter; ++i) {
            const int64_t r = r0 + i;
            if (r < rows) {
                const float4* p = reinterpret_cast<const float4*>(x + r * dim);
#pragma unroll
                for (int j = 0; j < NV; ++j) v[i][j] = ldg_stream_f4(p + lane + 32 * j);
            }
        }
#pragma unroll
        for (int i = 0; i < kRowsPerIter; ++i) {
            const int64_t r = r0 + i;
            if (r >= rows) break;
            float ss = 0.f;
#pragma unroll
            for (int j = 0; j < NV; ++j) {
                ss = fmaf(v[i][j].x, v[i][j].x, ss);
                ss = fmaf(v[i][j].y, v[i][j].y, ss);
                ss = fmaf(v[i][j].z, v[i][j].z, ss);
                ss = fmaf(v[i][j].w, v[i][j].w, ss);
            }
            ss = warp_sum(ss);
            const float nrm = sqrtf(ss);
            if (norms != nullptr && lane == 0) norms[r] = nrm;
#pragma unroll
            for (int j = 0; j < NV; ++j) {
                float4 o;
                o.x = __fdiv_rn(v[i][j].x, nrm);
                o.y = __fdiv_rn(v[i][j].y, nrm);
                o.z = __fdiv_rn(v[i][j].z, nrm);
                o.w = __fdiv_rn(v[i][j].w, nrm);
                if (y32 != nullptr) reinterpret_cast<float4*>(y32 + r * dim)[lane + 32 * j] = o;
                if (y16 != nullptr) {
                    __half2 h0 = __floats2half2_rn(o.x, o.y);
                    __half2 h1 = __floats2half2_rn(o.z, o.w);
                    uint2 pk;
                    pk.x = *reinterpret_cast<uint32_t*>(&h0);
                    pk.y = *reinterpret_cast<uint32_t*>(&h1);
                    reinterpret_cast<uint2*>(y16 + r * ld16)[lane + 32 * j] = pk;
                }
            }
            if (y16 != nullptr) {                                   // zero the K padding (dim..ld16)
                for (int c = dim + lane * 4; c < ld16; c += 128)
                    *reinterpret_cast<uint2*>(y16 + r * ld16 + c) = make_uint2(0u, 0u);
            }
        }
    }
}

// any dim: one warp per row, scalar accesses (row may be unaligned for float4)
__global__ void __launch_bounds__(kThreads)
l2norm_rows_generic_kernel(const float* __restrict__ x, int64_t rows, int32_t dim,
                           __half* __restrict__ y16, int32_t ld16, float* __restrict__ y32,
                           float* __restrict__ norms) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (static_cast<int64_t>(blockIdx.x) * kThreads + threadIdx.x) >> 5;
    const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * kThreads) >> 5;
    for (int64_t r = warp; r < rows; r += nwarps) {
        const float* p = x + r * dim;
        float ss = 0.f;
        for (int c = lane; c < dim; c += 32) { const float t = __ldg(p + c); ss = fmaf(t, t, ss); }
        ss = warp_sum(ss);
        const float nrm = sqrtf(ss);
        if (norms != nullptr && lane == 0) norms[r] = nrm;
        for (int c = lane; c < ld16 || c < dim; c += 32) {
            const float o = c < dim ? __fdiv_rn(__ldg(p + c), nrm) : 0.f;
            if (y32 != nullptr && c < dim) y32[r * dim + c] = o;
            if (y16 != nullptr && c < ld16) y16[r * ld16 + c] = __float2half_rn(o);
        }
    }
}

}  // namespace

int launch_l2norm(const float* x, int64_t rows, int32_t dim, __half* y16, int32_t ld16, float* y32, float* norms,
                  cudaStream_t s) {
    if (rows == 0) return FFR_OK;
    const int sms = num_sms();
    const int64_t warps_needed = (rows + kRowsPerIter - 1) / kRowsPerIter;
    const int64_t blocks_needed = (warps_needed + (kThreads / 32) - 1) / (kThreads / 32);
    // up to 8 resident CTAs per SM (2048 threads); whole multiples of the SM count when the matrix is big
    int64_t grid = blocks_needed < static_cast<int64_t>(sms) * 8 ? blocks_needed : static_cast<int64_t>(sms) * 8;
    if (grid < 1) grid = 1;
    const bool vec_ok = (dim % 128 == 0) && (y16 == nullptr || (ld16 % 4 == 0)) &&
                        ((reinterpret_cast<uintptr_t>(x) & 15) == 0) &&
                        (y32 == nullptr || (reinterpret_cast<uintptr_t>(y32) & 15) == 0) &&
                        (y16 == nullptr || (reinterpret_cast<uintptr_t>(y16) & 7) == 0);
    const dim3 g(static_cast<unsigned>(grid)), b(kThreads);
    if (vec_ok && dim == 128)       l2norm_rows_vec_kernel<1><<<g, b, 0, s>>>(x, rows, dim, y16, ld16, y32, norms);
    else if (vec_ok && dim == 256)  l2norm_rows_vec_kernel<2><<<g, b, 0, s>>>(x, rows, dim, y16, ld16, y32, norms);
    else if (vec_ok && dim == 384)  l2norm_rows_vec_kernel<3><<<g, b, 0, s>>>(x, rows, dim, y16, ld16, y32, norms);
    else if (vec_ok && dim == 512)  l2norm_rows_vec_kernel<4><<<g, b, 0, s>>>(x, rows, dim, y16, ld16, y32, norms);
    else if (vec_ok && dim == 1024) l2norm_rows_vec_kernel<8><<<g, b, 0, s>>>(x, rows, dim, y16, ld16, y32, norms);
    else l2norm_rows_generic_kernel<<<g, b, 0, s>>>(x, rows, dim, y16, ld16, y32, norms);
    FFR_LAUNCH_CHECK("l2norm_rows");
    return FFR_OK;
}

}  // namespace ffr
