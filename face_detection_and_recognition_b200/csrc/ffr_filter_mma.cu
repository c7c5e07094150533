// K2 -- fused reference x candidate cosine GEMM + threshold + running max/argmax, tcgen05/TMEM/TMA, sm_100a.
//
// Replaces the per-pair scan of
//     face_detection_and_extraction/face_extraction/extract_and_label_faces_from_dataset.py:101-116
// (cosine :106, threshold :110) and, through the unit-norm equivalence d <= t <=> cos >= 1 - t^2/2, the
// per-row test of similar_face_filtering/filter_faces_using_reference.py:186-189, for reference sets
// large enough to be tensor-core work.  The n_ref x n_cand similarity matrix only ever exists as
// 128 x 256 fp32 accumulator tiles in tensor memory; HBM sees the embeddings once and 9 bytes of result
// per candidate.
//
// Operands are the fp16, L2-normalised rows produced by K1 (leading dimension = dim rounded up to 64,
// zero padded), so the fp32 accumulator IS the cosine similarity.  fp16 operand rounding gives
// |score - fp32 score| ~ 2e-5 rms (<= ~1.5e-4 observed), which is why K3 re-checks in fp32 every row whose
// top-2 gap or distance to the threshold is <= delta.
//
// Decomposition (one CTA per SM, persistent over candidate tiles):
//   * candidates -> MMA M (TMEM lanes), references -> MMA N (TMEM columns).  A CTA keeps its 128-candidate
//     A tile (all of K) resident in shared memory and streams 256-reference B tiles, K-block by K-block,
//     out of L2 through a TMA/mbarrier ring: per 128 x 256 x K tile only B moves.
//   * warp 0: TMA producer.  warp 1: tcgen05.mma issuer (single thread) + TMEM owner.  warps 2-5: epilogue,
//     one thread per candidate row (TMEM lane), so the max/argmax over references is a pure in-register
//     reduction over columns -- no shuffles, no shared memory.
//   * the accumulator is double buffered in TMEM (2 x 256 of the 512 columns): the MMAs of reference tile
//     t+1 overlap the epilogue of tile t.
//   * epilogue per 32-column chunk: tcgen05.ld -> max tree (3-input max) -> only if the chunk max comes
//     within delta of the running best is the chunk rescanned (8 columns at a time) to update the running
//     top-3 {best, idx} {second, idx} third.  Ascending column order + strict '>' = np.argmax first
//     occurrence.  The top-3 is what K3 needs to make the index and keep bit exact in fp32.
#include <cuda.h>

#include "ffr_common.cuh"

namespace ffr {

namespace {

constexpr int kTileM = 128;          // candidates per CTA tile
constexpr int kTileN = 256;          // references per accumulator stage
constexpr int kBlockK = 64;          // fp16 per 128-byte swizzle row
constexpr int kUmmaK = 16;
constexpr int kThreads = 192;        // warp 0 TMA | warp 1 MMA | warps 2..5 epilogue
constexpr int kMaxAStages = 4;
constexpr int kMaxBStages = 8;
constexpr uint32_t kABlockBytes = kTileM * kBlockK * 2;      // 16 KiB: one K-block of the A tile
constexpr uint32_t kBStageBytes = kTileN * kBlockK * 2;      // 32 KiB: one K-block of a B tile
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kSmemLimit = 232448;                       // 227 KiB opt-in maximum
constexpr uint32_t kBarrierBytes = 1024;

struct Top3 {
    float b1, b2, b3;
    int32_t i1, i2;
};

__device__ __forceinline__ void top3_insert(Top3& t, float v, int32_t idx) {
    const bool g1 = v > t.b1, g2 = v > t.b2, g3 = v > t.b3;
    t.b3 = g2 ? t.b2 : (g3 ? v : t.b3);
    t.i2 = g1 ? t.i1 : (g2 ? idx : t.i2);
    t.b2 = g1 ? t.b1 : (g2 ? v : t.b2);
    t.i1 = g1 ? idx : t.i1;
    t.b1 = g1 ? v : t.b1;
}

// v: 32 consecutive scores of this thread's candidate; base = reference index of v[0]
__device__ __forceinline__ void process_chunk(const float (&v)[32], int32_t base, float delta, Top3& t) {
    float s[4];
#pragma unroll
    for (int k = 0; k < 4; ++k)
        s[k] = fmax3(fmax3(v[8 * k + 0], v[8 * k + 1], v[8 * k + 2]), fmax3(v[8 * k + 3], v[8 * k + 4], v[8 * k + 5]),
                     fmaxf(v[8 * k + 6], v[8 * k + 7]));
    const float cmax = fmax3(s[0], s[1], fmaxf(s[2], s[3]));
    if (__any_sync(0xffffffffu, cmax >= t.b1 - delta)) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (__any_sync(0xffffffffu, s[k] >= t.b1 - delta)) {
#pragma unroll
                for (int j = 0; j < 8; ++j) top3_insert(t, v[8 * k + j], base + 8 * k + j);
            }
        }
    }
}

__device__ __forceinline__ void mask_chunk(float (&v)[32], int valid) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = (j < valid) ? v[j] : -INFINITY;
}

__global__ void __launch_bounds__(kThreads, 1)
filter_mma_kernel(const __grid_constant__ CUtensorMap tmap_cand, const __grid_constant__ CUtensorMap tmap_ref,
                  int64_t n_ref, int64_t n_cand, int32_t kb_count, int32_t a_stages, int32_t b_stages,
                  float thr, float delta, float thr_band, int64_t ref_index_base,
                  uint8_t* __restrict__ keep, int32_t* __restrict__ best_idx, float* __restrict__ best_val,
                  RecheckLists lists, int no_recheck, float* __restrict__ dbg_scores) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // 1024-byte alignment is required by SWIZZLE_128B; the dynamic smem base is not guaranteed to have it.
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const uint32_t a_stage_bytes = static_cast<uint32_t>(kb_count) * kABlockBytes;
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem_a + static_cast<size_t>(a_stages) * a_stage_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_b + static_cast<size_t>(b_stages) * kBStageBytes);
    uint64_t* a_full = bars;
    uint64_t* a_empty = a_full + kMaxAStages;
    uint64_t* b_full = a_empty + kMaxAStages;
    uint64_t* b_empty = b_full + kMaxBStages;
    uint64_t* t_full = b_empty + kMaxBStages;
    uint64_t* t_empty = t_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_empty + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int64_t n_tiles = (n_cand + kTileM - 1) / kTileM;
    const int32_t n_rt = static_cast<int32_t>((n_ref + kTileN - 1) / kTileN);

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmap_cand);
        tma_prefetch_desc(&tmap_ref);
        for (int i = 0; i < kMaxAStages; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
        for (int i = 0; i < kMaxBStages; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&t_full[i], 1); mbar_init(&t_empty[i], 4); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<kTmemCols>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            uint32_t a_it = 0, b_it = 0;
            for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                const uint32_t as = a_it % a_stages, aph = (a_it / a_stages) & 1;
                mbar_wait(&a_empty[as], aph ^ 1);
                mbar_expect_tx(&a_full[as], a_stage_bytes);
                for (int kb = 0; kb < kb_count; ++kb)
                    tma_load_2d(smem_a + as * a_stage_bytes + kb * kABlockBytes, &tmap_cand, &a_full[as],
                                kb * kBlockK, static_cast<int32_t>(tile * kTileM), kEvictFirst);
                ++a_it;
                for (int rt = 0; rt < n_rt; ++rt) {
                    for (int kb = 0; kb < kb_count; ++kb) {
                        const uint32_t bs = b_it % b_stages, bph = (b_it / b_stages) & 1;
                        mbar_wait(&b_empty[bs], bph ^ 1);
                        mbar_expect_tx(&b_full[bs], kBStageBytes);
                        tma_load_2d(smem_b + bs * kBStageBytes, &tmap_ref, &b_full[bs], kb * kBlockK, rt * kTileN,
                                    kEvictLast);
                        ++b_it;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            uint32_t a_it = 0, b_it = 0, t_it = 0;
            for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                const uint32_t as = a_it % a_stages, aph = (a_it / a_stages) & 1;
                mbar_wait(&a_full[as], aph);
                tc_fence_after();
                const uint32_t a_base = smem_u32(smem_a + as * a_stage_bytes);
                for (int rt = 0; rt < n_rt; ++rt) {
                    const uint32_t acc = t_it & 1, tph = (t_it >> 1) & 1;
                    mbar_wait(&t_empty[acc], tph ^ 1);
                    tc_fence_after();
                    int64_t ncols = n_ref - static_cast<int64_t>(rt) * kTileN;
                    if (ncols > kTileN) ncols = kTileN;
                    const uint32_t n_mma = static_cast<uint32_t>((ncols + 15) & ~int64_t(15));
                    const uint32_t idesc = umma_idesc_f16(kTileM, n_mma);
                    const uint32_t d_tmem = tmem_base + acc * kTileN;
                    for (int kb = 0; kb < kb_count; ++kb) {
                        const uint32_t bs = b_it % b_stages, bph = (b_it / b_stages) & 1;
                        mbar_wait(&b_full[bs], bph);
                        tc_fence_after();
                        const uint32_t b_base = smem_u32(smem_b + bs * kBStageBytes);
#pragma unroll
                        for (int k = 0; k < kBlockK / kUmmaK; ++k) {
                            const uint64_t da = umma_desc_sw128(a_base + kb * kABlockBytes + k * (kUmmaK * 2));
                            const uint64_t db = umma_desc_sw128(b_base + k * (kUmmaK * 2));
                            umma_f16(d_tmem, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
                        }
                        umma_commit(&b_empty[bs]);            // B stage reusable once these MMAs retire
                        ++b_it;
                    }
                    umma_commit(&t_full[acc]);                // accumulator complete -> epilogue
                    ++t_it;
                }
                umma_commit(&a_empty[as]);                    // A stage reusable
                ++a_it;
            }
        }
    } else {
        // ===================== epilogue: one thread per candidate row =====================
        const int q = warp & 3;                               // TMEM lane quadrant this warp may access
        uint32_t t_it = 0;
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            Top3 t;
            t.b1 = t.b2 = t.b3 = -INFINITY;
            t.i1 = 0;
            t.i2 = -1;
            const int64_t row = tile * kTileM + q * 32 + lane;
            for (int rt = 0; rt < n_rt; ++rt) {
                const uint32_t acc = t_it & 1, tph = (t_it >> 1) & 1;
                mbar_wait(&t_full[acc], tph);
                tc_fence_after();
                int64_t ncols64 = n_ref - static_cast<int64_t>(rt) * kTileN;
                const int ncols = ncols64 > kTileN ? kTileN : static_cast<int>(ncols64);
                const int nchunks = (ncols + 31) >> 5;
                const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * kTileN;
                const int32_t col0 = rt * kTileN;
                float va[32], vb[32];
                tmem_ld_32x32(taddr, va);
                for (int c = 0; c < nchunks; c += 2) {
                    tmem_ld_wait();
                    if (c + 1 < nchunks) tmem_ld_32x32(taddr + (c + 1) * 32, vb);
                    if (ncols - c * 32 < 32) mask_chunk(va, ncols - c * 32);
                    if (dbg_scores != nullptr && row < n_cand) {
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (col0 + c * 32 + j < n_ref) dbg_scores[row * n_ref + col0 + c * 32 + j] = va[j];
                    }
                    process_chunk(va, col0 + c * 32, delta, t);
                    if (c + 1 < nchunks) {
                        tmem_ld_wait();
                        if (c + 2 < nchunks) tmem_ld_32x32(taddr + (c + 2) * 32, va);
                        if (ncols - (c + 1) * 32 < 32) mask_chunk(vb, ncols - (c + 1) * 32);
                        if (dbg_scores != nullptr && row < n_cand) {
#pragma unroll
                            for (int j = 0; j < 32; ++j)
                                if (col0 + (c + 1) * 32 + j < n_ref)
                                    dbg_scores[row * n_ref + col0 + (c + 1) * 32 + j] = vb[j];
                        }
                        process_chunk(vb, col0 + (c + 1) * 32, delta, t);
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&t_empty[acc]);
                ++t_it;
            }
            // ---- results of this candidate row ----
            const bool valid = row < n_cand;
            const bool near_tie = (t.i2 >= 0) && (t.b1 - t.b2 <= delta);
            const bool near_thr = fabsf(t.b1 - thr) <= thr_band;
            const bool flagged = valid && !no_recheck && (near_tie || near_thr);
            const bool full = flagged && (t.b1 - t.b3 <= delta);
            if (valid) {
                keep[row] = (t.b1 >= thr) ? 1 : 0;
                best_idx[row] = static_cast<int32_t>(t.i1 + ref_index_base);
                if (best_val != nullptr) best_val[row] = t.b1;
            }
            const bool pair = flagged && !full;
            const uint32_t pmask = __ballot_sync(0xffffffffu, pair);
            const uint32_t umask = __ballot_sync(0xffffffffu, full);
            if (pmask != 0) {                                  // warp-aggregated append to the two-candidate list
                int32_t slot0 = 0;
                if (lane == 0) slot0 = atomicAdd(&lists.hdr->recheck_count, __popc(pmask));
                slot0 = __shfl_sync(0xffffffffu, slot0, 0);
                if (pair) {
                    const int64_t slot = slot0 + __popc(pmask & ((1u << lane) - 1));
                    if (slot < lists.rec_cap) {
                        RecheckRec r;
                        r.row = static_cast<int32_t>(row);
                        r.idx1 = t.i1;
                        r.idx2 = near_tie ? t.i2 : -1;
                        r.full = 0;
                        lists.recs[slot] = r;
                    }
                }
            }
            if (umask != 0) {                                  // ... and to the full-rescan list
                int32_t slot0 = 0;
                if (lane == 0) slot0 = atomicAdd(&lists.hdr->full_count, __popc(umask));
                slot0 = __shfl_sync(0xffffffffu, slot0, 0);
                if (full) {
                    const int64_t slot = slot0 + __popc(umask & ((1u << lane) - 1));
                    if (slot < lists.full_cap) {
                        lists.full_rows[slot] = static_cast<int32_t>(row);
                        lists.full_keys[slot] = 0ull;
                        if (slot % kFullGroup == 0) lists.full_ctr[slot / kFullGroup] = 0;
                    }
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<kTmemCols>(tmem_base);
    }
}

// ---- host side -----------------------------------------------------------------------------------

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (fn == nullptr) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            return nullptr;
        fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// fp16 matrix [rows, ld] row-major, box = [box_rows, 64 cols], 128-byte swizzle, zero fill out of bounds
int make_tmap(CUtensorMap* m, const __half* base, int64_t rows, int32_t ld, int32_t box_rows) {
    EncodeTiledFn fn = get_encode_fn();
    if (fn == nullptr) { set_error("cuTensorMapEncodeTiled entry point not available (driver too old / no GPU)"); return FFR_ERR_CUDA; }
    cuuint64_t gdim[2] = {static_cast<cuuint64_t>(ld), static_cast<cuuint64_t>(rows)};
    cuuint64_t gstride[1] = {static_cast<cuuint64_t>(ld) * 2};
    cuuint32_t box[2] = {static_cast<cuuint32_t>(kBlockK), static_cast<cuuint32_t>(box_rows)};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<__half*>(base), gdim, gstride, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed: CUresult %d (rows=%lld ld=%d)", (int)r, (long long)rows, ld); return FFR_ERR_CUDA; }
    return FFR_OK;
}

int launch_filter_mma_impl(const __half* ref16, int64_t n_ref, const __half* cand16, int64_t n_cand, int32_t dim_pad,
                           float thr, float delta, float thr_band, int64_t ref_index_base, uint8_t* keep, int32_t* idx, float* val,
                           RecheckLists lists, int no_recheck, float* dbg_scores, cudaStream_t s) {
    if (dim_pad % kBlockK != 0 || dim_pad < kBlockK || dim_pad > 512) {
        set_error("filter_mma: padded dim %d not in {64..512 step 64}", dim_pad);
        return FFR_ERR_UNSUPPORTED;
    }
    if (n_ref >= (int64_t(1) << 31) - 512 || n_cand >= (int64_t(1) << 31) - 512) {
        set_error("filter_mma: n_ref / n_cand must be < 2^31");
        return FFR_ERR_UNSUPPORTED;
    }
    const int kb = dim_pad / kBlockK;
    const uint32_t a_stage = kb * kABlockBytes;
    int a_stages = a_stage <= 32768 ? 3 : (a_stage <= 65536 ? 2 : 1);
    const int64_t n_tiles = (n_cand + kTileM - 1) / kTileM;
    const uint32_t budget = kSmemLimit - kBarrierBytes - 1024;   // 1024: alignment slack
    int b_stages = static_cast<int>((budget - a_stages * a_stage) / kBStageBytes);
    if (b_stages > kMaxBStages) b_stages = kMaxBStages;
    if (b_stages < 2) { set_error("filter_mma: not enough shared memory for dim %d", dim_pad); return FFR_ERR_UNSUPPORTED; }
    const uint32_t smem = a_stages * a_stage + b_stages * kBStageBytes + kBarrierBytes + 1024;

    CUtensorMap tm_c, tm_r;
    int rc = make_tmap(&tm_c, cand16, n_cand, dim_pad, kTileM);
    if (rc != FFR_OK) return rc;
    rc = make_tmap(&tm_r, ref16, n_ref, dim_pad, kTileN);
    if (rc != FFR_OK) return rc;

    static bool attr_set = false;
    if (!attr_set) {
        FFR_CUDA_TRY(cudaFuncSetAttribute(filter_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit));
        attr_set = true;
    }
    const int sms = num_sms();
    const unsigned grid = static_cast<unsigned>(n_tiles < sms ? n_tiles : sms);
    filter_mma_kernel<<<grid, kThreads, smem, s>>>(tm_c, tm_r, n_ref, n_cand, kb, a_stages, b_stages, thr, delta, thr_band,
                                                   ref_index_base, keep, idx, val, lists, no_recheck, dbg_scores);
    FFR_LAUNCH_CHECK("filter_mma");
    return FFR_OK;
}

}  // namespace

int launch_filter_mma(const __half* ref16, int64_t n_ref, const __half* cand16, int64_t n_cand, int32_t dim_pad,
                      float thr, float delta, float thr_band, int64_t ref_index_base, uint8_t* keep, int32_t* idx, float* val,
                      RecheckLists lists, int no_recheck, cudaStream_t s) {
    return launch_filter_mma_impl(ref16, n_ref, cand16, n_cand, dim_pad, thr, delta, thr_band, ref_index_base, keep, idx, val,
                                  lists, no_recheck, nullptr, s);
}

// test hook (not part of the ABI in include/ffr.h): additionally dumps the full score matrix
int launch_filter_mma_debug(const __half* ref16, int64_t n_ref, const __half* cand16, int64_t n_cand, int32_t dim_pad,
                            float thr, float delta, uint8_t* keep, int32_t* idx, float* val, RecheckLists lists,
                            float* scores, cudaStream_t s) {
    return launch_filter_mma_impl(ref16, n_ref, cand16, n_cand, dim_pad, thr, delta, delta, 0, keep, idx, val, lists, 0,
                                  scores, s);
}

}  // namespace ffr
