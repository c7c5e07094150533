// Micro-benchmark 3: cost of the K2 epilogue's code STRUCTURE.  Eight warps read a 128 x 256 fp32 accumulator tile out of
// TMEM (values all zero, gate = +inf or 0 so the update path never / always runs) with different loop shapes:
//   0 bare loads   1 per-chunk max tree, update path inline after every chunk (K2 up to round 1)
//   2 four loads in flight, four trees, ONE branch per tile guarding the update paths
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/epi_bench tools/epi_bench.cu && tools/bin/epi_bench
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ float fmax3(float a, float b, float c) { float r; asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, float (&v)[32]) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
struct Top3 { float b1, b2, b3, b4; int32_t i1, i2, i3; };
__device__ __forceinline__ void top3_insert(Top3& t, float v, int32_t idx) {
    const bool g1 = v > t.b1, g2 = v > t.b2, g3 = v > t.b3, g4 = v > t.b4;
    t.b4 = g3 ? t.b3 : (g4 ? v : t.b4);
    t.i3 = g2 ? t.i2 : (g3 ? idx : t.i3);
    t.b3 = g2 ? t.b2 : (g3 ? v : t.b3);
    t.i2 = g1 ? t.i1 : (g2 ? idx : t.i2);
    t.b2 = g1 ? t.b1 : (g2 ? v : t.b2);
    t.i1 = g1 ? idx : t.i1;
    t.b1 = g1 ? v : t.b1;
}
__device__ __forceinline__ float chunk_max(const float (&v)[32]) {
    float s[4];
#pragma unroll
    for (int k = 0; k < 4; ++k)
        s[k] = fmax3(fmax3(v[8 * k + 0], v[8 * k + 1], v[8 * k + 2]), fmax3(v[8 * k + 3], v[8 * k + 4], v[8 * k + 5]),
                     fmaxf(v[8 * k + 6], v[8 * k + 7]));
    return fmax3(s[0], s[1], fmaxf(s[2], s[3]));
}
// compact update path: the chunk maximum is inserted; if other columns are inside the window too the row is marked
// ambiguous (amb = largest such chunk maximum) and left to the full fp32 rescan
__device__ __forceinline__ void slow_chunk(const float (&v)[32], float cmax, int32_t base, float delta, Top3& t, float& gate, float& amb) {
    const float w = fmaxf(t.b1, cmax) - delta;
    uint32_t g[4], e[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        g[k] = 0; e[k] = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            g[k] |= (v[8 * k + j] >= w) ? (1u << (8 * k + j)) : 0u;
            e[k] |= (v[8 * k + j] == cmax) ? (1u << (8 * k + j)) : 0u;
        }
    }
    const uint32_t ge = (g[0] | g[1]) | (g[2] | g[3]);
    const uint32_t eq = (e[0] | e[1]) | (e[2] | e[3]);
    if (ge & (ge - 1)) amb = fmaxf(amb, cmax);
    top3_insert(t, cmax, base + __ffs(eq) - 1);
    gate = t.b1 - delta;
}

__global__ void __launch_bounds__(256, 1) bench(int mode, float gate0, int iters, long long* out, float* sink) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t taddr = slot + ((uint32_t)((warp & 3) * 32) << 16) + (warp >> 2) * 128;
    Top3 t; t.b1 = t.b2 = t.b3 = t.b4 = -INFINITY; t.i1 = 0; t.i2 = t.i3 = -1;
    float gate = gate0, amb = -INFINITY;
    const float delta = 4e-4f;
    const long long t0 = clock64();
    for (int rt = 0; rt < iters; ++rt) {
        const int32_t col0 = rt * 256 + (warp >> 2) * 128;
        const uint32_t ta = taddr + (rt & 1) * 256;
        if (mode == 0) {
            float va[32];
#pragma unroll
            for (int c = 0; c < 4; ++c) { tmem_ld_32x32(ta + c * 32, va); tmem_ld_wait(); t.b1 = fmax3(t.b1, va[0], va[31]); }
        } else if (mode == 1) {
            float va[32], vb[32];
            tmem_ld_32x32(ta, va);
            tmem_ld_wait();
            tmem_ld_32x32(ta + 32, vb);
            { const float cm = chunk_max(va); if (cm >= gate) slow_chunk(va, cm, col0, delta, t, gate, amb); }
            tmem_ld_wait();
            tmem_ld_32x32(ta + 64, va);
            { const float cm = chunk_max(vb); if (cm >= gate) slow_chunk(vb, cm, col0 + 32, delta, t, gate, amb); }
            tmem_ld_wait();
            tmem_ld_32x32(ta + 96, vb);
            { const float cm = chunk_max(va); if (cm >= gate) slow_chunk(va, cm, col0 + 64, delta, t, gate, amb); }
            tmem_ld_wait();
            { const float cm = chunk_max(vb); if (cm >= gate) slow_chunk(vb, cm, col0 + 96, delta, t, gate, amb); }
        } else {
            float v[4][32];
#pragma unroll
            for (int c = 0; c < 4; ++c) tmem_ld_32x32(ta + c * 32, v[c]);
            tmem_ld_wait();
            float cm[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) cm[c] = chunk_max(v[c]);
            if (fmax3(cm[0], cm[1], fmaxf(cm[2], cm[3])) >= gate) {
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    if (cm[c] >= gate) slow_chunk(v[c], cm[c], col0 + 32 * c, delta, t, gate, amb);
            }
        }
        if (gate0 <= 0.f) gate = gate0;          // "always update" runs: keep the gate open
    }
    const long long t1 = clock64();
    if ((threadIdx.x & 31) == 0) out[warp] = t1 - t0;
    sink[threadIdx.x] = t.b1 + t.b2 + t.b3 + t.b4 + t.i1 + t.i2 + t.i3 + amb;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(slot) : "memory");
}

int main() {
    long long* out; float* sink;
    cudaMalloc(&out, 8 * sizeof(long long));
    cudaMalloc(&sink, 256 * sizeof(float));
    const int iters = 4000;
    const char* names[] = {"bare loads (serial)", "tree + inline update per chunk", "4 loads in flight, one branch per tile"};
    for (int mode = 0; mode < 3; ++mode)
        for (float gate0 : {INFINITY, -1.0f}) {
            bench<<<1, 256>>>(mode, gate0, iters, out, sink);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
            long long h[8];
            cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
            long long mx = 0;
            for (int w = 0; w < 8; ++w) mx = h[w] > mx ? h[w] : mx;
            printf("%-42s %-22s %7.1f cyc per 128x256 tile (8 warps)\n", names[mode], gate0 > 0 ? "update never taken" : "update always taken",
                   (double)mx / iters);
        }
    return 0;
}
