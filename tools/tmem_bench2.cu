// Micro-benchmark 2: tcgen05.ld throughput while the tensor core is (a) idle, (b) has just written the columns being
// read, (c) is running tcgen05.mma into the OTHER accumulator stage at the same time -- the situation of K2's epilogue.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/tmem_bench2 tools/tmem_bench2.cu && tools/bin/tmem_bench2
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

#define LD32(addr, v)                                                                                              \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                         \
                 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25," \
                 "%26,%27,%28,%29,%30,%31}, [%32];"                                                                \
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), \
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]),        \
                   "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]),      \
                   "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]),      \
                   "=r"(v[29]), "=r"(v[30]), "=r"(v[31])                                                           \
                 : "r"(addr) : "memory")
#define WAIT_LD() asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory")

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) { while (!mbar_try_wait(bar, parity)) { } }
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint64_t desc_sw128(uint32_t a) {
    return (uint64_t)((a & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}

// mma_mode 0: tensor core idle, TMEM never written.  1: 8 MMAs (128x256x16) write columns 0..255 first, then the loads run.
// 2: MMAs run continuously into columns 256..511 while columns 0..255 are read.  3 = 1 + 2.
// ld_mode 0: ld; wait (serial).  1: four loads in flight, one wait.
__global__ void __launch_bounds__(288, 1) bench(int mma_mode, int ld_mode, int nwarps, int iters, long long* out, uint32_t* sink) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint32_t slot;
    __shared__ uint64_t bar;
    __shared__ volatile int stop;
    const int warp = threadIdx.x >> 5;
    uint8_t* base_s = smem + ((1024 - (smem_u32(smem) & 1023)) & 1023);
    for (int i = threadIdx.x; i < 49152 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(base_s)[i] = 0x3c003c00u;   // fp16 1.0
    if (threadIdx.x == 0) { mbar_init(&bar, 1); stop = 0; asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tbase = slot;
    const uint32_t idesc = (1u << 4) | ((256u >> 3) << 17) | ((128u >> 4) << 24);
    const uint64_t da = desc_sw128(smem_u32(base_s)), db = desc_sw128(smem_u32(base_s + 16384));
    uint32_t ph = 0;
    if (warp == 8 && (mma_mode & 1)) {
        if (elect_one()) {
            for (int k = 0; k < 8; ++k) umma_f16(tbase, da + 2 * (k & 3), db + 2 * (k & 3), idesc, k != 0);
            umma_commit(&bar);
        }
        __syncwarp();
        mbar_wait(&bar, ph); ph ^= 1;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t acc = 0;
    long long t0 = 0, t1 = 0;
    if (warp == 8) {
        long long n = 0;
        if (mma_mode & 2) {
            while (!stop) {
                if (elect_one()) {
                    for (int k = 0; k < 8; ++k) umma_f16(tbase + 256, da + 2 * (k & 3), db + 2 * (k & 3), idesc, k != 0);
                    umma_commit(&bar);
                }
                __syncwarp();
                mbar_wait(&bar, ph); ph ^= 1;
                ++n;
            }
            if (threadIdx.x == 256) out[15] = n;
        }
    } else if (warp < nwarps) {
        const uint32_t base = tbase + ((uint32_t)((warp & 3) * 32) << 16) + (warp >> 2) * 128;
        uint32_t a[32], b[32], c[32], d[32];
        t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            if (ld_mode == 0) {
                LD32(base, a); WAIT_LD(); acc += a[0] ^ a[31];
                LD32(base + 32, b); WAIT_LD(); acc += b[0] ^ b[31];
                LD32(base + 64, c); WAIT_LD(); acc += c[0] ^ c[31];
                LD32(base + 96, d); WAIT_LD(); acc += d[0] ^ d[31];
            } else {
                LD32(base, a); LD32(base + 32, b); LD32(base + 64, c); LD32(base + 96, d); WAIT_LD();
                acc += a[0] ^ a[31] ^ b[0] ^ b[31] ^ c[0] ^ c[31] ^ d[0] ^ d[31];
            }
        }
        t1 = clock64();
        if ((threadIdx.x & 31) == 0) out[warp] = t1 - t0;
        __threadfence_block();
        if (threadIdx.x == 0) stop = 1;        // warp 0 finished: MMA stream may stop (all loaders run the same length)
    }
    sink[threadIdx.x] = acc;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(slot) : "memory");
}

int main() {
    long long* out;
    uint32_t* sink;
    cudaMalloc(&out, 16 * sizeof(long long));
    cudaMalloc(&sink, 288 * sizeof(uint32_t));
    cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 51200);
    const int iters = 2000;
    const char* mm[] = {"tensor core idle, TMEM never written", "columns written by MMA first", "MMA running into the other stage",
                        "written by MMA + MMA running"};
    for (int mma_mode = 0; mma_mode < 4; ++mma_mode)
        for (int ld_mode = 0; ld_mode < 2; ++ld_mode)
            for (int nw : {1, 4, 8}) {
                cudaMemset(out, 0, 16 * sizeof(long long));
                bench<<<1, 288, 51200>>>(mma_mode, ld_mode, nw, iters, out, sink);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
                long long h[16];
                cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
                long long mx = 0;
                for (int w = 0; w < nw; ++w) mx = h[w] > mx ? h[w] : mx;
                printf("%-40s %-12s warps=%d  %7.1f cyc per 128-col part per warp  %6.1f B/cyc/SM", mm[mma_mode],
                       ld_mode ? "4 in flight" : "serial", nw, (double)mx / iters, 16384.0 * nw * iters / (double)mx);
                if (mma_mode & 2) printf("   (%lld MMA batches of 8: %.0f cyc per 128x256x16 MMA)", h[15], (double)mx / (8.0 * h[15]));
                printf("\n");
            }
    return 0;
}
