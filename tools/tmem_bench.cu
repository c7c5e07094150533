// Micro-benchmark of tcgen05.ld (TMEM -> registers) on sm_100a: latency and throughput of the 32x32b shapes for
// 1..8 warps, with 1..4 loads in flight.  Build + run on the GPU box:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/tmem_bench tools/tmem_bench.cu && /tmp/tmem_bench
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
#define LD16(addr, v)                                                                                              \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 "                                                         \
                 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"                                 \
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), \
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]),        \
                   "=r"(v[15])                                                                                     \
                 : "r"(addr) : "memory")
#define WAIT_LD() asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory")

// mode 0: x32, one in flight (ld; wait) ; mode 1: x32, two in flight (ld; ld; wait) ; mode 2: x32 four in flight ;
// mode 3: x16 two in flight ; mode 4: x32 pipelined like K2 (wait; ld next; 20 dependent ALU ops on current)
__global__ void __launch_bounds__(256, 1) tmem_ld_bench(int mode, int nwarps, int iters, long long* out, uint32_t* sink) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16) + (warp >> 2) * 128;
    uint32_t acc = 0;
    long long t0 = 0, t1 = 0;
    if (warp < nwarps) {
        uint32_t a[32], b[32], c[32], d[32];
        t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            if (mode == 0) {
                LD32(base, a); WAIT_LD(); acc += a[0] ^ a[31];
                LD32(base + 32, b); WAIT_LD(); acc += b[0] ^ b[31];
                LD32(base + 64, c); WAIT_LD(); acc += c[0] ^ c[31];
                LD32(base + 96, d); WAIT_LD(); acc += d[0] ^ d[31];
            } else if (mode == 1) {
                LD32(base, a); LD32(base + 32, b); WAIT_LD(); acc += a[0] ^ a[31] ^ b[0] ^ b[31];
                LD32(base + 64, c); LD32(base + 96, d); WAIT_LD(); acc += c[0] ^ c[31] ^ d[0] ^ d[31];
            } else if (mode == 2) {
                LD32(base, a); LD32(base + 32, b); LD32(base + 64, c); LD32(base + 96, d); WAIT_LD();
                acc += a[0] ^ a[31] ^ b[0] ^ b[31] ^ c[0] ^ c[31] ^ d[0] ^ d[31];
            } else if (mode == 3) {
                for (int q = 0; q < 4; ++q) {
                    LD16(base + q * 32, a); LD16(base + q * 32 + 16, b); WAIT_LD(); acc += a[0] ^ a[15] ^ b[0] ^ b[15];
                }
            } else {
                LD32(base, a);
                for (int q = 0; q < 4; q += 2) {
                    WAIT_LD();
                    LD32(base + (q + 1) * 32, b);
                    uint32_t m = a[0];
#pragma unroll
                    for (int j = 1; j < 32; ++j) m = max(m, a[j] + j);
                    acc += m;
                    WAIT_LD();
                    if (q + 2 < 4) LD32(base + (q + 2) * 32, a);
                    m = b[0];
#pragma unroll
                    for (int j = 1; j < 32; ++j) m = max(m, b[j] + j);
                    acc += m;
                }
            }
        }
        t1 = clock64();
    }
    if ((threadIdx.x & 31) == 0 && warp < nwarps) out[warp] = t1 - t0;
    sink[threadIdx.x] = acc;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(slot) : "memory");
}

int main() {
    long long* out;
    uint32_t* sink;
    cudaMalloc(&out, 8 * sizeof(long long));
    cudaMalloc(&sink, 256 * sizeof(uint32_t));
    const int iters = 2000;
    const char* names[] = {"x32 1 in flight", "x32 2 in flight", "x32 4 in flight", "x16 2 in flight", "x32 K2-style pipeline + 31-op chain"};
    for (int mode = 0; mode < 5; ++mode) {
        for (int nw : {1, 2, 4, 8}) {
            tmem_ld_bench<<<1, 256>>>(mode, nw, iters, out, sink);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
            long long h[8];
            cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
            long long mx = 0;
            for (int w = 0; w < nw; ++w) mx = h[w] > mx ? h[w] : mx;
            // every iteration moves 4 x (32 lanes x 32 cols x 4 B) = 16 KB per warp
            printf("%-40s warps=%d  %7.1f cyc per 128-col part per warp   %6.1f B/cyc/SM\n", names[mode], nw,
                   (double)mx / iters, 16384.0 * nw * iters / (double)mx);
        }
    }
    return 0;
}
