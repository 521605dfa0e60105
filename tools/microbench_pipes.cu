// Micro-benchmark: do PRMT (ALU pipe) and IDP4A / IMAD (FMA pipe?) issue in parallel on sm_100a?
// Times three loops at full occupancy: PRMT only, DP4A only, and the two interleaved.  If the
// interleaved loop takes max(a, b) rather than a + b, the two run on different pipes.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o microbench_pipes microbench_pipes.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) k(uint32_t* out, int iters)
{
    uint32_t a = threadIdx.x * 2654435761u, b = blockIdx.x + 12345u, c = a ^ b, d = a + b;
    uint32_t e = a * 3u, f = b * 5u, g = c * 7u, h = d * 11u;
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (MODE == 0 || MODE == 2) {
                asm volatile("prmt.b32 %0, %0, %1, 0x5140;" : "+r"(a) : "r"(b));
                asm volatile("prmt.b32 %0, %0, %1, 0x3610;" : "+r"(c) : "r"(d));
            }
            if (MODE == 1 || MODE == 2) {
                asm volatile("dp4a.u32.u32 %0, %1, 0x00000100, %0;" : "+r"(e) : "r"(f));
                asm volatile("dp4a.u32.u32 %0, %1, 0x00010000, %0;" : "+r"(g) : "r"(h));
            }
            if (MODE == 3) {
                asm volatile("mad.lo.u32 %0, %1, 269, %0;" : "+r"(e) : "r"(f));
                asm volatile("mad.lo.u32 %0, %1, 77, %0;" : "+r"(g) : "r"(h));
            }
            if (MODE == 4) {
                asm volatile("prmt.b32 %0, %0, %1, 0x5140;" : "+r"(a) : "r"(b));
                asm volatile("prmt.b32 %0, %0, %1, 0x3610;" : "+r"(c) : "r"(d));
                asm volatile("mad.lo.u32 %0, %1, 269, %0;" : "+r"(e) : "r"(f));
                asm volatile("mad.lo.u32 %0, %1, 77, %0;" : "+r"(g) : "r"(h));
            }
        }
    }
    out[blockIdx.x * 256 + threadIdx.x] = a ^ c ^ e ^ g;
}

template <int MODE>
float run(uint32_t* d_out, int iters)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<148 * 8, 256>>>(d_out, iters);
    cudaEventRecord(e0);
    k<MODE><<<148 * 8, 256>>>(d_out, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    return ms;
}

int main()
{
    uint32_t* d_out;
    cudaMalloc(&d_out, 148 * 8 * 256 * 4);
    const int iters = 20000;
    printf("prmt only            %.3f ms\n", run<0>(d_out, iters));
    printf("dp4a only            %.3f ms\n", run<1>(d_out, iters));
    printf("prmt + dp4a          %.3f ms\n", run<2>(d_out, iters));
    printf("imad only            %.3f ms\n", run<3>(d_out, iters));
    printf("prmt + imad          %.3f ms\n", run<4>(d_out, iters));
    return 0;
}
