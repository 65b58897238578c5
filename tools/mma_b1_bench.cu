// Micro-benchmark: the only tensor-core formulation of AND + popcount, the legacy warp-level
//   mma.sync.aligned.m16n8k256.row.col.s32.b1.b1.s32.and.popc
// on sm_100a (tcgen05.mma has no 1-bit kind).  One instruction = 16 x 8 outputs x 256 bits = 32768 bit-ANDs counted, i.e. the
// work of 1024 POPC32.  Prints the sustained rate per SM and for the chip next to the POPC pipe's (hpgv_epi_pipe_peak, 15.5
// POPC32/clk/SM measured).  SURVEY 7.3 asked for this number; DESIGN.md section 8 quotes it.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/mma_b1_bench tools/mma_b1_bench.cu && tools/mma_b1_bench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void mma_b1_kernel(int iters, uint32_t seed, int *sink) {
    uint32_t a[4], b[2];
    int c[4][4];                                            // four independent accumulator sets: no dependent-issue stall
    for (int x = 0; x < 4; x++) a[x] = seed * (threadIdx.x + 1) + x * 0x9E3779B9u;
    for (int x = 0; x < 2; x++) b[x] = seed * (threadIdx.x + 7) + x * 0x85EBCA6Bu + blockIdx.x;
    for (int s = 0; s < 4; s++) for (int x = 0; x < 4; x++) c[s][x] = 0;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int s = 0; s < 4; s++)
            asm volatile("mma.sync.aligned.m16n8k256.row.col.s32.b1.b1.s32.and.popc {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                         : "+r"(c[s][0]), "+r"(c[s][1]), "+r"(c[s][2]), "+r"(c[s][3])
                         : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
    }
    int r = 0;
    for (int s = 0; s < 4; s++) for (int x = 0; x < 4; x++) r ^= c[s][x];
    if (r == 0x12345678) sink[0] = r;
}

int main() {
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    int *sink;
    cudaMalloc(&sink, 4);
    const int threads = 256, grid = prop.multiProcessorCount * 8, iters = 20000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    mma_b1_kernel<<<grid, threads>>>(iters / 10, 12345u, sink);
    cudaEventRecord(e0);
    mma_b1_kernel<<<grid, threads>>>(iters, 12345u, sink);
    cudaEventRecord(e1);
    cudaError_t e = cudaEventSynchronize(e1);
    if (e != cudaSuccess) { printf("mma b1: %s\n", cudaGetErrorString(e)); return 1; }
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double mmas = 4.0 * iters * (threads / 32) * (double) grid;              // warp-level instructions
    const double popc_equiv = mmas * 1024.0 / (ms * 1e-3);                         // POPC32-equivalents per second
    const double clk = prop.clockRate * 1e3;
    printf("%s, %d SMs, %.0f MHz\n", prop.name, prop.multiProcessorCount, clk / 1e6);
    printf("mma.sync m16n8k256 b1 and.popc: %.3f G warp-instructions/s = %.2f per clk per SM\n", mmas / (ms * 1e-3) / 1e9,
           mmas / (ms * 1e-3) / clk / prop.multiProcessorCount);
    printf("  = %.1f T POPC32-equivalents/s (%.0f per clk per SM) vs 4.5 T/s (15.5 per clk per SM) of the POPC pipe\n", popc_equiv / 1e12,
           popc_equiv / clk / prop.multiProcessorCount);
    return 0;
}
