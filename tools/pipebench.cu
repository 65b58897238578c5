// pipebench.cu -- per-SM issue rates of the integer instructions the search kernels are built from (sm_100a).
// Build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/pipebench tools/pipebench.cu
// Prints ops/clk/SM for single instructions and for pairs issued together (a pair that takes max(a, b) runs on
// two pipes, a pair that takes a + b shares one).  Test tooling: not linked into the product.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define DEV __device__ __forceinline__

struct OpPopc  { static DEV void f(uint32_t &a, uint32_t &b, uint32_t c) { a += __popc(b ^ a); } static constexpr int n = 1; static const char *name() { return "POPC(+LOP+IADD)"; } };
struct OpPopcOnly { static DEV void f(uint32_t &a, uint32_t &b, uint32_t c) { a = __popc(a) | c; } static constexpr int n = 1; static const char *name() { return "POPC(+LOP)"; } };
struct OpLop3  { static DEV void f(uint32_t &a, uint32_t &b, uint32_t c) { asm("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a) : "r"(b), "r"(c)); } static constexpr int n = 1; static const char *name() { return "LOP3"; } };
struct OpIadd3 { static DEV void f(uint32_t &a, uint32_t &b, uint32_t c) { a = a + b + c; } static constexpr int n = 1; static const char *name() { return "IADD3"; } };
struct OpImad  { static DEV void f(uint32_t &a, uint32_t &b, uint32_t c) { asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(a) : "r"(b), "r"(c)); } static constexpr int n = 1; static const char *name() { return "IMAD"; } };
struct OpIdp4a { static DEV void f(uint32_t &a, uint32_t &b, uint32_t c) { asm("dp4a.u32.s32 %0, %1, %2, %0;" : "+r"(a) : "r"(b), "r"(c)); } static constexpr int n = 1; static const char *name() { return "IDP.4A"; } };
struct OpVabs  { static DEV void f(uint32_t &a, uint32_t &b, uint32_t c) { asm("sad.s32 %0, %1, %2, %0;" : "+r"(a) : "r"(b), "r"(c)); } static constexpr int n = 1; static const char *name() { return "VABSDIFF"; } };
struct OpVimnmxRelu { static DEV void f(uint32_t &a, uint32_t &b, uint32_t c) { asm("max.s32.relu %0, %0, %1;" : "+r"(a) : "r"(b)); a ^= c; } static constexpr int n = 1; static const char *name() { return "VIMNMX.RELU(+LOP)"; } };
struct OpViadd16 { static DEV void f(uint32_t &a, uint32_t &b, uint32_t c) { asm("add.s16x2 %0, %0, %1;" : "+r"(a) : "r"(b)); } static constexpr int n = 1; static const char *name() { return "VIADD.16x2"; } };
struct OpVimnmx16 { static DEV void f(uint32_t &a, uint32_t &b, uint32_t c) { asm("max.s16x2.relu %0, %0, %1;" : "+r"(a) : "r"(b)); a ^= c; } static constexpr int n = 1; static const char *name() { return "VIMNMX.S16x2.RELU(+LOP)"; } };
struct OpPrmt  { static DEV void f(uint32_t &a, uint32_t &b, uint32_t c) { asm("prmt.b32 %0, %0, %1, %2;" : "+r"(a) : "r"(b), "r"(c & 0x7777u)); } static constexpr int n = 1; static const char *name() { return "PRMT(+LOP)"; } };
struct OpSetpAdd { static DEV void f(uint32_t &a, uint32_t &b, uint32_t c) { asm("{ .reg .pred p; setp.ge.s32 p, %0, %2; @p add.u32 %0, %0, %1; }" : "+r"(a) : "r"(b), "r"(c)); } static constexpr int n = 2; static const char *name() { return "ISETP+@p IADD (2 ops)"; } };
struct OpShf   { static DEV void f(uint32_t &a, uint32_t &b, uint32_t c) { a = __funnelshift_l(a, b, c); } static constexpr int n = 1; static const char *name() { return "SHF"; } };
struct OpNone  { static DEV void f(uint32_t &a, uint32_t &b, uint32_t c) {} static constexpr int n = 0; static const char *name() { return "-"; } };

// RA x op A and RB x op B per inner step on 8 independent chains
template <class A, int RA, class B, int RB>
__global__ void __launch_bounds__(512, 1) bench_kernel(int iters, uint32_t seed, uint32_t *sink) {
    uint32_t x[8], y[8], z[8];
#pragma unroll
    for (int q = 0; q < 8; q++) { x[q] = seed * (threadIdx.x + 1) + q * 0x9E3779B9u; y[q] = x[q] * 2654435761u + blockIdx.x; z[q] = y[q] ^ 0x5bd1e995u; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
#pragma unroll
            for (int ra = 0; ra < RA; ra++)
#pragma unroll
                for (int q = 0; q < 8; q++) A::f(x[q], y[q], z[(q + ra) & 7]);
#pragma unroll
            for (int rb = 0; rb < RB; rb++)
#pragma unroll
                for (int q = 0; q < 8; q++) B::f(z[q], y[(q + 1) & 7], x[(q + rb) & 7] | 1u);
        }
    }
    uint32_t r = 0;
#pragma unroll
    for (int q = 0; q < 8; q++) r ^= x[q] ^ y[q] ^ z[q];
    if (r == 0x12345678u) sink[0] = r;
}

template <class A, int RA, class B, int RB>
void run(int sms, double mhz, uint32_t *sink) {
    const int iters = 2000, threads = 512;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    bench_kernel<A, RA, B, RB><<<sms, threads>>>(200, 7u, sink);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int rep = 0; rep < 3; rep++) {
        cudaEventRecord(e0);
        bench_kernel<A, RA, B, RB><<<sms, threads>>>(iters, 7u, sink);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    const double steps = (double) iters * 4 * 8 * threads;              // per SM
    const double clk = best * 1e-3 * mhz * 1e6;
    const double a_ops = steps * RA * A::n, b_ops = steps * RB * B::n;
    printf("%-26s x%d  %-26s x%d  %8.3f ms  A %7.2f /clk/SM  B %7.2f /clk/SM  (A+B %7.2f)\n", A::name(), RA, B::name(), RB, best,
           a_ops / clk, b_ops / clk, (a_ops + b_ops) / clk);
    fflush(stdout);
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double mhz = khz / 1000.0;
    printf("%s, %d SMs, %.0f MHz (rates assume the SM clock stays at max)\n", p.name, p.multiProcessorCount, mhz);
    uint32_t *sink; cudaMalloc(&sink, 4);
    const int sms = p.multiProcessorCount;
    run<OpPopc, 1, OpNone, 0>(sms, mhz, sink);
    run<OpPopcOnly, 1, OpNone, 0>(sms, mhz, sink);
    run<OpLop3, 1, OpNone, 0>(sms, mhz, sink);
    run<OpIadd3, 1, OpNone, 0>(sms, mhz, sink);
    run<OpImad, 1, OpNone, 0>(sms, mhz, sink);
    run<OpIdp4a, 1, OpNone, 0>(sms, mhz, sink);
    run<OpVabs, 1, OpNone, 0>(sms, mhz, sink);
    run<OpVimnmxRelu, 1, OpNone, 0>(sms, mhz, sink);
    run<OpViadd16, 1, OpNone, 0>(sms, mhz, sink);
    run<OpVimnmx16, 1, OpNone, 0>(sms, mhz, sink);
    run<OpPrmt, 1, OpNone, 0>(sms, mhz, sink);
    run<OpSetpAdd, 1, OpNone, 0>(sms, mhz, sink);
    run<OpShf, 1, OpNone, 0>(sms, mhz, sink);
    // pairs: does B share a pipe with A?
    run<OpPopcOnly, 1, OpIdp4a, 2>(sms, mhz, sink);
    run<OpPopcOnly, 1, OpVabs, 2>(sms, mhz, sink);
    run<OpPopcOnly, 1, OpImad, 2>(sms, mhz, sink);
    run<OpPopcOnly, 1, OpLop3, 2>(sms, mhz, sink);
    run<OpPopcOnly, 1, OpLop3, 3>(sms, mhz, sink);
    run<OpLop3, 1, OpImad, 1>(sms, mhz, sink);
    run<OpLop3, 1, OpIdp4a, 1>(sms, mhz, sink);
    run<OpLop3, 1, OpVabs, 1>(sms, mhz, sink);
    run<OpLop3, 1, OpIadd3, 1>(sms, mhz, sink);
    run<OpImad, 1, OpIdp4a, 1>(sms, mhz, sink);
    run<OpImad, 1, OpVabs, 1>(sms, mhz, sink);
    run<OpLop3, 1, OpViadd16, 1>(sms, mhz, sink);
    run<OpLop3, 1, OpPrmt, 1>(sms, mhz, sink);
    return 0;
}
