// Pipe-throughput microbenchmarks used to size the fused FFT kernel (DESIGN.md §"instruction budget").
// Measures per-SM issue rates of scalar FFMA/FADD vs the packed fma/add.f32x2 forms and MUFU.
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 4096
template <int OP>
__global__ void __launch_bounds__(256) k(float* out, float a, float b) {
    float x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    unsigned long long p0, p1, p2, p3, pa, pb;
    asm("mov.b64 %0, {%1,%2};" : "=l"(p0) : "f"(x0), "f"(x1));
    asm("mov.b64 %0, {%1,%2};" : "=l"(p1) : "f"(x2), "f"(x3));
    asm("mov.b64 %0, {%1,%2};" : "=l"(p2) : "f"(x4), "f"(x5));
    asm("mov.b64 %0, {%1,%2};" : "=l"(p3) : "f"(x6), "f"(x7));
    asm("mov.b64 %0, {%1,%1};" : "=l"(pa) : "f"(a));
    asm("mov.b64 %0, {%1,%1};" : "=l"(pb) : "f"(b));
#pragma unroll 1
    for (int i = 0; i < ITERS; ++i) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (OP == 0) {  // 8 scalar FFMA
                x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
                x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
            } else if (OP == 1) {  // 4 packed FFMA2 (= 8 lanes-flops)
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p0) : "l"(pa), "l"(pb));
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p1) : "l"(pa), "l"(pb));
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p2) : "l"(pa), "l"(pb));
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p3) : "l"(pa), "l"(pb));
            } else if (OP == 2) {  // 8 scalar FADD
                x0 += a; x1 += a; x2 += a; x3 += a; x4 += a; x5 += a; x6 += a; x7 += a;
            } else if (OP == 3) {  // 4 packed FADD2
                asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p0) : "l"(pa));
                asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p1) : "l"(pa));
                asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p2) : "l"(pa));
                asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p3) : "l"(pa));
            } else if (OP == 4) {  // 8 MUFU.EX2
                asm volatile("ex2.approx.f32 %0, %0;" : "+f"(x0)); asm volatile("ex2.approx.f32 %0, %0;" : "+f"(x1));
                asm volatile("ex2.approx.f32 %0, %0;" : "+f"(x2)); asm volatile("ex2.approx.f32 %0, %0;" : "+f"(x3));
                asm volatile("ex2.approx.f32 %0, %0;" : "+f"(x4)); asm volatile("ex2.approx.f32 %0, %0;" : "+f"(x5));
                asm volatile("ex2.approx.f32 %0, %0;" : "+f"(x6)); asm volatile("ex2.approx.f32 %0, %0;" : "+f"(x7));
            } else if (OP == 5) {  // 4 FFMA + 4 FADD interleaved (scalar mix like a butterfly)
                x0 = fmaf(x0, a, b); x1 += a; x2 = fmaf(x2, a, b); x3 += a;
                x4 = fmaf(x4, a, b); x5 += a; x6 = fmaf(x6, a, b); x7 += a;
            } else if (OP == 6) {  // 2 FFMA2 + 2 FADD2
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p0) : "l"(pa), "l"(pb));
                asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p1) : "l"(pa));
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p2) : "l"(pa), "l"(pb));
                asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p3) : "l"(pa));
            }
        }
    }
    float r0, r1;
    asm("mov.b64 {%0,%1}, %2;" : "=f"(r0), "=f"(r1) : "l"(p0 ^ p1 ^ p2 ^ p3));
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7 + r0 + r1;
}

template <int OP>
void run(const char* name, int lanes_per_iter, int instr_per_iter, int sms) {
    float* out;
    const int blocks = sms * 8, threads = 256;
    cudaMalloc(&out, blocks * threads * 4);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<OP><<<blocks, threads>>>(out, 1.0001f, 0.0001f);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    for (int r = 0; r < 5; ++r) k<OP><<<blocks, threads>>>(out, 1.0001f, 0.0001f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
    const double thr = (double)blocks * threads * ITERS * 4;
    printf("%-28s %8.3f ms  %8.2f T lane-ops/s  %8.2f T warp-instr-lanes/s (issue)\n", name, ms,
           thr * lanes_per_iter / ms * 1e-9, thr * instr_per_iter / ms * 1e-9);
    cudaFree(out);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("%s sms=%d clock=%d kHz\n", p.name, p.multiProcessorCount, clk);
    const int s = p.multiProcessorCount;
    run<0>("FFMA scalar x8", 8, 8, s);
    run<1>("FFMA2 packed x4", 8, 4, s);
    run<2>("FADD scalar x8", 8, 8, s);
    run<3>("FADD2 packed x4", 8, 4, s);
    run<4>("MUFU.EX2 x8", 8, 8, s);
    run<5>("FFMA/FADD mix x8", 8, 8, s);
    run<6>("FFMA2/FADD2 mix x4", 8, 4, s);
    return 0;
}
