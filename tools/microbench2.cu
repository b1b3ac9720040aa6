// Does a packed FFMA2 occupy the warp scheduler's issue port for 2 cycles, or only the FMA pipe?
// Mixes 4 FFMA2 with 4 independent integer (IADD3/LOP3) or shared-memory (LDS) instructions per iteration.
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 4096
template <int OP>
__global__ void __launch_bounds__(256) k(float* out, float a, float b, int ia) {
    __shared__ float sm[1024];
    sm[threadIdx.x] = a; sm[threadIdx.x + 256] = b; sm[threadIdx.x + 512] = a; sm[threadIdx.x + 768] = b;
    __syncthreads();
    unsigned long long p0, p1, p2, p3, pa, pb;
    float x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3;
    asm("mov.b64 %0, {%1,%2};" : "=l"(p0) : "f"(x0), "f"(x1));
    asm("mov.b64 %0, {%1,%2};" : "=l"(p1) : "f"(x2), "f"(x3));
    asm("mov.b64 %0, {%1,%2};" : "=l"(p2) : "f"(x1), "f"(x2));
    asm("mov.b64 %0, {%1,%2};" : "=l"(p3) : "f"(x3), "f"(x0));
    asm("mov.b64 %0, {%1,%1};" : "=l"(pa) : "f"(a));
    asm("mov.b64 %0, {%1,%1};" : "=l"(pb) : "f"(b));
    int i0 = threadIdx.x, i1 = i0 + 1, i2 = i0 + 2, i3 = i0 + 3;
    float l0 = 0, l1 = 0, l2 = 0, l3 = 0;
    const float* sp = sm + (threadIdx.x & 255);
#pragma unroll 1
    for (int i = 0; i < ITERS; ++i) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (OP == 0 || OP == 2 || OP == 4) {
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p0) : "l"(pa), "l"(pb));
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p1) : "l"(pa), "l"(pb));
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p2) : "l"(pa), "l"(pb));
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p3) : "l"(pa), "l"(pb));
            }
            if (OP == 1 || OP == 2) {
                asm volatile("add.s32 %0, %0, %1;" : "+r"(i0) : "r"(ia));
                asm volatile("xor.b32 %0, %0, %1;" : "+r"(i1) : "r"(ia));
                asm volatile("add.s32 %0, %0, %1;" : "+r"(i2) : "r"(ia));
                asm volatile("xor.b32 %0, %0, %1;" : "+r"(i3) : "r"(ia));
            }
            if (OP == 3 || OP == 4) {
                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(l0) : "l"(sp));
                asm volatile("ld.shared.f32 %0, [%1+1024];" : "=f"(l1) : "l"(sp));
                asm volatile("ld.shared.f32 %0, [%1+2048];" : "=f"(l2) : "l"(sp));
                asm volatile("ld.shared.f32 %0, [%1+3072];" : "=f"(l3) : "l"(sp));
            }
        }
    }
    float r0, r1;
    asm("mov.b64 {%0,%1}, %2;" : "=f"(r0), "=f"(r1) : "l"(p0 ^ p1 ^ p2 ^ p3));
    out[blockIdx.x * blockDim.x + threadIdx.x] = r0 + r1 + i0 + i1 + i2 + i3 + l0 + l1 + l2 + l3;
}
template <int OP> float run(const char* name, int sms) {
    float* out; const int blocks = sms * 8, threads = 256;
    cudaMalloc(&out, blocks * threads * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<OP><<<blocks, threads>>>(out, 1.0001f, 0.0001f, 3); cudaDeviceSynchronize();
    cudaEventRecord(e0);
    for (int r = 0; r < 5; ++r) k<OP><<<blocks, threads>>>(out, 1.0001f, 0.0001f, 3);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
    printf("%-36s %8.3f ms\n", name, ms); cudaFree(out); return ms;
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0); const int s = p.multiProcessorCount;
    run<0>("4 FFMA2", s); run<1>("4 INT (IADD/LOP)", s); run<2>("4 FFMA2 + 4 INT", s);
    run<3>("4 LDS.32", s); run<4>("4 FFMA2 + 4 LDS.32", s);
    return 0;
}
