// Microbenchmark (diagnostic): issue cost of SFU and FP32 instructions per SM sub-partition on this GPU.
// Each thread runs ILP independent chains of one instruction type; cycles per warp-instruction per SMSP is reported for
// 1, 2, 4 and 8 warps per SMSP.  Build: nvcc -arch=sm_100a -O3 -o pipes pipes.cu
#include <cstdio>
#include <cuda_runtime.h>
#define ILP 8
#define ITERS 512
template <int OP>
__global__ void k(float* out, long long* cyc, float seed) {
    float v[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) v[i] = seed + 0.001f * (threadIdx.x + i);
    float c = seed * 0.5f + 1.0f;
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            if (OP == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(v[i]));
            if (OP == 1) asm volatile("lg2.approx.ftz.f32 %0, %0;" : "+f"(v[i]));
            if (OP == 2) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(v[i]));
            if (OP == 3) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(v[i]) : "f"(c), "f"(seed));
            if (OP == 4) asm volatile("mul.f32 %0, %0, %1;" : "+f"(v[i]) : "f"(c));
            if (OP == 5) asm volatile("add.f32 %0, %0, %1;" : "+f"(v[i]) : "f"(c));
            if (OP == 6) asm volatile("fma.rn.f32 %0, %0, 0f3F800347, 0f3A83126F;" : "+f"(v[i]));
            if (OP == 7) asm volatile("max.f32 %0, %0, %1;" : "+f"(v[i]) : "f"(c));
            if (OP == 8 && (i & 1) == 0) {  // packed: two floats per instruction (counted as one instruction per pair)
                unsigned long long a, b2, c2;
                asm volatile("mov.b64 %0, {%1, %2};" : "=l"(a) : "f"(v[i]), "f"(v[i + 1]));
                asm volatile("mov.b64 %0, {%1, %1};" : "=l"(b2) : "f"(c));
                asm volatile("mov.b64 %0, {%1, %1};" : "=l"(c2) : "f"(seed));
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(a) : "l"(b2), "l"(c2));
                asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(v[i]), "=f"(v[i + 1]) : "l"(a));
            }
        }
    }
    long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int OP>
void run(const char* name) {
    float* out; long long* cyc;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
    printf("%-10s", name);
    for (int wps : {1, 2, 4, 8}) {
        int threads = wps * 4 * 32;
        k<OP><<<148, threads>>>(out, cyc, 1.0f);
        k<OP><<<148, threads>>>(out, cyc, 1.0f);
        cudaDeviceSynchronize();
        long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
        // warp-instructions per SMSP = wps * ITERS * ILP
        printf("  %dw/smsp: %.2f cyc/inst", wps, avg / ((double)wps * ITERS * ILP));
    }
    printf("\n");
}
int main() {
    run<0>("ex2"); run<1>("lg2"); run<2>("rcp"); run<3>("ffma 3reg"); run<4>("fmul"); run<5>("fadd"); run<6>("ffma imm"); run<7>("fmnmx"); run<8>("ffma2 x4");
    return 0;
}
