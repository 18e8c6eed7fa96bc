// Micro-benchmark (GPU box only): issue rate of scalar vs packed fp32 on sm_100a.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o f32x2 f32x2.cu && ./f32x2
// Prints warp-instructions per clock per SM sub-partition for FFMA, FMUL+FADD (separately rounded),
// fma.rn.f32x2 and add.rn.f32x2 chains with ILP 8, at 4/8/16 warps per SM sub-partition... (1 CTA/SM).
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned long long pk(float a, float b) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long r;
    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b) {
    unsigned long long r;
    asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}

constexpr int ILP = 8, ITERS = 4096;

template <int MODE>
__global__ void bench(float* out, long long* cycles, float c0, float c1) {
    float acc[ILP];
    unsigned long long acc2[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { acc[i] = threadIdx.x * 0.001f + i; acc2[i] = pk(acc[i], acc[i] + 1.f); }
    const unsigned long long k0 = pk(c0, c0), k1 = pk(c1, c1);
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            if (MODE == 0) acc[i] = fmaf(acc[i], c0, c1);
            if (MODE == 1) acc[i] = __fadd_rn(__fmul_rn(acc[i], c0), c1);
            if (MODE == 2) acc2[i] = fma2(acc2[i], k0, k1);
            if (MODE == 3) acc2[i] = add2(acc2[i], k1);
            if (MODE == 4) acc2[i] = add2(fma2(acc2[i], k0, k1), k1);   // fma result into add: must stay 2 instrs
        }
    }
    long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += acc[i] + __uint_as_float((unsigned)(acc2[i] & 0xffffffffu)) + __uint_as_float((unsigned)(acc2[i] >> 32));
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int instr_per_elem) {
    float* out; long long* cyc;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
    for (int warps : {4, 8, 16, 32}) {
        bench<MODE><<<148, warps * 32>>>(out, cyc, 0.999f, 0.001f);
        cudaDeviceSynchronize();
        bench<MODE><<<148, warps * 32>>>(out, cyc, 0.999f, 0.001f);
        cudaDeviceSynchronize();
        long long h[148]; cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
        double avg = 0; for (auto v : h) avg += v; avg /= 148;
        double winstr = (double)warps * ITERS * ILP * instr_per_elem;     // warp-instructions per SM
        printf("%-28s warps/SM %2d  cycles %9.0f  warp-instr/clk/SMSP %.3f\n", name, warps, avg, winstr / avg / 4);
    }
    cudaFree(out); cudaFree(cyc);
}

int main() {
    run<0>("FFMA (scalar)", 1);
    run<1>("FMUL+FADD (scalar, rn)", 2);
    run<2>("fma.rn.f32x2", 1);
    run<3>("add.rn.f32x2", 1);
    run<4>("fma.f32x2 -> add.f32x2", 2);
    cudaError_t e = cudaGetLastError();
    printf("status: %s\n", cudaGetErrorString(e));
    return 0;
}
