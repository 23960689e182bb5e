// pipe_rate.cu -- developer probe: issue rate of a few instructions the requantise epilogues could use, on one SM's
// worth of warps (results per clock per SM).  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_rate pipe_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int OP>
__global__ void __launch_bounds__(1024) rate_kernel(float seed, int iters, int *out, long long *cycles)
{
    float f[8];
    int q[8];
#pragma unroll
    for (int j = 0; j < 8; j++) f[j] = seed + threadIdx.x * 0.37f + j * 1.7f, q[j] = threadIdx.x + j;
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int j = 0; j < 8; j++) {
            if (OP == 0) asm volatile("cvt.rni.sat.s8.f32 %0, %1;" : "=r"(q[j]) : "f"(f[j] + (float)q[j]));
            if (OP == 1) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(f[j]) : "f"(1.0f));
            if (OP == 2) asm volatile("max.s32 %0, %0, %1;" : "+r"(q[j]) : "r"(i));
            if (OP == 3) asm volatile("cvt.rni.s32.f32 %0, %1;" : "=r"(q[j]) : "f"(f[j] + (float)q[j]));
            if (OP == 4) asm volatile("prmt.b32 %0, %0, %1, 0x3201;" : "+r"(q[j]) : "r"(i));
            if (OP == 5) asm volatile("cvt.rn.f32.s32 %0, %1;" : "=f"(f[j]) : "r"(q[j] + i));
        }
    }
    const long long t1 = clock64();
    int acc = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) acc += q[j] + (int)f[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

int main()
{
    int *out;
    long long *cyc, h;
    cudaMalloc(&out, 1024 * 4);
    cudaMalloc(&cyc, 8);
    const int iters = 4096;
    const char *names[] = {"cvt.rni.sat.s8.f32 (F2I.S8) + FADD + I2F", "add.f32 (FADD)", "max.s32 (IMNMX)", "cvt.rni.s32.f32 (F2I) + FADD + I2F",
                           "prmt (PRMT)", "cvt.rn.f32.s32 (I2F) + IADD"};
#define RUN(OP)                                                                                          \
    rate_kernel<OP><<<1, 1024>>>(1.0f, iters, out, cyc);                                               \
    cudaDeviceSynchronize();                                                                             \
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);                                                      \
    printf("%-45s %8.1f thread-ops per clock per SM (%lld cycles)\n", names[OP], 1024.0 * 8 * iters / h, h);
    RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5)
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
