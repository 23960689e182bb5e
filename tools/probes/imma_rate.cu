// imma_rate.cu -- developer probe: what the legacy warp-level tensor path (mma.sync.m16n8k32 s8, SASS IMMA.16832.S8.S8) delivers
// on one B200 SM, alone and beside integer-ALU work.  A depthwise 3x3 written as diagonal-B warp MMAs needs ~5 of them per
// 128 outputs; whether that beats 12 dp4a + 6 PRMT depends on this rate.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o imma_rate imma_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void imma(int (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2])
{
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// the same with eight different A and B operands in rotation (what a real kernel issues: the first variant's single A / B
// pair sits in the operand-reuse cache)
__global__ void __launch_bounds__(512) imma_varied_kernel(int iters, int *out, long long *cycles)
{
    int d[8][4];
    uint32_t a[8][4], b[8][2];
#pragma unroll
    for (int j = 0; j < 8; j++) {
#pragma unroll
        for (int e = 0; e < 4; e++) d[j][e] = 0, a[j][e] = threadIdx.x * 0x01010101u + j * 4 + e;
        b[j][0] = 0x01020304u + j, b[j][1] = 0x04030201u + 3 * j;
    }
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < iters; i += 8) {
#pragma unroll
        for (int ii = 0; ii < 8; ii++)
#pragma unroll
            for (int j = 0; j < 8; j++) imma(d[j], a[(j + ii) & 7], b[(j * 3 + ii) & 7]);
    }
    const long long t1 = clock64();
    int acc = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) acc += d[j][0] + d[j][1] + d[j][2] + d[j][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

// ALU = extra integer instructions (max.s32) issued per MMA by the same warp
template <int ALU>
__global__ void __launch_bounds__(1024) imma_kernel(int iters, int *out, long long *cycles)
{
    int d[8][4];
    uint32_t a[4], b[2];
    int q[8];
#pragma unroll
    for (int j = 0; j < 8; j++) {
        q[j] = threadIdx.x + j;
#pragma unroll
        for (int e = 0; e < 4; e++) d[j][e] = 0;
    }
    a[0] = threadIdx.x * 0x01010101u, a[1] = a[0] + 1, a[2] = a[0] + 2, a[3] = a[0] + 3, b[0] = 0x01020304u, b[1] = 0x04030201u;
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int j = 0; j < 8; j++) {
            imma(d[j], a, b);
#pragma unroll
            for (int e = 0; e < ALU; e++) asm volatile("max.s32 %0, %0, %1;" : "+r"(q[(j + e) & 7]) : "r"(i));
        }
    }
    const long long t1 = clock64();
    int acc = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) acc += d[j][0] + d[j][1] + d[j][2] + d[j][3] + q[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int ALU>
static void run(int threads, int *out, long long *cyc)
{
    const int iters = 2048;
    long long h;
    imma_kernel<ALU><<<1, threads>>>(iters, out, cyc);
    cudaDeviceSynchronize();
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    const double mmas = (threads / 32) * 8.0 * iters;
    printf("warps %2d  alu/mma %2d : %7.2f cycles per MMA per SM, %8.1f MAC/clk/SM, %6.1f alu thread-ops/clk/SM (%lld cycles)\n", threads / 32,
           ALU, h / mmas, mmas * 16 * 8 * 32 / h, mmas * ALU * 32 / h, h);
}

int main()
{
    int *out;
    long long *cyc;
    cudaMalloc(&out, 1024 * 4);
    cudaMalloc(&cyc, 8);
    for (int t : {128, 256, 512, 1024}) run<0>(t, out, cyc);
    for (int t : {256, 512, 1024}) run<4>(t, out, cyc);
    for (int t : {256, 512, 1024}) run<8>(t, out, cyc);
    for (int t : {512, 1024}) run<16>(t, out, cyc);
    for (int t : {128, 256, 512}) {
        long long h;
        imma_varied_kernel<<<1, t>>>(2048, out, cyc);
        cudaDeviceSynchronize();
        cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        const double mmas = (t / 32) * 8.0 * 2048;
        printf("warps %2d  eight A / B operand sets in rotation: %7.2f cycles per MMA per SM, %8.1f MAC/clk/SM\n", t / 32, h / mmas,
               mmas * 16 * 8 * 32 / h);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
