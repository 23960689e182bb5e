// umma_probe.cu -- test hook: does a SWIZZLE_128B K-major shared-memory descriptor accept a start address
// that is shifted by whole 128-byte rows which are NOT a multiple of the 8-row swizzle period?
//
// Why it matters: an implicit-GEMM 3x3 convolution (and a tensor-core depthwise kernel with a 128-byte pixel
// pitch) can read the A operand of tap (ky, kx) straight out of ONE halo tile in shared memory -- a flat pixel
// sequence [pixel][128 B] written by a swizzled TMA load -- by starting the descriptor (ky * row + kx) pixels
// further on.  That only works if the tensor core applies the swizzle XOR to the absolute shared-memory
// address (as TMA does when it writes), not to the row index relative to the descriptor's start.
// csrc/dwconv3x3_umma.cu proved the shifted-start idea for the no-swizzle layout; this probe answers it for
// SWIZZLE_128B.  One M128 N32 K32 kind::i8 MMA: D[m][n] = sum_k A[shift + m][k0*32 + k] * B[n][k0*32 + k].
#include "common.cuh"

namespace b200 {

constexpr int kProbeRows = 144;  // 128 + the largest shift

__global__ void __launch_bounds__(128) umma_shift_probe_kernel(const __grid_constant__ CUtensorMap tmap_a,
                                                               const __grid_constant__ CUtensorMap tmap_b, int shift,
                                                               int k0, int32_t *out)
{
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t *sa = smem;                       // [kProbeRows][128 B], swizzled by the TMA
    uint8_t *sb = smem + kProbeRows * 128 + ((1024 - (kProbeRows * 128) % 1024) % 1024);  // [32][128 B]
    __shared__ uint64_t full_bar, done_bar;
    __shared__ uint32_t tmem_ptr;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) {
        mbar_init(&full_bar, 1);
        mbar_init(&done_bar, 1);
        mbar_fence_init();
    }
    if (warp == 0) {
        tmem_alloc(&tmem_ptr, 32);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_ptr;
    if (tid == 0) {
        mbar_expect_tx(&full_bar, kProbeRows * 128 + 32 * 128);
        tma_load_2d(sa, &tmap_a, &full_bar, 0, 0);
        tma_load_2d(sb, &tmap_b, &full_bar, 0, 0);
        mbar_wait(&full_bar, 0);
        tc_fence_after();
        const uint64_t adesc = umma_desc_sw128(smem_u32(sa) + shift * 128) + 2 * k0;
        const uint64_t bdesc = umma_desc_sw128(smem_u32(sb)) + 2 * k0;
        tc_mma_i8(tmem_base, adesc, bdesc, umma_idesc(2 /*S32*/, 1 /*S8*/, 128, 32), 0u);
        tc_commit(&done_bar);
    }
    mbar_wait(&done_bar, 0);
    tc_fence_after();
    uint32_t r[32];
    tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(warp * 32) << 16), r);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 32; j++) out[tid * 32 + j] = static_cast<int32_t>(r[j]);
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 32);
    }
}

// rate probe: `reps` M128 x N x K32 kind::i8 MMAs over uninitialised swizzled operand tiles, round-robin over
// `nacc` accumulators (1 = one dependent chain); cycles from the first issue to the completion of the last
__global__ void __launch_bounds__(128) umma_rate_probe_kernel(int n, int nacc, int reps, long long *cycles, int m, int f16, int issuers)
{
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t *sa = smem;               // [128][128 B]
    uint8_t *sb = smem + 128 * 128;   // [256][128 B]
    __shared__ uint64_t done_bar;
    __shared__ uint32_t tmem_ptr;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < (128 + 256) * 128 / 16; i += 128) reinterpret_cast<uint4 *>(smem)[i] = make_uint4(i, 1, 2, 3);
    if (tid == 0) {
        mbar_init(&done_bar, issuers);
        mbar_fence_init();
    }
    if (warp == 0) {
        tmem_alloc(&tmem_ptr, 512);
        tmem_relinquish();
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_ptr;
    // `issuers` warps (lane 0 of warps 0 .. issuers-1) issue reps / issuers MMAs each, into disjoint accumulators
    if ((tid & 31) == 0 && warp < issuers) {
        const uint64_t a0 = umma_desc_sw128(smem_u32(sa)), b0 = umma_desc_sw128(smem_u32(sb));
        const uint32_t idesc = f16 ? umma_idesc(1, 0, m, n) : umma_idesc(2, 1, m, n);
        const int per = nacc / issuers > 0 ? nacc / issuers : 1;
        const uint32_t tb = tmem_base + (issuers > 1 ? warp * per * n : 0);
        const int mine = reps / issuers;
        const long long t0 = clock64();
        if (f16)
            for (int i = 0; i < mine; i++) tc_mma_f16(tb + (i % per) * n, a0 + 2 * (i & 3), b0 + 2 * (i & 3), idesc, 1u);
        else
            for (int i = 0; i < mine; i++) tc_mma_i8(tb + (i % per) * n, a0 + 2 * (i & 3), b0 + 2 * (i & 3), idesc, 1u);
        tc_commit(&done_bar);
        mbar_wait(&done_bar, 0);
        if (warp == 0) *cycles = clock64() - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

}  // namespace b200

using namespace b200;

// a_dev: [144][128] int8, b_dev: [32][128] int8, out_dev: [128][32] int32
extern "C" int b200_test_umma_shifted_start(const void *a_dev, const void *b_dev, int shift, int k0, void *out_dev,
                                            void *stream)
{
    if (!a_dev || !b_dev || !out_dev || shift < 0 || shift > kProbeRows - 128 || k0 < 0 || k0 > 3) {
        set_error("b200_test_umma_shifted_start: bad arguments (shift %d, k0 %d)", shift, k0);
        return B200_ERR_ARG;
    }
    alignas(64) CUtensorMap ta, tb;
    int rc = encode_tmap_2d(&ta, 1, a_dev, 128, kProbeRows, 128, 128, kProbeRows, 128);
    if (rc) return rc;
    rc = encode_tmap_2d(&tb, 1, b_dev, 128, 32, 128, 128, 32, 128);
    if (rc) return rc;
    const size_t smem = kProbeRows * 128 + 1024 + 32 * 128 + 1024;
    B200_CUDA_CHECK(cudaFuncSetAttribute(umma_shift_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    B200_CUDA_CHECK(launch_kernel(umma_shift_probe_kernel, dim3(1), dim3(128), smem, (cudaStream_t)stream, ta, tb, shift, k0,
                                  static_cast<int32_t *>(out_dev)));
    B200_LAUNCH_CHECK();
    return B200_OK;
}

// cycles for `reps` back-to-back M128 x n x K32 int8 MMAs on one SM, `nacc` accumulators in rotation
extern "C" int b200_test_umma_rate(int n, int nacc, int reps, long long *cycles_host, void *stream)
{
    return b200_test_umma_rate2(128, n, 0, nacc, reps, cycles_host, stream);
}

// the same for M = 64 / 128 and kind::i8 (f16 = 0, K = 32 bytes) / kind::f16 (f16 = 1, K = 16 halves)
extern "C" int b200_test_umma_rate2(int m, int n, int f16, int nacc, int reps, long long *cycles_host, void *stream)
{
    return b200_test_umma_rate3(m, n, f16, nacc, reps, 1, cycles_host, stream);
}

// ... issued by `issuers` (1, 2 or 4) warps, each into its own accumulators
extern "C" int b200_test_umma_rate3(int m, int n, int f16, int nacc, int reps, int issuers, long long *cycles_host,
                                    void *stream)
{
    if (issuers != 1 && issuers != 2 && issuers != 4) {
        set_error("b200_test_umma_rate3: issuers %d", issuers);
        return B200_ERR_ARG;
    }
    if ((m != 64 && m != 128) || n < 16 || n > 256 || n % 16 || nacc < 1 || nacc * n > 512 || reps < 1 || !cycles_host) {
        set_error("b200_test_umma_rate: bad arguments (n %d, nacc %d, reps %d)", n, nacc, reps);
        return B200_ERR_ARG;
    }
    long long *d = nullptr;
    B200_CUDA_CHECK(cudaMalloc(&d, sizeof(long long)));
    B200_CUDA_CHECK(cudaFuncSetAttribute(umma_rate_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    B200_CUDA_CHECK(launch_kernel(umma_rate_probe_kernel, dim3(1), dim3(128), (128 + 256) * 128 + 1024, (cudaStream_t)stream, n,
                                  nacc, reps, d, m, f16, issuers));
    B200_CUDA_CHECK(cudaStreamSynchronize((cudaStream_t)stream));
    B200_CUDA_CHECK(cudaMemcpy(cycles_host, d, sizeof(long long), cudaMemcpyDeviceToHost));
    cudaFree(d);
    return B200_OK;
}


// ---- TMA im2col-mode probe ---------------------------------------------------------------------------------
// One im2col load of `pixels` (<= 256) output pixels x `chans` channels of one filter tap, no swizzle, copied
// out as it lies in shared memory: pins the coordinate / corner / offset conventions the implicit-GEMM
// convolution (csrc/conv_igemm.cu) is built on against a numpy gather (tests/test_gpu_igemm.py).
namespace b200 {
__global__ void __launch_bounds__(128) tma_im2col_probe_kernel(const __grid_constant__ CUtensorMap tmap, int c0, int w0,
                                                               int h0, int n0, int woff, int hoff, int bytes, uint8_t *out)
{
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    __shared__ uint64_t bar;
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_expect_tx(&bar, bytes);
        tma_load_im2col_4d(smem, &tmap, &bar, c0, w0, h0, n0, static_cast<uint16_t>(woff), static_cast<uint16_t>(hoff));
    }
    mbar_wait(&bar, 0);
    for (int i = threadIdx.x; i < bytes; i += blockDim.x) out[i] = smem[i];
}
}  // namespace b200

extern "C" int b200_test_tma_im2col(const void *in_dev, int n, int h, int w, int c, int cp, int lower_w, int lower_h,
                                    int upper_w, int upper_h, int stride_w, int stride_h, int chans, int pixels, int c0, int w0,
                                    int h0, int n0, int woff, int hoff, void *out_dev, void *stream)
{
    using namespace b200;
    alignas(64) CUtensorMap tm;
    int rc = encode_tmap_im2col_u8(&tm, in_dev, n, h, w, c, cp, lower_w, lower_h, upper_w, upper_h, stride_w, stride_h, chans,
                                   pixels, 0);
    if (rc) return rc;
    const int bytes = chans * pixels;
    B200_CUDA_CHECK(cudaFuncSetAttribute(tma_im2col_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024));
    tma_im2col_probe_kernel<<<1, 128, bytes + 1024, (cudaStream_t)stream>>>(tm, c0, w0, h0, n0, woff, hoff, bytes,
                                                                           static_cast<uint8_t *>(out_dev));
    B200_LAUNCH_CHECK();
    return B200_OK;
}
