// dwconv3x3_umma128.cu -- int8 depthwise 3x3, stride 1, "same" padding, on the tensor cores, 128-byte pixels.
//
// Second take on csrc/dwconv3x3_umma.cu (whose no-swizzle N = 16 MMAs turned out operand-fetch bound).  Here a
// tile is a flat pixel sequence [pixel][128 channels] written by ONE SWIZZLE_128B TMA box; csrc/umma_probe.cu
// showed that a swizzled descriptor may start any number of rows into such a tile, so the A operand of tap
// (ky, kx) and 32-channel slab kq is  umma_desc_sw128(tile + (p0 + ky*(W+2) + kx) * 128) + 2*kq.  B is a diagonal
// 32 x 32 matrix per (slab, tap) kept in swizzled [32][128 B] tiles of four taps each.  Nine M128 N32 K32 MMAs
// give 128 pixels x 32 channels; a tile is 256 pixels x 128 channels (8 accumulators of 32 TMEM columns, double
// buffered = all 512 columns, one CTA per SM), drained by 16 epilogue warps (one pixel x 32 channels per thread).
// Zero-point padding through per-class seeds as in the other depthwise kernels.  Opt-in: SHL_B200_DW_UMMA=2.
#include <stdlib.h>

#include "common.cuh"

namespace b200 {

constexpr int kVStagesMax = 3;
constexpr int kVMaxMB = 2;                         // 128-pixel blocks per tile
constexpr int kVEpiWarps = 16;
constexpr int kVThreads = 96 + kVEpiWarps * 32;    // TMA warp, MMA warp, epilogue warps, second MMA warp (last)
constexpr int kVBufCols = kVMaxMB * 4 * 32;        // TMEM columns of one accumulator buffer
constexpr int kVTmemCols = 2 * kVBufCols;          // 512
constexpr int kVBBytes = 4 * 3 * 4096;             // diagonal B tiles: [slab][tap group][32 rows][128 B]

struct DwV128Args {
    int n, cp, h, w;
    int th, thi, twi, nb;
    int ybands, cchunks, ntiles;
    int box_bytes;     // bytes one TMA box delivers
    int stage_stride;  // allocation per ring slot (multiple of 1024): covers the reads of the dropped pixels too
    int stages;
    uint32_t inv_twi, inv_thi;
    uint32_t idesc;
    int diag;
    const uint32_t *wrow;  // [3 (ky)][cp] words (w[ky][0][c], w[ky][1][c], w[ky][2][c], 0)
    int8_t *out;
    int zp_in;
    EpiScalars ep;
};

template <int MODE>
__global__ void __launch_bounds__(kVThreads, 1)
dw3x3_umma128_kernel(const __grid_constant__ CUtensorMap tmap, const DwV128Args a)
{
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t *s_b = smem;                 // kVBBytes, 1024-aligned tiles
    uint8_t *ring = smem + kVBBytes;     // a.stages slots of a.stage_stride bytes
    __shared__ __align__(16) int s_seed[16 * 128];
    __shared__ __align__(16) float s_mu[128], s_ba[128];
    __shared__ uint8_t s_lut[256];
    __shared__ uint64_t full_bar[kVStagesMax], empty_bar[kVStagesMax], acc_full[2], acc_empty[2];
    __shared__ uint32_t tmem_ptr;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    pdl_launch_dependents();
    const int cc = blockIdx.x % a.cchunks;          // this CTA's 128-channel chunk, for good
    const int t0 = blockIdx.x / a.cchunks, tstep = gridDim.x / a.cchunks;

    // ---- constants of the chunk ----
    for (int i = tid; i < kVBBytes / 16; i += kVThreads) reinterpret_cast<uint4 *>(s_b)[i] = make_uint4(0, 0, 0, 0);
    __syncthreads();
    for (int i = tid; i < 128 * 9; i += kVThreads) {
        const int t = i % 9, c = i / 9;                 // tap, channel inside the chunk
        const int kq = c >> 5, nrow = c & 31;           // slab, row (= column) of the diagonal
        const uint32_t wv = __ldg(a.wrow + (t / 3) * a.cp + cc * 128 + c);
        const int tg = t >> 2, j = t & 3;               // tile of four taps, 32-byte K slice inside its rows
        const int c16 = j * 2 + (nrow >> 4);            // 16-byte chunk of the row that holds k = nrow
        s_b[(kq * 3 + tg) * 4096 + nrow * 128 + ((c16 ^ (nrow & 7)) << 4) + (nrow & 15)] =
            static_cast<uint8_t>(wv >> (8 * (t % 3)));
    }
    for (int i = tid; i < 16 * 128; i += kVThreads) {
        const int c = i & 127, cls = i >> 7;
        const int cg = cc * 128 + c;
        int padsum = 0;
#pragma unroll
        for (int ky = 0; ky < 3; ky++) {
            const uint32_t wv = __ldg(a.wrow + ky * a.cp + cg);
            const bool rowpad = (ky == 0 && (cls & 4)) || (ky == 2 && (cls & 8));
#pragma unroll
            for (int kx = 0; kx < 3; kx++) {
                const bool colpad = (kx == 0 && (cls & 1)) || (kx == 2 && (cls & 2));
                if (rowpad || colpad) padsum += static_cast<int8_t>(wv >> (8 * kx));
            }
        }
        s_seed[i] = __ldg(a.ep.ibias + cg) + kMagicI + a.zp_in * padsum;
    }
    if (tid < 128) {
        s_mu[tid] = __ldg(a.ep.mult + cc * 128 + tid);
        s_ba[tid] = __ldg(a.ep.badd + cc * 128 + tid);
    }
    if (a.ep.post_lut != nullptr && tid < 256) s_lut[tid] = static_cast<uint8_t>(a.ep.post_lut[tid]);
    if (tid == 0) {
        tma_prefetch_desc(&tmap);
        for (int i = 0; i < kVStagesMax; i++) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 2);  // both MMA issuers have consumed the slot
        }
        for (int i = 0; i < 2; i++) {
            mbar_init(&acc_full[i], 2);
            mbar_init(&acc_empty[i], kVEpiWarps);
        }
        mbar_fence_init();
    }
    if (warp == 1) {
        tmem_alloc(&tmem_ptr, kVTmemCols);
        tmem_relinquish();
    }
    fence_proxy_async_smem();  // s_b was written through the generic proxy
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_ptr;

    auto blocks_of = [&](int yb) {
        const int rows_out = min(a.th, a.h - yb * a.th);
        const int flat = ((a.nb - 1) * a.thi + rows_out) * a.twi;
        return min(kVMaxMB, (flat + 127) >> 7);
    };

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            pdl_wait();  // the input is the predecessor's output
            int stage = 0;
            uint32_t phase = 0;
            for (int t = t0; t < a.ntiles; t += tstep) {
                const int yb = t % a.ybands, img0 = (t / a.ybands) * a.nb;
                mbar_wait(&empty_bar[stage], phase ^ 1);
                mbar_expect_tx(&full_bar[stage], a.box_bytes);
                tma_load_4d(ring + static_cast<size_t>(stage) * a.stage_stride, &tmap, &full_bar[stage], cc * 128, -1,
                            yb * a.th - 1, img0);
                if (++stage == a.stages) stage = 0, phase ^= 1;
            }
        }
    } else if (warp == 1 || warp == 2 + kVEpiWarps) {
        // ===== MMA issuers: one thread gets an MMA out every ~113 cycles whatever its shape (csrc/umma_probe.cu), so
        // two warps issue, each the two 32-channel slabs (= accumulators) it owns =====
        const int kq_lo = warp == 1 ? 0 : 2;
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            uint32_t a_off[9];  // tap offsets in 16-byte units (a pixel is 8 of them)
#pragma unroll
            for (int t = 0; t < 9; t++) a_off[t] = static_cast<uint32_t>((t / 3) * a.twi + t % 3) * 8;
            const uint64_t b0 = umma_desc_sw128(smem_u32(s_b));
            for (int t = t0; t < a.ntiles; t += tstep, it++) {
                const int buf = it & 1;
                const int nmb = blocks_of(t % a.ybands);
                mbar_wait(&acc_empty[buf], ((it >> 1) & 1) ^ 1);
                mbar_wait(&full_bar[stage], phase);
                tc_fence_after();
                const uint64_t a0 = umma_desc_sw128(smem_u32(ring) + static_cast<uint32_t>(stage) * a.stage_stride);
                // tap-major: consecutive MMAs go to different accumulators (8 per tile), so the nine accumulating
                // MMAs of one accumulator are eight instructions apart and never wait for each other's latency
                const uint32_t d0 = tmem_base + buf * kVBufCols;
                const int nmb_run = (a.diag & 2) ? 0 : nmb;
#pragma unroll
                for (int tp = 0; tp < 9; tp++) {
                    for (int mb = 0; mb < nmb_run; mb++) {
#pragma unroll
                        for (int kq = kq_lo; kq < kq_lo + 2; kq++) {
                            const uint64_t adesc = a0 + (static_cast<uint32_t>(mb) * 128 * 8 + a_off[tp] + 2 * kq);
                            const uint64_t bdesc = b0 + ((kq * 3 + (tp >> 2)) * (4096 >> 4) + 2 * (tp & 3));
                            tc_mma_i8(d0 + (mb * 4 + kq) * 32, adesc, bdesc, a.idesc, tp > 0 ? 1u : 0u);
                        }
                    }
                }
                tc_commit(&empty_bar[stage]);
                tc_commit(&acc_full[buf]);
                if (++stage == a.stages) stage = 0, phase ^= 1;
            }
        }
    } else {
        // ===== epilogue: one pixel x 32 channels (two halves of 16) per thread and 128-pixel block =====
        const int ew = warp - 2;
        const int kq = ew >> 2;      // 32-channel slab
        const int q = warp & 3;      // TMEM lane quarter this warp may read
        const bool has_lut = a.ep.post_lut != nullptr;
        const int zp_m = a.ep.zp_out - kMagicI;
        const int lut_base = static_cast<int>(smem_u32(s_lut));
        int lut_lo = kMagicI - a.ep.zp_out - 128 - lut_base;
        asm("mov.b32 %0, %0;" : "+r"(lut_lo));
        const int chs = cc * 128 + kq * 32;
        const uint32_t seed0 = smem_u32(s_seed + kq * 32);
        const uint32_t mu0 = smem_u32(s_mu + kq * 32), ba0 = smem_u32(s_ba + kq * 32);
        uint32_t pix[kVMaxMB], ooff[kVMaxMB];
#pragma unroll
        for (int mb = 0; mb < kVMaxMB; mb++) {
            const uint32_t p = mb * 128 + q * 32 + lane;
            const uint32_t r = __umulhi(p, a.inv_twi);
            const uint32_t xx = p - r * a.twi;
            const uint32_t nbi = __umulhi(r, a.inv_thi);
            const uint32_t yy = r - nbi * a.thi;
            const uint32_t colc = (xx == 0 ? 1u : 0u) | (xx == static_cast<uint32_t>(a.w - 1) ? 2u : 0u);
            const bool okc = xx < static_cast<uint32_t>(a.w) && nbi < static_cast<uint32_t>(a.nb) && yy < static_cast<uint32_t>(a.th);
            pix[mb] = (nbi << 24) | (yy << 16) | (colc << 1) | (okc ? 1u : 0u);
            ooff[mb] = ((nbi * a.h + yy) * a.w + xx) * a.cp + chs;
        }
        pdl_wait();  // the output buffer may alias a tensor the predecessor still reads
        int it = 0;
        for (int t = t0; t < a.ntiles; t += tstep, it++) {
            const int buf = it & 1;
            const int yb = t % a.ybands, img0 = (t / a.ybands) * a.nb;
            const int y0 = yb * a.th;
            const int rows_out = min(a.th, a.h - y0);
            const int nmb = blocks_of(yb);
            int8_t *obase = a.out + (static_cast<size_t>(img0) * a.h + y0) * a.w * a.cp;
            mbar_wait(&acc_full[buf], (it >> 1) & 1);
            tc_fence_after();
#pragma unroll
            for (int mb = 0; mb < kVMaxMB; mb++) {
                if (mb >= ((a.diag & 4) ? 0 : nmb)) break;
                const int yy = (pix[mb] >> 16) & 0xFF, nbi = pix[mb] >> 24;
                const int y = y0 + yy;
                const bool ok = (pix[mb] & 1) && yy < rows_out && img0 + nbi < a.n;
                const int cls = ((pix[mb] >> 1) & 3) | (y == 0 ? 4 : 0) | (y == a.h - 1 ? 8 : 0);
                const uint32_t sa = seed0 + cls * (128 * 4);
                uint32_t acc32[32];
                tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + buf * kVBufCols + (mb * 4 + kq) * 32, acc32);
                tmem_ld_wait();
#pragma unroll
                for (int hf = 0; hf < 2; hf++) {
                    const uint32_t *acc = acc32 + hf * 16;
                    uint32_t o[4];
#pragma unroll
                    for (int v = 0; v < 4; v++) {
                        int s0, s1, s2, s3;
                        asm volatile("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(s0), "=r"(s1), "=r"(s2), "=r"(s3) : "r"(sa + hf * 64 + v * 16));
                        uint64_t m01, m23, b01, b23;
                        asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(m01), "=l"(m23) : "r"(mu0 + hf * 64 + v * 16));
                        asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(b01), "=l"(b23) : "r"(ba0 + hf * 64 + v * 16));
                        int tt[4];
                        requant_pair<true>(acc[4 * v] + s0, acc[4 * v + 1] + s1, m01, b01, tt[0], tt[1]);
                        requant_pair<true>(acc[4 * v + 2] + s2, acc[4 * v + 3] + s3, m23, b23, tt[2], tt[3]);
                        o[v] = finish4<MODE>(tt, a.ep, s_lut, has_lut, zp_m, lut_lo, lut_base);
                    }
                    if (ok) *reinterpret_cast<uint4 *>(obase + ooff[mb] + hf * 16) = make_uint4(o[0], o[1], o[2], o[3]);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[buf]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, kVTmemCols);
    }
}

}  // namespace b200

using namespace b200;

// 1: handled here; 0: not this kernel's case; < 0: error
int b200_dwconv3x3_umma128_launch(const b200_dwconv_desc *d, const void *wrow, void *stream, int *handled)
{
    *handled = 0;
    if (d->stride_h != 1 || d->stride_w != 1 || d->pad_top != 1 || d->pad_left != 1 || d->oh != d->h || d->ow != d->w ||
        d->cp % 128)
        return B200_OK;
    const int twi = d->w + 2;
    const int max_flat = kVMaxMB * 128;
    if (twi > 256 || twi > max_flat) return B200_OK;
    int th = max_flat / twi;
    if (th > d->h) th = d->h;
    if (th > 253) th = 253;
    int nb = 1;
    if (th == d->h) {
        nb = 1 + (max_flat - d->h * twi) / ((d->h + 2) * twi);
        if (nb > d->n) nb = d->n;
        if (nb > 255) nb = 255;
    }
    int ybands = (d->h + th - 1) / th;
    th = (d->h + ybands - 1) / ybands;
    const int cchunks = d->cp / 128;
    while (static_cast<long long>((d->n + nb - 1) / nb) * ybands * cchunks < sm_count() && (nb > 1 || th > 2)) {
        if (nb > 1)
            nb = (nb + 1) / 2;
        else {
            th = (th + 1) / 2;
            ybands = (d->h + th - 1) / th;
            th = (d->h + ybands - 1) / ybands;
        }
    }
    const int thi = th + 2;
    const long long ntiles = static_cast<long long>((d->n + nb - 1) / nb) * ybands;
    if (ntiles * cchunks >= (1ll << 31)) return B200_OK;
    if (static_cast<long long>(nb) * d->h * d->w * d->cp >= (1ll << 32)) return B200_OK;

    DwV128Args a;
    a.n = d->n, a.cp = d->cp, a.h = d->h, a.w = d->w;
    a.th = th, a.thi = thi, a.twi = twi, a.nb = nb;
    a.ybands = ybands, a.cchunks = cchunks, a.ntiles = static_cast<int>(ntiles);
    a.box_bytes = nb * thi * twi * 128;
    int need = (max_flat + 2 * twi + 2) * 128;
    if (need < a.box_bytes) need = a.box_bytes;
    a.stage_stride = (need + 1023) & ~1023;
    a.stages = static_cast<int>((200 * 1024 - kVBBytes) / a.stage_stride);
    if (a.stages > kVStagesMax) a.stages = kVStagesMax;
    if (a.stages < 2) return B200_OK;
    a.inv_twi = static_cast<uint32_t>((1ull << 32) / static_cast<uint32_t>(twi)) + 1;
    a.inv_thi = static_cast<uint32_t>((1ull << 32) / static_cast<uint32_t>(thi)) + 1;
    a.idesc = umma_idesc(2 /*S32*/, 1 /*S8*/, 128, 32);
    a.diag = getenv("SHL_B200_DW_UMMA_DIAG") ? atoi(getenv("SHL_B200_DW_UMMA_DIAG")) : 0;
    a.wrow = static_cast<const uint32_t *>(wrow);
    a.out = static_cast<int8_t *>(d->out);
    a.zp_in = d->zp_in;
    a.ep = make_epi(d->ep);

    alignas(64) CUtensorMap tm;
    int rc = encode_tmap_nhwc_u8_sw128(&tm, d->in, d->n, d->h, d->w, d->cp, twi, thi, nb);
    if (rc) return rc;

    long long cap = sm_count();
    if (cap > cchunks) cap -= cap % cchunks;
    if (cap < cchunks) cap = cchunks;
    const long long want = ntiles * cchunks;
    const int grid = static_cast<int>(want < cap ? want : cap);
    const size_t smem = static_cast<size_t>(kVBBytes) + static_cast<size_t>(a.stages) * a.stage_stride + 1024;
    int mode;
    if (d->ep.post_lut)
        mode = d->ep.act == B200_ACT_NONE ? EPI_LUT : EPI_GENERIC;
    else
        mode = d->ep.act == B200_ACT_NONE ? EPI_PLAIN : (d->ep.act == B200_ACT_RELU ? EPI_RELU : EPI_RELU6);
    int dev = 0;
    B200_CUDA_CHECK(cudaGetDevice(&dev));
    cudaStream_t s = (cudaStream_t)stream;
#define B200_DWV_CASE(M)                                                                                       \
    case M: {                                                                                                  \
        static bool attr[64] = {};                                                                             \
        if (dev >= 0 && dev < 64 && !attr[dev]) {                                                              \
            B200_CUDA_CHECK(cudaFuncSetAttribute(dw3x3_umma128_kernel<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                                 204 * 1024));                                                 \
            attr[dev] = true;                                                                                  \
        }                                                                                                      \
        B200_CUDA_CHECK(launch_kernel(dw3x3_umma128_kernel<M>, dim3(grid), dim3(kVThreads), smem, s, tm, a));  \
        break;                                                                                                 \
    }
    switch (mode) {
        B200_DWV_CASE(EPI_PLAIN)
        B200_DWV_CASE(EPI_RELU)
        B200_DWV_CASE(EPI_RELU6)
        B200_DWV_CASE(EPI_LUT)
        default:
            B200_DWV_CASE(EPI_GENERIC)
    }
#undef B200_DWV_CASE
    B200_LAUNCH_CHECK();
    *handled = 1;
    return B200_OK;
}
