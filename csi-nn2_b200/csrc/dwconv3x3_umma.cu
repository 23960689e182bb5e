// dwconv3x3_umma.cu -- int8 depthwise 3x3, stride 1, "same" padding, on the tensor cores.
//
// The dp4a kernel (dwconv3x3_tma.cu) issues ~13 instructions per output and is bound by the
// integer pipe; the requantising epilogue alone is ~5.  Here the nine taps are accumulated by
// tcgen05.mma kind::i8 and the CUDA cores only run the epilogue:
//
//   * a tile = NB images x TH output rows x the full width of 16 channels ("plane"), fetched with
//     its halo by ONE 4-D TMA box {16 ch, W+2, TH+2, NB} into shared memory as a flat pixel
//     sequence [pixel][16 B] (padding zero-filled by the hardware, zero point folded into the
//     accumulator seeds exactly as in the dp4a kernel);
//   * in that flat sequence the input of output pixel p for tap (ky, kx) is pixel
//     p + ky*(W+2) + kx: the A operand of a tap is the SAME buffer at a shifted start address.
//     A pixel is 16 bytes, eight pixels are 128 contiguous bytes = one K-major no-swizzle core
//     matrix, so a 128-pixel x 32-byte A tile of two taps (t, t') is a descriptor with start
//     p0 + off(t), SBO 128 B and LBO (off(t') - off(t)) * 16 B;
//   * B (per plane and tap pair, 16 x 32 bytes, built in shared memory from the packed depthwise
//     weights) is diagonal: B[n][k = (tap half, c)] = w[tap][c] if c == n else 0.  Five
//     M128 N16 K32 MMAs produce 128 pixels x 16 channels; outputs of the two halo columns are
//     computed and dropped;
//   * accumulators live in TMEM (two buffers of 4 pixel blocks x 2 planes x 16 columns); eight
//     epilogue warps read them back (one pixel x 16 channels per thread), add the seed of the
//     pixel's border class, requantise with the contract of include/b200nn.h and store 16 bytes.
//
// Replaces shl_rvv_dwconv3x3s1_int8 (source/thead_rvv/int8/depthwise_convolution_3x3_int8.c:31);
// semantics shl_ref_depthwise_conv2d_quant (source/reference/convolution.c:416).
#include <stdlib.h>

#include "common.cuh"

namespace b200 {

constexpr int kUStages = 3;                        // shared-memory ring slots (2 planes each)
constexpr int kUMaxMB = 4;                         // 128-pixel blocks per tile
constexpr int kUEpiWarps = 8;
constexpr int kUThreads = 64 + kUEpiWarps * 32;    // TMA warp, MMA warp, epilogue warps
constexpr int kUBufCols = kUMaxMB * 2 * 16;        // TMEM columns of one accumulator buffer
constexpr int kUTmemCols = 2 * kUBufCols;          // 256: two CTAs per SM fit the 512 columns

struct DwUmmaArgs {
    int n, cp, h, w;
    int th, thi, twi, nb;  // tile: nb images x th output rows; thi = th + 2, twi = w + 2
    int ybands, cchunks, ntiles;
    int plane_bytes;   // bytes one TMA box delivers
    int plane_stride;  // allocation per plane: covers the reads of the dropped pixels too
    uint32_t inv_twi, inv_thi;  // floor(2^32 / d) + 1: exact quotients for the small flat indices
    uint32_t idesc;
    int swap_lbo_sbo;  // diagnostic: exchange the two descriptor strides
    int diag;          // diagnostic (wrong results): 1 = no TMA after the ring's first fill, 2 = no MMAs, 4 = no epilogue math / stores
    const uint32_t *wrow;  // [3 (ky)][cp] words (w[ky][0][c], w[ky][1][c], w[ky][2][c], 0)
    int8_t *out;
    int zp_in;
    EpiScalars ep;
};

// K-major, no swizzle: rows of a core matrix 16 bytes apart, 8-row groups SBO apart, the two
// 16-byte K halves LBO apart
__device__ __forceinline__ uint64_t umma_desc_nosw(uint32_t smem_addr, uint32_t lbo, uint32_t sbo)
{
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
    d |= static_cast<uint64_t>((lbo >> 4) & 0x3FFF) << 16;
    d |= static_cast<uint64_t>((sbo >> 4) & 0x3FFF) << 32;
    d |= static_cast<uint64_t>(1) << 46;  // descriptor version (sm_100); layout bits 61..63 = 0: no swizzle
    return d;
}

template <int MODE>
__global__ void __launch_bounds__(kUThreads, 2)
dw3x3_umma_kernel(const __grid_constant__ CUtensorMap tmap, const DwUmmaArgs a)
{
    extern __shared__ __align__(128) uint8_t smem_raw[];
    uint8_t *smem = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);
    __shared__ __align__(128) uint8_t s_b[2 * 5 * 512];
    __shared__ __align__(16) int s_seed[16 * 32];
    __shared__ uint8_t s_lut[256];
    __shared__ uint64_t full_bar[kUStages], empty_bar[kUStages], acc_full[2], acc_empty[2];
    __shared__ uint32_t tmem_ptr;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    pdl_launch_dependents();
    const int cc = blockIdx.x % a.cchunks;          // this CTA's 32-channel chunk, for good
    const int t0 = blockIdx.x / a.cchunks, tstep = gridDim.x / a.cchunks;

    // ---- constants of the chunk: diagonal B matrices, seeds, post table ----
    for (int i = tid; i < 2 * 5 * 512 / 4; i += kUThreads) reinterpret_cast<uint32_t *>(s_b)[i] = 0;
    __syncthreads();
    for (int i = tid; i < 2 * 16 * 9; i += kUThreads) {
        const int t = i % 9, c = (i / 9) % 16, pl = i / 144;
        const int cg = cc * 32 + pl * 16 + c;
        if (cg < a.cp) {
            const uint32_t wv = __ldg(a.wrow + (t / 3) * a.cp + cg);
            const int j = t >> 1, hf = t & 1;
            s_b[(pl * 5 + j) * 512 + hf * 256 + (c >> 3) * 128 + (c & 7) * 16 + c] = static_cast<uint8_t>(wv >> (8 * (t % 3)));
        }
    }
    for (int i = tid; i < 16 * 32; i += kUThreads) {
        const int c = i % 32, cls = i / 32;
        const int cg = cc * 32 + c;
        int v = 0;
        if (cg < a.cp) {
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
            v = __ldg(a.ep.ibias + cg) + kMagicI + a.zp_in * padsum;
        }
        s_seed[i] = v;
    }
    if (a.ep.post_lut != nullptr && tid < 256) s_lut[tid] = static_cast<uint8_t>(a.ep.post_lut[tid]);
    if (tid == 0) {
        tma_prefetch_desc(&tmap);
        for (int i = 0; i < kUStages; i++) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < 2; i++) {
            mbar_init(&acc_full[i], 1);
            mbar_init(&acc_empty[i], kUEpiWarps);
        }
        mbar_fence_init();
    }
    if (warp == 1) {
        tmem_alloc(&tmem_ptr, kUTmemCols);
        tmem_relinquish();
    }
    fence_proxy_async_smem();  // s_b was written through the generic proxy
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_ptr;

    // tile t of this chunk -> (first image, row band); flat pixels the tile has to produce
    auto blocks_of = [&](int yb) {
        const int rows_out = min(a.th, a.h - yb * a.th);
        const int flat = ((a.nb - 1) * a.thi + rows_out) * a.twi;
        return min(kUMaxMB, (flat + 127) >> 7);
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
                if ((a.diag & 1) && phase) {
                    mbar_arrive(&full_bar[stage]);
                    if (++stage == kUStages) stage = 0, phase ^= 1;
                    continue;
                }
                mbar_expect_tx(&full_bar[stage], 2 * a.plane_bytes);
                uint8_t *dst = smem + static_cast<size_t>(stage) * 2 * a.plane_stride;
                tma_load_4d(dst, &tmap, &full_bar[stage], cc * 32, -1, yb * a.th - 1, img0);
                tma_load_4d(dst + a.plane_stride, &tmap, &full_bar[stage], cc * 32 + 16, -1, yb * a.th - 1, img0);
                if (++stage == kUStages) stage = 0, phase ^= 1;
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            // descriptors are constants up to the A start address: the upper word (SBO, version) and the
            // LBO field of each tap pair once, per MMA one add into the 14-bit address field
            uint64_t a_tmpl[5], b_desc[2][5];
            uint32_t a_off[5];
#pragma unroll
            for (int j = 0; j < 5; j++) {
                const int ta = 2 * j, tb = 2 * j + 1;  // taps of this K = 32 step: flat offsets (ky * twi + kx) pixels
                const int offa = (ta / 3) * a.twi + ta % 3;
                const int offb = j < 4 ? (tb / 3) * a.twi + tb % 3 : offa + 1;
                const uint32_t lbo = static_cast<uint32_t>(offb - offa) * 16;
                a_tmpl[j] = a.swap_lbo_sbo ? umma_desc_nosw(0, 128, lbo) : umma_desc_nosw(0, lbo, 128);
                a_off[j] = static_cast<uint32_t>(offa);
#pragma unroll
                for (int pl = 0; pl < 2; pl++)
                    b_desc[pl][j] = a.swap_lbo_sbo ? umma_desc_nosw(smem_u32(s_b) + (pl * 5 + j) * 512, 128, 256)
                                                   : umma_desc_nosw(smem_u32(s_b) + (pl * 5 + j) * 512, 256, 128);
            }
            const uint32_t plane16 = static_cast<uint32_t>(a.plane_stride) >> 4;
            for (int t = t0; t < a.ntiles; t += tstep, it++) {
                const int buf = it & 1;
                const int nmb = blocks_of(t % a.ybands);
                mbar_wait(&acc_empty[buf], ((it >> 1) & 1) ^ 1);
                mbar_wait(&full_bar[stage], phase);
                tc_fence_after();
                // start addresses in 16-byte units (= pixels); the whole ring lies below 256 KB
                const uint32_t tile16 = (smem_u32(smem) + static_cast<uint32_t>(stage) * 2 * a.plane_stride) >> 4;
                // tap-pair-major: consecutive MMAs go to different accumulators, so the accumulating MMAs of one
                // accumulator never wait for each other's latency
                const uint32_t d0 = tmem_base + buf * kUBufCols;
                const int nmb_run = (a.diag & 2) ? 0 : nmb;
#pragma unroll
                for (int j = 0; j < 5; j++) {
                    for (int mb = 0; mb < nmb_run; mb++) {
#pragma unroll
                        for (int pl = 0; pl < 2; pl++)
                            tc_mma_i8(d0 + (mb * 2 + pl) * 16, a_tmpl[j] + (tile16 + pl * plane16 + mb * 128 + a_off[j]),
                                      b_desc[pl][j], a.idesc, j > 0 ? 1u : 0u);
                    }
                }
                tc_commit(&empty_bar[stage]);  // the slot is free once these MMAs have read it
                tc_commit(&acc_full[buf]);
                if (++stage == kUStages) stage = 0, phase ^= 1;
            }
        }
    } else {
        // ===== epilogue: one pixel x 16 channels per thread and 128-pixel block =====
        const int ew = warp - 2;
        const int g = ew >> 2;       // plane
        const int q = warp & 3;      // TMEM lane quarter this warp may read
        const bool has_lut = a.ep.post_lut != nullptr;
        const int zp_m = a.ep.zp_out - kMagicI;
        const int lut_base = static_cast<int>(smem_u32(s_lut));
        int lut_lo = kMagicI - a.ep.zp_out - 128 - lut_base;
        asm("mov.b32 %0, %0;" : "+r"(lut_lo));
        const int chs = cc * 32 + g * 16;
        const bool ch_ok = chs < a.cp;
        uint64_t mu[8], ba[8];
        {
            const int c0 = ch_ok ? chs : 0;
#pragma unroll
            for (int v = 0; v < 4; v++) {
                const float4 m4 = __ldg(reinterpret_cast<const float4 *>(a.ep.mult + c0) + v);
                const float4 b4 = __ldg(reinterpret_cast<const float4 *>(a.ep.badd + c0) + v);
                mu[2 * v] = f2_pack(m4.x, m4.y), mu[2 * v + 1] = f2_pack(m4.z, m4.w);
                ba[2 * v] = f2_pack(b4.x, b4.y), ba[2 * v + 1] = f2_pack(b4.z, b4.w);
            }
        }
        const uint32_t seed0 = smem_u32(s_seed + g * 16);
        // the pixel a thread owns in 128-pixel block mb does not depend on the tile: decode once.
        // pix[mb] = image in the tile (8 bits) | tile row (8) | column class (2) | column valid (1); ooff[mb] = its
        // byte offset inside the tile's output (host-checked < 2^32)
        uint32_t pix[kUMaxMB], ooff[kUMaxMB];
#pragma unroll
        for (int mb = 0; mb < kUMaxMB; mb++) {
            const uint32_t p = mb * 128 + q * 32 + lane;
            const uint32_t r = __umulhi(p, a.inv_twi);
            const uint32_t xx = p - r * a.twi;
            const uint32_t nbi = __umulhi(r, a.inv_thi);
            const uint32_t yy = r - nbi * a.thi;
            const uint32_t colc = (xx == 0 ? 1u : 0u) | (xx == static_cast<uint32_t>(a.w - 1) ? 2u : 0u);
            const bool okc = ch_ok && xx < static_cast<uint32_t>(a.w) && nbi < static_cast<uint32_t>(a.nb) && yy < static_cast<uint32_t>(a.th);
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
            for (int mb = 0; mb < kUMaxMB; mb++) {
                if (mb >= ((a.diag & 4) ? 0 : nmb)) break;
                uint32_t acc[16];
                tmem_ld_32x16(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + buf * kUBufCols + (mb * 2 + g) * 16, acc);
                const int yy = (pix[mb] >> 16) & 0xFF, nbi = pix[mb] >> 24;
                const int y = y0 + yy;
                const bool ok = (pix[mb] & 1) && yy < rows_out && img0 + nbi < a.n;
                const int cls = ((pix[mb] >> 1) & 3) | (y == 0 ? 4 : 0) | (y == a.h - 1 ? 8 : 0);
                const uint32_t sa = seed0 + cls * (32 * 4);
                tmem_ld_wait();
                uint32_t o[4];
#pragma unroll
                for (int v = 0; v < 4; v++) {
                    int s0, s1, s2, s3;
                    asm volatile("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(s0), "=r"(s1), "=r"(s2), "=r"(s3) : "r"(sa + v * 16));
                    int tt[4];
                    requant_pair<true>(acc[4 * v] + s0, acc[4 * v + 1] + s1, mu[2 * v], ba[2 * v], tt[0], tt[1]);
                    requant_pair<true>(acc[4 * v + 2] + s2, acc[4 * v + 3] + s3, mu[2 * v + 1], ba[2 * v + 1], tt[2], tt[3]);
                    o[v] = finish4<MODE>(tt, a.ep, s_lut, has_lut, zp_m, lut_lo, lut_base);
                }
                if (ok) *reinterpret_cast<uint4 *>(obase + ooff[mb]) = make_uint4(o[0], o[1], o[2], o[3]);
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
        tmem_dealloc(tmem_base, kUTmemCols);
    }
}

}  // namespace b200

using namespace b200;

// 1: handled here; 0: not this kernel's case (the caller goes on to the dp4a kernel); < 0: error
int b200_dwconv3x3_umma_launch(const b200_dwconv_desc *d, const void *wrow, void *stream, int *handled)
{
    *handled = 0;
    const char *e_on = getenv("SHL_B200_DW_UMMA"), *e_swap = getenv("SHL_B200_DW_UMMA_SWAP");
    const int enabled = e_on ? atoi(e_on) == 1 : 0, swap = e_swap ? atoi(e_swap) : 0;
    if (!enabled) return B200_OK;
    if (d->stride_h != 1 || d->stride_w != 1 || d->pad_top != 1 || d->pad_left != 1 || d->oh != d->h || d->ow != d->w)
        return B200_OK;
    const int twi = d->w + 2;
    const int max_flat = kUMaxMB * 128;
    if (twi > 256 || twi > max_flat) return B200_OK;
    int th = max_flat / twi;
    if (th > d->h) th = d->h;
    if (th + 2 > 256) th = 254;
    int nb = 1;
    if (th == d->h) {
        nb = 1 + (max_flat - d->h * twi) / ((d->h + 2) * twi);
        if (nb > d->n) nb = d->n;
        if (nb > 256) nb = 256;
    }
    int ybands = (d->h + th - 1) / th;
    th = (d->h + ybands - 1) / ybands;
    const int cchunks = (d->cp + 31) / 32;
    // small batches: shorter tiles until every SM has one
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
    if (static_cast<long long>(nb) * d->h * d->w * d->cp >= (1ll << 32) || nb > 255 || th > 255) return B200_OK;

    DwUmmaArgs a;
    a.n = d->n, a.cp = d->cp, a.h = d->h, a.w = d->w;
    a.th = th, a.thi = thi, a.twi = twi, a.nb = nb;
    a.ybands = ybands, a.cchunks = cchunks, a.ntiles = static_cast<int>(ntiles);
    a.plane_bytes = nb * thi * twi * 16;
    int need = (max_flat + 2 * twi + 2) * 16;
    if (need < a.plane_bytes) need = a.plane_bytes;
    a.plane_stride = (need + 127) & ~127;
    a.inv_twi = static_cast<uint32_t>((1ull << 32) / static_cast<uint32_t>(twi)) + 1;
    a.inv_thi = static_cast<uint32_t>((1ull << 32) / static_cast<uint32_t>(thi)) + 1;
    a.idesc = umma_idesc(2 /*S32*/, 1 /*S8*/, 128, 16);
    a.swap_lbo_sbo = swap;
    a.diag = getenv("SHL_B200_DW_UMMA_DIAG") ? atoi(getenv("SHL_B200_DW_UMMA_DIAG")) : 0;
    a.wrow = static_cast<const uint32_t *>(wrow);
    a.out = static_cast<int8_t *>(d->out);
    a.zp_in = d->zp_in;
    a.ep = make_epi(d->ep);

    alignas(64) CUtensorMap tm;
    int rc = encode_tmap_nhwc_u8_nb(&tm, d->in, d->n, d->h, d->w, d->cp, 16, twi, thi, nb);
    if (rc) return rc;

    long long cap = static_cast<long long>(sm_count()) * 2;
    if (cap > cchunks) cap -= cap % cchunks;
    if (cap < cchunks) cap = cchunks;
    const long long want = ntiles * cchunks;
    const int grid = static_cast<int>(want < cap ? want : cap);
    const size_t smem = static_cast<size_t>(kUStages) * 2 * a.plane_stride + 128;
    int mode;
    if (d->ep.post_lut)
        mode = d->ep.act == B200_ACT_NONE ? EPI_LUT : EPI_GENERIC;
    else
        mode = d->ep.act == B200_ACT_NONE ? EPI_PLAIN : (d->ep.act == B200_ACT_RELU ? EPI_RELU : EPI_RELU6);
    int dev = 0;
    B200_CUDA_CHECK(cudaGetDevice(&dev));
    cudaStream_t s = (cudaStream_t)stream;
#define B200_DWU_CASE(M)                                                                                     \
    case M: {                                                                                                \
        static bool attr[64] = {};                                                                           \
        if (dev >= 0 && dev < 64 && !attr[dev]) {                                                            \
            B200_CUDA_CHECK(cudaFuncSetAttribute(dw3x3_umma_kernel<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                                 100 * 1024));                                               \
            attr[dev] = true;                                                                                \
        }                                                                                                    \
        B200_CUDA_CHECK(launch_kernel(dw3x3_umma_kernel<M>, dim3(grid), dim3(kUThreads), smem, s, tm, a));   \
        break;                                                                                               \
    }
    if (smem > 100 * 1024) return B200_OK;
    switch (mode) {
        B200_DWU_CASE(EPI_PLAIN)
        B200_DWU_CASE(EPI_RELU)
        B200_DWU_CASE(EPI_RELU6)
        B200_DWU_CASE(EPI_LUT)
        default:
            B200_DWU_CASE(EPI_GENERIC)
    }
#undef B200_DWU_CASE
    B200_LAUNCH_CHECK();
    *handled = 1;
    return B200_OK;
}
