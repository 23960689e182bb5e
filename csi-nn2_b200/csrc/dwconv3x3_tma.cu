// dwconv3x3_tma.cu -- int8 depthwise 3x3 (stride 1 / 2, dilation 1) on pixel-major tensors,
// TMA-fed.  The HBM-bound half of a MobileNet block.
//
// A CTA walks tiles of TH x TW output pixels x CC channels.  A producer warp fetches each tile's
// input halo ((S*(TH-1)+3) x (S*(TW-1)+3) x CC bytes) with ONE 4-D TMA box into a 3-deep shared
// memory ring (negative / overhanging coordinates are zero-filled by the hardware and patched to
// the input zero point, which is what a padded tap holds in the quantised domain); 256 consumer
// threads never touch global memory for input, so HBM latency is hidden by the ring instead of by
// registers and occupancy.
//
// A consumer thread owns four channels (one 32-bit word) of ONE output column and slides down the
// tile: per input row it reads the three horizontally adjacent words, permutes them into four
// "tap words" (x[-1][c], x[0][c], x[+1][c], -) and issues dp4a against weight words
// (w[ky][0][c], w[ky][1][c], w[ky][2][c], 0): 3 dp4a per output, no unpacking.  Each input row
// feeds the three output rows it belongs to through rotating accumulator sets (12 registers), so
// the kernel needs ~60 registers.  Accumulators start at ibias + kMagicI (see common.cuh); the
// epilogue is the contract of include/b200nn.h, specialised at compile time like the GEMM's.
//
// Replaces shl_rvv_dwconv3x3s1_int8 / shl_rvv_dwconv3x3s2_int8
// (source/thead_rvv/int8/depthwise_convolution_3x3_int8.c:31,244); semantics
// shl_ref_depthwise_conv2d_quant (source/reference/convolution.c:416).
#include <stdlib.h>

#include "common.cuh"

namespace b200 {

enum { DW_PLAIN = EPI_PLAIN, DW_RELU = EPI_RELU, DW_RELU6 = EPI_RELU6, DW_LUT = EPI_LUT, DW_GENERIC = EPI_GENERIC };
constexpr int kDwStages = 3;
// consumer threads per CTA: 256 or 224 (tile widths 32/16/8 or 28/14/7 columns -- the second family
// divides the 112 / 56 / 28 / 14 / 7 wide maps of the ImageNet networks without idle columns)

struct DwTmaArgs {
    int n, c, cp, h, w, oh, ow, pt, pl;
    int th, thi;            // output rows per tile, input rows per tile
    int ybands, xbands, cchunks;
    int dxb, dyb, db;  // the grid's stride over spatial tiles, gridDim.x / cchunks, as (x band, y band, image) digits
    int stage_bytes;   // bytes one TMA box delivers
    int stage_stride;  // distance between ring slots: stage_bytes rounded up to 128 (TMA destination alignment)
    const uint32_t *wrow;   // [3 (ky)][cp] words: (w[ky][0][c], w[ky][1][c], w[ky][2][c], 0)
    int8_t *out;
    int zp_in;
    EpiScalars ep;
};

template <int MODE>
__device__ __forceinline__ uint32_t dw_requant4(const int (&acc)[4], const uint64_t (&mu2)[2], const uint64_t (&ba2)[2],
                                                const EpiScalars &ep, const uint8_t *lut, bool has_lut, int zp_m,
                                                int lut_lo, int lut_base)
{
    int t[4];
    // packed f32x2: magic -> float, fma, round (3 issue slots per two outputs)
    requant_pair<true>(acc[0], acc[1], mu2[0], ba2[0], t[0], t[1]);
    requant_pair<true>(acc[2], acc[3], mu2[1], ba2[1], t[2], t[3]);
    return finish4<MODE>(t, ep, lut, has_lut, zp_m, lut_lo, lut_base);  // DW_* == EPI_* (common.cuh)
}

// tap words of four channels from three horizontally adjacent input words
__device__ __forceinline__ void taps3(uint32_t a, uint32_t b, uint32_t c, uint32_t (&v)[4])
{
    const uint32_t lo = __byte_perm(a, b, 0x5140);  // (a.c0, b.c0, a.c1, b.c1)
    const uint32_t hi = __byte_perm(a, b, 0x7362);  // (a.c2, b.c2, a.c3, b.c3)
    v[0] = __byte_perm(lo, c, 0x4410);
    v[1] = __byte_perm(lo, c, 0x5532);
    v[2] = __byte_perm(hi, c, 0x6610);
    v[3] = __byte_perm(hi, c, 0x7732);
}

__device__ __forceinline__ void dp4(int (&acc)[4], const uint32_t (&v)[4], const uint32_t (&w)[4])
{
#pragma unroll
    for (int e = 0; e < 4; e++) acc[e] = __dp4a(static_cast<int>(v[e]), static_cast<int>(w[e]), acc[e]);
}

// This CTA's walk over the tile space (image, y band, x band, channel chunk): the chunk is
// blockIdx.x % cchunks for good; the spatial tile advances by gridDim.x / cchunks per step, applied as
// mixed-radix digits with carries -- no division in the loop (the divisions were ~60 instructions
// per tile and thread, a third of the work on the 7x7 and 14x14 maps).
struct TileWalk {
    int cc, xb, yb, b;
    __device__ __forceinline__ explicit TileWalk(const DwTmaArgs &a)
    {
        cc = blockIdx.x % a.cchunks;
        uint32_t q = blockIdx.x / a.cchunks;
        xb = q % a.xbands;
        q /= a.xbands;
        yb = q % a.ybands;
        b = q / a.ybands;
    }
    __device__ __forceinline__ void next(const DwTmaArgs &a)
    {
        xb += a.dxb;
        if (xb >= a.xbands) xb -= a.xbands, yb++;
        yb += a.dyb;
        if (yb >= a.ybands) yb -= a.ybands, b++;
        b += a.db;
    }
};

template <int S, int CC, int TW, int MODE>
__global__ void __launch_bounds__(CC / 4 * TW + 32, 2)
dw3x3_tma_kernel(const __grid_constant__ CUtensorMap tmap, const DwTmaArgs a)
{
    constexpr int TWI = S * (TW - 1) + 3;
    constexpr int WORDS = CC / 4;
    constexpr int kDwConsumers = WORDS * TW;
    static_assert(kDwConsumers % 32 == 0 && kDwConsumers <= 256, "tile shape must give whole consumer warps");
    extern __shared__ __align__(128) uint8_t smem_raw[];
    // TMA destinations must be 128-byte aligned; stay on the shared pointer (LDS / STS codegen)
    uint8_t *smem = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);
    __shared__ uint64_t full_bar[kDwStages], empty_bar[kDwStages];
    __shared__ uint8_t s_lut[256];
    // accumulator seeds [row class][column class][CC]: ibias + kMagicI + zp_in * (sum of the weights
    // of the taps that fall into the padding for that class).  The TMA zero-fills padded taps where
    // the contract wants zp_in, so the difference -- zp_in * w summed over the padded taps -- is added
    // through the seed instead of patching the tile in shared memory (no CTA barrier, no extra pass).
    // class bits: row 1 = top row padded (ky 0), 2 = bottom (ky 2); column 1 = left (kx 0), 2 = right (kx 2)
    __shared__ __align__(16) int s_seed[16 * CC];

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    pdl_launch_dependents();  // see launch_kernel (common.cuh)
    if (tid == 0) {
        tma_prefetch_desc(&tmap);
        for (int i = 0; i < kDwStages; i++) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], kDwConsumers / 32);
        }
        mbar_fence_init();
    }
    if (a.ep.post_lut != nullptr && tid < 256) s_lut[tid] = static_cast<uint8_t>(a.ep.post_lut[tid]);
    __syncthreads();


    if (warp == kDwConsumers / 32) {
        // ===== TMA producer =====
        if (elect_one()) {
            pdl_wait();  // the input is the predecessor's output
            int stage = 0;
            uint32_t phase = 0;
            TileWalk tw(a);
            for (; tw.b < a.n; tw.next(a)) {
                const int cc = tw.cc, xb = tw.xb, yb = tw.yb, b = tw.b;
                mbar_wait(&empty_bar[stage], phase ^ 1);
                mbar_expect_tx(&full_bar[stage], a.stage_bytes);
                tma_load_4d(smem + static_cast<size_t>(stage) * a.stage_stride, &tmap, &full_bar[stage], cc * CC,
                            xb * TW * S - a.pl, yb * a.th * S - a.pt, b);
                if (++stage == kDwStages) {
                    stage = 0;
                    phase ^= 1;
                }
            }
        }
        return;
    }

    // ===== consumers =====
    const int cw = tid % WORDS;  // channel word inside the chunk
    const int x = tid / WORDS;   // output column inside the tile
    const bool has_lut = a.ep.post_lut != nullptr;
    const int zp_m = a.ep.zp_out - kMagicI;
    // table index clamp and shared-memory address in the same two instructions (see finish4)
    const int lut_base = static_cast<int>(smem_u32(s_lut));
    int lut_lo = kMagicI - a.ep.zp_out - 128 - lut_base;
    asm("mov.b32 %0, %0;" : "+r"(lut_lo));
    int stage = 0;
    uint32_t phase = 0;
    uint32_t wk[3][4];
    uint64_t mu[2], ba[2];
    const bool top_pad = a.pt > 0;                                   // output row 0 reads a padded row
    const bool bot_pad = (a.oh - 1) * S - a.pt + 2 >= a.h;           // output row oh-1 does
    // The grid is a multiple of the channel-chunk count (host), so this CTA keeps ONE chunk for its
    // whole life: per-channel constants and the seed table are set up once, before the tile loop
    // (and before the griddepcontrol wait: they are constants).
    TileWalk walk(a);
    const int cc = walk.cc;
    const int ch = cc * CC + cw * 4;          // first of this thread's four channels
    const bool ch_ok = ch < a.cp;
    {
        {
            const int chs = ch_ok ? ch : 0;
#pragma unroll
            for (int ky = 0; ky < 3; ky++) {
                const uint4 wv = __ldg(reinterpret_cast<const uint4 *>(a.wrow + ky * a.cp + chs));
                wk[ky][0] = wv.x, wk[ky][1] = wv.y, wk[ky][2] = wv.z, wk[ky][3] = wv.w;
            }
            const float4 m4 = __ldg(reinterpret_cast<const float4 *>(a.ep.mult + chs));
            const float4 b4 = __ldg(reinterpret_cast<const float4 *>(a.ep.badd + chs));
            mu[0] = f2_pack(m4.x, m4.y), mu[1] = f2_pack(m4.z, m4.w);
            ba[0] = f2_pack(b4.x, b4.y), ba[1] = f2_pack(b4.z, b4.w);
            // the seed table of this chunk
            for (int i = tid; i < 16 * CC; i += kDwConsumers) {
                const int c = i % CC, cls = i / CC;
                const int cg = cc * CC + c;
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
            asm volatile("bar.sync 1, %0;" ::"n"(kDwConsumers) : "memory");
        }
    }
    for (; walk.b < a.n; walk.next(a)) {
        const int xb = walk.xb, yb = walk.yb, b = walk.b;
        const int ox = xb * TW + x;
        const int oy0 = yb * a.th;
        const int rows_out = min(a.th, a.oh - oy0);
        const bool col_ok = ox < a.ow && ch_ok;

        // the output buffer may alias a tensor the predecessor still reads (returns at once after
        // the first tile; the chunk setup above -- constants only -- overlaps the predecessor's tail)
        pdl_wait();
        mbar_wait(&full_bar[stage], phase);
        uint8_t *tile = smem + static_cast<size_t>(stage) * a.stage_stride;

        if (col_ok) {
            const uint32_t *tw = reinterpret_cast<const uint32_t *>(tile) + x * S * WORDS + cw;
            // 32-bit byte offset into the output (host-checked < 2^32)
            uint32_t ooff = ((static_cast<uint32_t>(b) * a.oh + oy0) * a.ow + ox) * a.cp + ch;
            const uint32_t orow = static_cast<uint32_t>(a.ow) * a.cp;

            auto taps = [&](int r, uint32_t (&v)[4]) {
                const uint32_t *p = tw + r * (TWI * WORDS);
                taps3(p[0], p[WORDS], p[2 * WORDS], v);
            };
            // a fresh accumulator set is produced by the ky = 0 dp4a itself: the addend is the seed
            // of the output row it starts (row `yo` of the tile), fetched with one 128-bit shared load
            const int colc = ((ox * S - a.pl < 0) ? 1 : 0) | ((ox * S - a.pl + 2 >= a.w) ? 2 : 0);
            // shared-memory addresses of this thread's seeds: interior rows, and the tile row (if any)
            // that is the image's last row with a padded ky = 2; only tile row 0 can have ky = 0 padded
            const uint32_t seed_mid = smem_u32(s_seed + colc * CC + cw * 4);
            const uint32_t seed_bot = seed_mid + 8 * CC * 4;
            const int y_bot = bot_pad ? a.oh - 1 - oy0 : -1;       // tile row that is the last image row
            const uint32_t seed_row0 = seed_mid + (((top_pad && oy0 == 0) ? 4 : 0) | (y_bot == 0 ? 8 : 0)) * CC * 4;
            auto first_at = [&](int (&acc)[4], const uint32_t (&v)[4], uint32_t seed_addr) {
                int s0, s1, s2, s3;
                asm volatile("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(s0), "=r"(s1), "=r"(s2), "=r"(s3) : "r"(seed_addr));
                acc[0] = __dp4a(static_cast<int>(v[0]), static_cast<int>(wk[0][0]), s0);
                acc[1] = __dp4a(static_cast<int>(v[1]), static_cast<int>(wk[0][1]), s1);
                acc[2] = __dp4a(static_cast<int>(v[2]), static_cast<int>(wk[0][2]), s2);
                acc[3] = __dp4a(static_cast<int>(v[3]), static_cast<int>(wk[0][3]), s3);
            };
            // a fresh accumulator set is produced by the ky = 0 dp4a itself: the addend is the seed
            // of the output row it starts (tile row yo >= 1), one 128-bit shared load
            auto first = [&](int (&acc)[4], const uint32_t (&v)[4], int yo) {
                first_at(acc, v, yo == y_bot ? seed_bot : seed_mid);
            };
            auto store = [&](const int (&acc)[4]) {
                *reinterpret_cast<uint32_t *>(a.out + ooff) = dw_requant4<MODE>(acc, mu, ba, a.ep, s_lut, has_lut, zp_m, lut_lo, lut_base);
                ooff += orow;
            };
            int accA[4], accB[4], accC[4];
            uint32_t v[4];
            if (S == 1) {
                // input row r feeds output rows r (ky 0), r - 1 (ky 1), r - 2 (ky 2, completes it).
                // rows_out is uniform over the CTA: the exits below do not diverge.
                taps(0, v);
                first_at(accA, v, seed_row0);
                taps(1, v);
                dp4(accA, v, wk[1]), first(accB, v, 1);
                for (int y = 0;; y += 3) {
                    taps(y + 2, v);
                    dp4(accA, v, wk[2]), dp4(accB, v, wk[1]), first(accC, v, y + 2);
                    store(accA);
                    if (y + 1 >= rows_out) break;
                    taps(y + 3, v);
                    dp4(accB, v, wk[2]), dp4(accC, v, wk[1]), first(accA, v, y + 3);
                    store(accB);
                    if (y + 2 >= rows_out) break;
                    taps(y + 4, v);
                    dp4(accC, v, wk[2]), dp4(accA, v, wk[1]), first(accB, v, y + 4);
                    store(accC);
                    if (y + 3 >= rows_out) break;
                }
            } else {
                // output row y reads input rows 2y (ky 0), 2y + 1 (ky 1), 2y + 2 (ky 2 = ky 0 of row y + 1)
                taps(0, v);
                first_at(accA, v, seed_row0);
                for (int y = 0;; y += 2) {
                    taps(2 * y + 1, v);
                    dp4(accA, v, wk[1]);
                    taps(2 * y + 2, v);
                    dp4(accA, v, wk[2]), first(accB, v, y + 1);
                    store(accA);
                    if (y + 1 >= rows_out) break;
                    taps(2 * y + 3, v);
                    dp4(accB, v, wk[1]);
                    taps(2 * y + 4, v);
                    dp4(accB, v, wk[2]), first(accA, v, y + 2);
                    store(accB);
                    if (y + 2 >= rows_out) break;
                }
            }
        }

        // hand the slot back (it was only read)
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[stage]);
        if (++stage == kDwStages) {
            stage = 0;
            phase ^= 1;
        }
    }
}

struct DwCfg {
    int cc, tw;
};

template <int S, int CC, int TW>
static int launch_cfg(int mode, int grid, size_t smem, cudaStream_t s, const CUtensorMap &tm, const DwTmaArgs &a,
                      int dev)
{
#define B200_DW_CASE(M)                                                                              \
    case M: {                                                                                        \
        static bool attr[64] = {};                                                                   \
        if (dev >= 0 && dev < 64 && !attr[dev]) {                                                    \
            B200_CUDA_CHECK(cudaFuncSetAttribute(dw3x3_tma_kernel<S, CC, TW, M>,                     \
                                                 cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024)); \
            attr[dev] = true;                                                                        \
        }                                                                                            \
        B200_CUDA_CHECK(launch_kernel(dw3x3_tma_kernel<S, CC, TW, M>, dim3(grid), dim3(CC / 4 * TW + 32), smem, s, tm, a)); \
        break;                                                                                       \
    }
    switch (mode) {
        B200_DW_CASE(DW_PLAIN)
        B200_DW_CASE(DW_RELU)
        B200_DW_CASE(DW_RELU6)
        B200_DW_CASE(DW_LUT)
        default:
            B200_DW_CASE(DW_GENERIC)
    }
#undef B200_DW_CASE
    return B200_OK;
}

}  // namespace b200

using namespace b200;

// called by b200_dwconv2d (dwconv.cu) for int8 3x3, dilation 1, stride 1 / 2; `wrow` is the
// ky-major repack of the depthwise weights built by b200_opt/quant.c
int b200_dwconv3x3_tma_launch(const b200_dwconv_desc *d, const void *wrow, void *stream)
{
    const int S = d->stride_h;
    // tile shape: (CC, TW) with CC/4 * TW = 256 or 224 threads; prefer full column use and a thin halo
    static const DwCfg cfgs[6] = {{32, 32}, {64, 16}, {128, 8}, {32, 28}, {64, 14}, {128, 7}};
    int best = 0;
    double best_score = -1;
    for (int i = 0; i < 6; i++) {
        const int cc = cfgs[i].cc, tw = cfgs[i].tw;
        const int xb = (d->ow + tw - 1) / tw, cb = (d->cp + cc - 1) / cc;
        const int twi = S * (tw - 1) + 3;
        const double use = (double)d->ow / (xb * tw) * (double)d->cp / (cb * cc);
        const double halo = (double)(S * tw) / twi;
        const double score = use * halo;
        if (score > best_score + 1e-9) best_score = score, best = i;
    }
    const int CC = cfgs[best].cc, TW = cfgs[best].tw;
    const int TWI = S * (TW - 1) + 3;
    // rows per tile: ~33 KB per halo slot (3 slots + the seed table, 2 CTAs per SM): tall tiles amortise the per-tile
    // prologue (tile decode, per-channel constants, zero-point patch) over more row steps
    static const int slot_kb = getenv("SHL_B200_DW_SLOT_KB") ? atoi(getenv("SHL_B200_DW_SLOT_KB")) : 33;
    static const int ctas_per_sm = getenv("SHL_B200_DW_CTAS") ? atoi(getenv("SHL_B200_DW_CTAS")) : 2;
    int thi_max = (slot_kb * 1024) / (TWI * CC);
    if (thi_max > 256) thi_max = 256;  // TMA box limit
    int th = (thi_max - 3) / S + 1;
    if (th < 1) th = 1;
    if (th > d->oh) th = d->oh;
    int ybands = (d->oh + th - 1) / th;
    th = (d->oh + ybands - 1) / ybands;
    // small batches: shorter tiles until every SM has one (the halo rows cost less than idle SMs)
    {
        const long long per_band = static_cast<long long>(d->n) * ((d->ow + TW - 1) / TW) * ((d->cp + CC - 1) / CC);
        while (per_band * ybands < sm_count() && th > 4) {
            th = (th + 1) / 2;
            ybands = (d->oh + th - 1) / th;
            th = (d->oh + ybands - 1) / ybands;
        }
    }
    const int thi = S * (th - 1) + 3;

    DwTmaArgs a;
    a.n = d->n, a.c = d->c, a.cp = d->cp, a.h = d->h, a.w = d->w, a.oh = d->oh, a.ow = d->ow;
    a.pt = d->pad_top, a.pl = d->pad_left, a.th = th, a.thi = thi;
    a.ybands = ybands, a.xbands = (d->ow + TW - 1) / TW, a.cchunks = (d->cp + CC - 1) / CC;
    a.stage_bytes = thi * TWI * CC;
    a.stage_stride = (a.stage_bytes + 127) & ~127;
    a.wrow = static_cast<const uint32_t *>(wrow);
    a.out = static_cast<int8_t *>(d->out);
    a.zp_in = d->zp_in;
    a.ep = make_epi(d->ep);

    alignas(64) CUtensorMap tm;
    int rc = encode_tmap_nhwc_u8(&tm, d->in, d->n, d->h, d->w, d->cp, CC, TWI, thi);
    if (rc) return rc;

    const long long tiles = static_cast<long long>(d->n) * a.ybands * a.xbands * a.cchunks;
    if (tiles >= (1ll << 31) || static_cast<long long>(d->n) * d->oh * d->ow * d->cp >= (1ll << 32)) {
        set_error("b200_dwconv2d: %lld tiles / output bytes exceed the 32-bit indices of the 3x3 kernel", tiles);
        return B200_ERR_UNSUPPORTED;
    }
    // the grid is a multiple of the channel-chunk count: a CTA sees one chunk only (its constants
    // are set up once) and walks the spatial tiles with a fixed stride
    long long cap = static_cast<long long>(sm_count()) * ctas_per_sm;
    if (cap > a.cchunks) cap -= cap % a.cchunks;
    if (cap < a.cchunks) cap = a.cchunks;
    const int grid = static_cast<int>(tiles < cap ? tiles : cap);  // tiles is a multiple of cchunks
    {
        int q = grid / a.cchunks;  // spatial tiles skipped per step, as mixed-radix digits
        a.dxb = q % a.xbands;
        q /= a.xbands;
        a.dyb = q % a.ybands;
        a.db = q / a.ybands;
    }
    const size_t smem = static_cast<size_t>(kDwStages) * a.stage_stride + 128;
    int mode;
    if (d->ep.post_lut)
        mode = d->ep.act == B200_ACT_NONE ? DW_LUT : DW_GENERIC;
    else
        mode = d->ep.act == B200_ACT_NONE ? DW_PLAIN : (d->ep.act == B200_ACT_RELU ? DW_RELU : DW_RELU6);
    int dev = 0;
    B200_CUDA_CHECK(cudaGetDevice(&dev));
    cudaStream_t s = (cudaStream_t)stream;
#define B200_DW_CFG(I, CCv, TWv)                                                        \
    case I:                                                                             \
        rc = S == 1 ? launch_cfg<1, CCv, TWv>(mode, grid, smem, s, tm, a, dev)          \
                    : launch_cfg<2, CCv, TWv>(mode, grid, smem, s, tm, a, dev);         \
        break;
    switch (best) {
        B200_DW_CFG(0, 32, 32)
        B200_DW_CFG(1, 64, 16)
        B200_DW_CFG(2, 128, 8)
        B200_DW_CFG(3, 32, 28)
        B200_DW_CFG(4, 64, 14)
        default:
            B200_DW_CFG(5, 128, 7)
    }
#undef B200_DW_CFG
    if (rc) return rc;
    B200_LAUNCH_CHECK();
    return B200_OK;
}
