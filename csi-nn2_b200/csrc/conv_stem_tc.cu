// conv_stem_tc.cu -- a network's FIRST conv layer (3 input channels, 3x3 or 7x7, NCHW input) as an
// implicit GEMM on tcgen05: the im2col matrix is never written to HBM.  A tile is a segment of up
// to 128 output pixels of ONE output row.  The kh input rows x 3 channels it needs arrive by TMA (a 4-D map over the
// NCHW image: box = {160 bytes of a row, kh rows, 3 channels}; two boxes side by side for stride 2, whose 128 pixels span
// more than the 256-byte box limit), issued one tile ahead into a double-buffered staging area -- no thread spends
// registers, address arithmetic or scoreboard waits on the image.  The TMA zero-fills outside the image; the quantised
// domain wants zp_in there, so the (16-byte aligned) out-of-image chunks of the staged rows are overwritten before the
// gather.  Each thread then picks its pixel's K bytes out of the staged rows into a 128B-swizzled A tile: K is laid out
// as one group per staged row (c, ky) holding the KW bytes from the window's first column on, rounded up to whole words
// (the surplus bytes meet zero weights), so a pixel's A row is a handful of unaligned WORD reads, not K byte reads.  One
// thread issues tcgen05.mma kind::i8 against the resident weights, and the same workers requantise the accumulators
// (contract of include/b200nn.h) and store pixel-major rows.
//
// Why: on CUDA cores this layer costs K/4 dp4a per output plus the epilogue -- 38 instructions per
// output for 3x3x3 -> 32 channels, 136 us at batch 256, six times its HBM floor.  On the tensor
// core the per-output cost is the epilogue alone; the gather is paid once per pixel, not per channel.
//
// CTA = 4 or 6 worker groups of 4 warps + 1 MMA warp, one CTA per SM, persistent.  A group owns one A tile, two staging
// buffers and two TMEM accumulators and software-pipelines itself: wait for tile i's rows -> gather -> signal the MMA warp
// -> epilogue of tile i-1 (whose MMA ran meanwhile; the A tile is free again by then), with tile i+1's TMA in flight.  The
// groups run staggered, so the gather of one overlaps the epilogue of another; no CTA-wide barrier inside the loop.
// Accumulators are seeded with ibias (+ the magic constant) by tcgen05.st, like the GEMM's.
//
// Replaces, for this shape, shl_rvv_conv_im2col_gemm_int8
// (source/thead_rvv/int8/convolution_gemm_int8.c:106-170); semantics shl_ref_conv2d_quant
// (source/reference/convolution.c:370).
#include <stdlib.h>

#include "common.cuh"

namespace b200 {

// worker groups per CTA: 6 when the tiles are small (K <= 128, N <= 32: 24 worker warps keep the
// schedulers busy through the shared-memory and TMEM latencies of gather and epilogue), else 4
// (TMEM: groups x 2 x N columns <= 512; shared memory: groups x (A tile + 2 staging buffers))
__host__ __device__ constexpr int stem_groups(int k, int nch) { return (k <= 128 && nch <= 2) ? 6 : 4; }
// staging geometry: bytes of an image row one TMA box covers, boxes per tile, bytes between boxes / buffers
__host__ __device__ constexpr int stem_hb() { return 160; }
__host__ __device__ constexpr int stem_boxes(int sw) { return sw == 2 ? 2 : 1; }
__host__ __device__ constexpr int stem_box_stride(int rows) { return (rows * stem_hb() + 127) / 128 * 128; }
__host__ __device__ constexpr int stem_threads(int k, int nch) { return (stem_groups(k, nch) * 4 + 1) * 32; }

struct StemArgs {
    int n, h, w, o, oh, ow, cp_out;
    int sh, sw, pt, pl;
    int ldw;            // weight row pitch in bytes
    const int8_t *in;   // NCHW
    const int8_t *wt;   // [O][ldw], k = (ky, kx, c)
    int8_t *out;        // [n*oh*ow][cp_out]
    int zp_in;
    uint32_t idesc;
    uint32_t oh_inv;    // ceil(2^32 / oh): tile row -> image by multiply-high, corrected by one compare
    EpiScalars ep;
};

__device__ __forceinline__ void group_bar_sync(int g)
{
    asm volatile("bar.sync %0, 128;" ::"r"(g + 1) : "memory");
}

// 32 lanes x 16 columns seed store / load share the GEMM's helpers (common.cuh)

template <int C, int KH, int KW, int SW, int NCH, int MODE>
__global__ void __launch_bounds__(stem_threads(C * KH * KW, NCH), 1)
conv_stem_tc_kernel(const __grid_constant__ CUtensorMap tmap, const StemArgs a)
{
    constexpr int kStemGroups = stem_groups(C * KH * KW, NCH);
    constexpr int kStemThreads = stem_threads(C * KH * KW, NCH);
    constexpr int ROWS = KH * C;                          // staged input rows per tile, (c, ky) order (the TMA box's)
    constexpr int HB = stem_hb();                         // bytes of an image row per box
    constexpr int NB = stem_boxes(SW);                    // boxes per tile (stride 2: pixels 0-63 and 64-127)
    constexpr int PB = 128 / NB;                          // pixels per box
    constexpr uint32_t BOX = stem_box_stride(ROWS);       // bytes between boxes (TMA destinations are 128-byte aligned)
    constexpr uint32_t STG = NB * BOX;                    // one staging buffer
    constexpr int K = C * KH * KW;
    // K order of the A tile: one group of KWP = KW rounded up to whole words per staged row (c, ky), the group's bytes
    // being kx = 0 .. KWP - 1 -- i.e. a pixel's A row is ROWS unaligned word reads out of the staged rows, not K byte
    // reads.  The bytes kx >= KW (the next pixels) meet zero weights.
    constexpr int KWP = (KW + 3) / 4 * 4;
    constexpr int GW = KWP / 4;                 // words per group
    constexpr int KG = ROWS * KWP;              // K as laid out
    constexpr int KP = (KG + 31) / 32 * 32;     // padded to whole MMA k-steps
    constexpr int KSTEPS = KP / 32;
    constexpr int ATOMS = (KP + 127) / 128;     // 128-byte swizzle atoms per row
    constexpr int KCHUNKS = (KG + 15) / 16;     // 16-byte chunks of an A row that hold data
    constexpr int N = NCH * 16;
    constexpr bool MAGIC = K <= 128;
    constexpr uint32_t A_TILE = ATOMS * 128 * 128;   // bytes
    constexpr uint32_t B_TILE = ATOMS * N * 128;
    // a pixel's last word read ends inside its box: (PB - 1) * SW + 15 (alignment shift) + KWP + 3 (the funnel's second word)
    static_assert((PB - 1) * SW + 15 + KWP + 3 < HB, "TMA box too narrow for the tile");

    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t *smem_a = smem;                                   // [group][A_TILE]
    uint8_t *smem_b = smem_a + kStemGroups * A_TILE;           // [ATOMS][N rows][128 B]
    uint8_t *smem_s = smem_b + B_TILE;                         // [group][2][NB][ROWS][HB] staged input rows
    float *s_mu = reinterpret_cast<float *>(smem_s + kStemGroups * 2 * STG);
    float *s_ba = s_mu + N;
    int *s_ib = reinterpret_cast<int *>(s_ba + N);
    uint8_t *s_lut = reinterpret_cast<uint8_t *>(s_ib + N);
    uint64_t *a_full = reinterpret_cast<uint64_t *>(s_lut + 256);   // [group][2]: A tile written, accumulator stage s
    uint64_t *mma_done = a_full + kStemGroups * 2;                  // [group][2]
    uint64_t *in_full = mma_done + kStemGroups * 2;                 // [group][2]: staging buffer filled by the TMA
    uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(in_full + kStemGroups * 2);

    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const bool has_lut = a.ep.post_lut != nullptr;
    pdl_launch_dependents();  // see launch_kernel (common.cuh): the setup below reads constants only

    // ---- one-time setup: weights into the swizzled B tile, per-channel parameters, barriers, TMEM
    for (int i = tid; i < ATOMS * N * 8; i += kStemThreads) {
        const int chunk = i & 7, row = (i >> 3) % N, atom = i / (8 * N);
        const int k0 = atom * 128 + chunk * 16;
        uint32_t wv[4] = {0, 0, 0, 0};
        if (row < a.o) {
#pragma unroll
            for (int e = 0; e < 16; e++) {
                const int k = k0 + e;                       // = (c * KH + ky) * KWP + kx
                const int grp = k / KWP, kx = k - grp * KWP;
                const int c = grp / KH, ky = grp - c * KH;
                if (grp < ROWS && kx < KW)                   // a.wt: k = (ky, kx, c)
                    wv[e >> 2] |= static_cast<uint32_t>(static_cast<uint8_t>(a.wt[row * a.ldw + (ky * KW + kx) * C + c])) << (8 * (e & 3));
            }
        }
        *reinterpret_cast<uint4 *>(smem_b + atom * (N * 128) + row * 128 + ((chunk ^ (row & 7)) << 4)) =
            make_uint4(wv[0], wv[1], wv[2], wv[3]);
    }
    for (int o = tid; o < N; o += kStemThreads) {
        s_mu[o] = o < a.o ? a.ep.mult[o] : 0.f;
        s_ba[o] = o < a.o ? a.ep.badd[o] : 0.f;
        s_ib[o] = (o < a.o ? a.ep.ibias[o] : 0) + (MAGIC ? kMagicI : 0);
    }
    if (has_lut)
        for (int i = tid; i < 256; i += kStemThreads) s_lut[i] = static_cast<uint8_t>(a.ep.post_lut[i]);
    if (tid == 0) {
        tma_prefetch_desc(&tmap);
        for (int i = 0; i < kStemGroups * 2; i++) {
            mbar_init(&a_full[i], 1);
            mbar_init(&mma_done[i], 1);
            mbar_init(&in_full[i], 1);
        }
        mbar_fence_init();
    }
    if (warp == kStemGroups * 4) {
        tmem_alloc(tmem_ptr, 512);
        tmem_relinquish();
    }
    fence_proxy_async_smem();  // the B tile was written through the generic proxy
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    const int nseg = (a.ow + 127) / 128;
    const int tiles = a.n * a.oh * nseg;                // < 2^31, host-checked
    // tile (i, g) of this CTA: consecutive tiles go to consecutive CTAs
    auto tile_of = [&](int i, int g) { return (i * kStemGroups + g) * static_cast<int>(gridDim.x) + static_cast<int>(blockIdx.x); };

    if (warp == kStemGroups * 4) {
        // ===== MMA issuer =====
        if (elect_one()) {
            for (int i = 0;; i++) {
                bool any = false;
                for (int g = 0; g < kStemGroups; g++) {
                    if (tile_of(i, g) >= tiles) continue;
                    any = true;
                    const int s = i & 1;
                    mbar_wait(&a_full[g * 2 + s], (i >> 1) & 1);
                    tc_fence_after();
                    const uint32_t tmem_d = tmem_base + (g * 2 + s) * N;
                    const uint32_t a_addr = smem_u32(smem_a + g * A_TILE);
#pragma unroll
                    for (int ks = 0; ks < KSTEPS; ks++) {
                        const uint64_t adesc = umma_desc_sw128(a_addr + (ks >> 2) * (128 * 128)) + 2 * (ks & 3);
                        const uint64_t bdesc = umma_desc_sw128(smem_u32(smem_b) + (ks >> 2) * (N * 128)) + 2 * (ks & 3);
                        tc_mma_i8(tmem_d, adesc, bdesc, a.idesc, 1u);  // accumulators are pre-seeded
                    }
                    tc_commit(&mma_done[g * 2 + s]);
                }
                if (!any) break;
            }
        }
    } else {
        // ===== worker groups: gather + epilogue =====
        const int g = warp >> 2;
        const int r = tid & 127;                      // row of the tile = TMEM lane
        const uint32_t tlane = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16);
        const EpiScalars &ep = a.ep;
        const int zp_m = ep.zp_out - kMagicI;
        const int lut_base = static_cast<int>(smem_u32(s_lut));
        int lut_lo = kMagicI - ep.zp_out - 128 - lut_base;
        asm("mov.b32 %0, %0;" : "+r"(lut_lo));
        const uint32_t zpw = 0x01010101u * static_cast<uint32_t>(a.zp_in & 0xFF);
        uint8_t *stg = smem_s + g * (2 * STG);

        // seed both accumulators of this group
        for (int ch = 0; ch < NCH; ch++) {
            uint32_t ib[16];
#pragma unroll
            for (int j4 = 0; j4 < 4; j4++) {
                const int4 i4 = *reinterpret_cast<const int4 *>(&s_ib[ch * 16 + j4 * 4]);
                ib[j4 * 4] = i4.x, ib[j4 * 4 + 1] = i4.y, ib[j4 * 4 + 2] = i4.z, ib[j4 * 4 + 3] = i4.w;
            }
            tmem_st_32x16(tlane + (g * 2 + 0) * N + ch * 16, ib);
            tmem_st_32x16(tlane + (g * 2 + 1) * N + ch * 16, ib);
        }
        tmem_st_wait();
        tc_fence_before();

        // tile -> (image, output row, first output column); decoded when its TMA is issued and carried through the gather
        // (one iteration later) and the epilogue (two later)
        struct TilePos { int b, oy, ox0; };
        auto decode = [&](int t) {
            TilePos tp;
            const int seg = nseg == 1 ? 0 : t % nseg;
            const int row = nseg == 1 ? t : t / nseg;
            tp.b = static_cast<int>(__umulhi(static_cast<uint32_t>(row), a.oh_inv));
            if (tp.b * a.oh > row) tp.b--;
            if ((tp.b + 1) * a.oh <= row) tp.b++;
            tp.oy = row - tp.b * a.oh;
            tp.ox0 = seg * 128;
            return tp;
        };
        // the tile's input rows: boxes of HB bytes starting at the tile's first input column rounded down to 16 (so that
        // whatever lies outside the image is whole 16-byte chunks), kh rows from its first input row, all channels
        auto issue = [&](const TilePos &tp, int buf) {
            const int x0 = (tp.ox0 * SW - a.pl) & ~15;
            mbar_expect_tx(&in_full[g * 2 + buf], NB * ROWS * HB);
#pragma unroll
            for (int bx = 0; bx < NB; bx++)
                tma_load_4d(stg + buf * STG + bx * BOX, &tmap, &in_full[g * 2 + buf], x0 + bx * PB * SW, tp.oy * a.sh - a.pt, 0, tp.b);
        };

        auto epilogue = [&](int i, const TilePos &tp) {
            const int s = i & 1;
            const int ox = tp.ox0 + r;
            const int p = (tp.b * a.oh + tp.oy) * a.ow + ox;
            const bool pix_ok = ox < a.ow;
            mbar_wait(&mma_done[g * 2 + s], (i >> 1) & 1);
            tc_fence_after();
            int8_t *dst = a.out + static_cast<size_t>(p) * a.cp_out;
#pragma unroll 1
            for (int ch = 0; ch < NCH; ch++) {
                uint32_t acc[16], ib[16];
                const uint32_t taddr = tlane + (g * 2 + s) * N + ch * 16;
                tmem_ld_32x16(taddr, acc);
#pragma unroll
                for (int j4 = 0; j4 < 4; j4++) {
                    const int4 i4 = *reinterpret_cast<const int4 *>(&s_ib[ch * 16 + j4 * 4]);
                    ib[j4 * 4] = i4.x, ib[j4 * 4 + 1] = i4.y, ib[j4 * 4 + 2] = i4.z, ib[j4 * 4 + 3] = i4.w;
                }
                tmem_ld_wait();
                tmem_st_32x16(taddr, ib);  // re-seed for tile i + 2
                uint32_t packed[4];
#pragma unroll
                for (int j4 = 0; j4 < 4; j4++) {
                    const float4 m4 = *reinterpret_cast<const float4 *>(&s_mu[ch * 16 + j4 * 4]);
                    const float4 b4 = *reinterpret_cast<const float4 *>(&s_ba[ch * 16 + j4 * 4]);
                    int t[4];
                    requant_pair<MAGIC>(acc[j4 * 4 + 0], acc[j4 * 4 + 1], f2_pack(m4.x, m4.y), f2_pack(b4.x, b4.y), t[0], t[1]);
                    requant_pair<MAGIC>(acc[j4 * 4 + 2], acc[j4 * 4 + 3], f2_pack(m4.z, m4.w), f2_pack(b4.z, b4.w), t[2], t[3]);
                    packed[j4] = finish4<MODE>(t, ep, s_lut, has_lut, zp_m, lut_lo, lut_base);
                }
                if (pix_ok && ch * 16 < a.cp_out) *reinterpret_cast<uint4 *>(dst + ch * 16) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
            }
            tmem_st_wait();
            tc_fence_before();
        };

        pdl_wait();  // image and output buffer belong to the predecessor until here
        TilePos tp_prev = {0, 0, 0}, tp_cur = {0, 0, 0}, tp_next = {0, 0, 0};
        if (tile_of(0, g) < tiles) {
            tp_next = decode(tile_of(0, g));
            if (r == 0) issue(tp_next, 0);
        }
        for (int i = 0;; i++) {
            const bool has = tile_of(i, g) < tiles;
            tp_prev = tp_cur;
            tp_cur = tp_next;
            if (has) {
                const int buf = i & 1;
                uint8_t *sb = stg + buf * STG;
                // the next tile's rows fly during this tile's gather and the previous tile's epilogue; their buffer was last
                // read by the gather of tile i - 1, which every thread of the group left through a barrier
                if (tile_of(i + 1, g) < tiles) {
                    tp_next = decode(tile_of(i + 1, g));
                    if (r == 0) issue(tp_next, buf ^ 1);
                }
                mbar_wait(&in_full[g * 2 + buf], (i >> 1) & 1);
                const int xs = tp_cur.ox0 * SW - a.pl;       // image column of the window of pixel 0
                const int x0 = xs & ~15;
                if (zpw != 0u) {
                    // ---- out-of-image chunks of the staged rows: zero-filled by the TMA, zp_in by contract
                    const int iy0 = tp_cur.oy * a.sh - a.pt;
                    constexpr int CH = HB / 16;
                    for (int idx = r; idx < NB * ROWS * CH; idx += 128) {
                        const int chunk = idx % CH, rr = (idx / CH) % ROWS, bx = idx / (CH * ROWS);
                        const int ky = rr % KH;
                        const int x = x0 + bx * PB * SW + chunk * 16;
                        if (static_cast<unsigned>(iy0 + ky) >= static_cast<unsigned>(a.h) || x < 0 || x >= a.w)
                            *reinterpret_cast<uint4 *>(sb + bx * BOX + rr * HB + chunk * 16) = make_uint4(zpw, zpw, zpw, zpw);
                    }
                    group_bar_sync(g);
                }
                // the A tile is free once the MMAs of tile i - 1 have read it
                if (i > 0) mbar_wait(&mma_done[g * 2 + ((i - 1) & 1)], ((i - 1) >> 1) & 1);
                // ---- gather this thread's pixel: per staged row (c, ky) the KWP bytes from its window's first column on,
                // as GW unaligned words (two aligned loads + one byte permute each; the sub-word offset is the same on
                // every row because rows start word-aligned)
                const int boff = (r % PB) * SW + (xs - x0);
                const uint32_t sel = 0x3210u + 0x1111u * static_cast<uint32_t>(boff & 3);
                const uint32_t *srow = reinterpret_cast<const uint32_t *>(sb + (r / PB) * BOX) + (boff >> 2);
                uint32_t xw[KCHUNKS * 4];
#pragma unroll
                for (int j = ROWS * GW; j < KCHUNKS * 4; j++) xw[j] = 0;
#pragma unroll
                for (int row = 0; row < ROWS; row++) {
                    uint32_t w[GW + 1];
#pragma unroll
                    for (int j = 0; j <= GW; j++) w[j] = srow[row * (HB / 4) + j];
#pragma unroll
                    for (int j = 0; j < GW; j++) xw[row * GW + j] = __byte_perm(w[j], w[j + 1], sel);
                }
                uint8_t *arow = smem_a + g * A_TILE + r * 128;
                // chunks past KCHUNKS are never written: whatever they hold multiplies the zero weights of k >= KG
#pragma unroll
                for (int j = 0; j < KCHUNKS; j++) {
                    const int atom = j >> 3, chunk = j & 7;
                    *reinterpret_cast<uint4 *>(arow + atom * (128 * 128) + ((chunk ^ (r & 7)) << 4)) =
                        make_uint4(xw[j * 4], xw[j * 4 + 1], xw[j * 4 + 2], xw[j * 4 + 3]);
                }
                fence_proxy_async_smem();
            }
            // every thread of the group has written its row (and finished re-seeding, one
            // iteration ago, the accumulator this tile will use)
            group_bar_sync(g);
            if (has && r == 0) mbar_arrive(&a_full[g * 2 + (i & 1)]);
            if (i > 0) epilogue(i - 1, tp_prev);
            if (!has) break;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == kStemGroups * 4) tmem_dealloc(tmem_base, 512);
}

template <int C, int KH, int KW, int SW, int NCH>
static int stem_launch(int mode, int grid, cudaStream_t s, const CUtensorMap &tm, const StemArgs &a, int dev)
{
    constexpr int KP = (KH * C * ((KW + 3) / 4 * 4) + 31) / 32 * 32;
    constexpr int ATOMS = (KP + 127) / 128;
    constexpr int N = NCH * 16;
    constexpr int STG = stem_boxes(SW) * stem_box_stride(KH * C);
    constexpr int kStemGroups = stem_groups(C * KH * KW, NCH);
    constexpr int kStemThreads = stem_threads(C * KH * KW, NCH);
    const size_t smem = 1024 + static_cast<size_t>(kStemGroups) * ATOMS * 128 * 128 + ATOMS * N * 128 +
                        kStemGroups * 2 * STG + N * 12 + 256 + kStemGroups * 6 * 8 + 16;
#define B200_STEM_CASE(M)                                                                                   \
    case M: {                                                                                               \
        static bool attr[64] = {};                                                                          \
        if (dev >= 0 && dev < 64 && !attr[dev]) {                                                           \
            B200_CUDA_CHECK(cudaFuncSetAttribute(conv_stem_tc_kernel<C, KH, KW, SW, NCH, M>,                    \
                                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));  \
            attr[dev] = true;                                                                               \
        }                                                                                                   \
        B200_CUDA_CHECK(launch_kernel(conv_stem_tc_kernel<C, KH, KW, SW, NCH, M>, dim3(grid), dim3(kStemThreads),  \
                                      smem, s, tm, a));                                                      \
        break;                                                                                              \
    }
    switch (mode) {
        B200_STEM_CASE(EPI_PLAIN)
        B200_STEM_CASE(EPI_RELU)
        B200_STEM_CASE(EPI_RELU6)
        B200_STEM_CASE(EPI_LUT)
        default:
            B200_STEM_CASE(EPI_GENERIC)
    }
#undef B200_STEM_CASE
    return B200_OK;
}

}  // namespace b200

using namespace b200;

// called by b200_conv2d_direct (conv_direct.cu); returns B200_ERR_UNSUPPORTED when the shape is not
// one of the tensor-core stems (the caller then runs the dp4a kernel)
int b200_conv_stem_tc_launch(const b200_conv_direct_desc *d, void *stream)
{
    const bool s3 = d->c == 3 && d->kh == 3 && d->kw == 3;
    const bool s7 = d->c == 3 && d->kh == 7 && d->kw == 7;
    const int nch = (d->o + 15) / 16;
    // the image rows arrive by TMA: every row must start on a 16-byte boundary (base and row pitch)
    if (!(s3 || s7) || d->dil_h != 1 || d->dil_w != 1 || nch > 4 || d->cp_out < nch * 16 || d->cp_out % 16 ||
        (reinterpret_cast<uintptr_t>(d->out) & 15) || d->stride_h != d->stride_w || d->stride_w > 2 ||
        (s7 && d->stride_w != 2) || d->w % 16 || (reinterpret_cast<uintptr_t>(d->in) & 15) || d->pad_left > 15)
        return B200_ERR_UNSUPPORTED;
    const long long total = static_cast<long long>(d->n) * d->oh * d->ow;
    if (total >= (1ll << 31) - 128 || static_cast<long long>(d->n) * d->c * d->h * d->w >= (1ll << 31))
        return B200_ERR_UNSUPPORTED;
    StemArgs a;
    a.n = d->n, a.h = d->h, a.w = d->w, a.o = d->o, a.oh = d->oh, a.ow = d->ow, a.cp_out = d->cp_out;
    a.sh = d->stride_h, a.sw = d->stride_w, a.pt = d->pad_top, a.pl = d->pad_left, a.ldw = d->ldw;
    a.in = static_cast<const int8_t *>(d->in), a.wt = static_cast<const int8_t *>(d->wt);
    a.out = static_cast<int8_t *>(d->out), a.zp_in = d->zp_in, a.ep = make_epi(d->ep);
    // column chunks the kernel is instantiated for: 3x3 -> 1 / 2 / 4, 7x7 -> 2 / 4
    const int nch_k = s3 ? (nch <= 2 ? nch : 4) : (nch <= 2 ? 2 : 4);
    a.idesc = umma_idesc(2 /*S32*/, 1 /*S8*/, 128, nch_k * 16);
    a.oh_inv = d->oh == 1 ? 0xFFFFFFFFu : static_cast<uint32_t>(((1ull << 32) + d->oh - 1) / d->oh);
    const long long tiles_ll = static_cast<long long>(d->n) * d->oh * ((d->ow + 127) / 128);
    if (tiles_ll >= (1ll << 31) / 8) return B200_ERR_UNSUPPORTED;
    const int tiles = static_cast<int>(tiles_ll);
    const int grid = tiles < sm_count() ? tiles : sm_count();
    int mode;
    if (d->ep.post_lut)
        mode = d->ep.act == B200_ACT_NONE ? EPI_LUT : EPI_GENERIC;
    else
        mode = d->ep.act == B200_ACT_NONE ? EPI_PLAIN : (d->ep.act == B200_ACT_RELU ? EPI_RELU : EPI_RELU6);
    int dev = 0;
    B200_CUDA_CHECK(cudaGetDevice(&dev));
    cudaStream_t s = (cudaStream_t)stream;
    // the NCHW image as a 4-D tensor (row bytes, rows, channels, images); box = {160 bytes, kh rows, 3 channels, 1}
    alignas(64) CUtensorMap tm;
    int rc = encode_tmap_nhwc_u8_nb(&tm, d->in, d->n, d->c, d->h, d->w, stem_hb(), d->kh, d->c, 1);
    if (rc) return rc;
    if (s3 && d->stride_w == 2)
        rc = nch <= 1 ? stem_launch<3, 3, 3, 2, 1>(mode, grid, s, tm, a, dev)
                      : (nch == 2 ? stem_launch<3, 3, 3, 2, 2>(mode, grid, s, tm, a, dev)
                                  : stem_launch<3, 3, 3, 2, 4>(mode, grid, s, tm, a, dev));
    else if (s3)
        rc = nch <= 1 ? stem_launch<3, 3, 3, 1, 1>(mode, grid, s, tm, a, dev)
                      : (nch == 2 ? stem_launch<3, 3, 3, 1, 2>(mode, grid, s, tm, a, dev)
                                  : stem_launch<3, 3, 3, 1, 4>(mode, grid, s, tm, a, dev));
    else
        rc = nch <= 2 ? stem_launch<3, 7, 7, 2, 2>(mode, grid, s, tm, a, dev)
                      : stem_launch<3, 7, 7, 2, 4>(mode, grid, s, tm, a, dev);
    if (rc) return rc;
    B200_LAUNCH_CHECK();
    return B200_OK;
}
