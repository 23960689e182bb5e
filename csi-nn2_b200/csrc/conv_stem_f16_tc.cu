// conv_stem_f16_tc.cu -- the fp16 twin of conv_stem_tc.cu: a network's first 3x3 conv layer (3 input channels, NCHW
// fp16 input, stride 1 / 2) as an implicit GEMM on tcgen05 kind::f16 with f32 accumulators in TMEM.
//
// Why: conv_direct_f16_kernel (one thread per pixel, all output channels in registers, 27 x 16 packed FMAs against
// shared-memory weights) takes 266 us for MobileNetV1's first layer at batch 256 -- a sixth of the fp16 step and six
// times the layer's HBM floor.  On the tensor core the per-pixel work is the gather (K halves out of the staged rows)
// and a bias / relu / convert epilogue.
//
// Same structure as the int8 stem: a tile = up to 128 output pixels of one output row; its 3 input rows x 3 channels
// arrive by TMA (the fp16 image seen as a byte tensor: boxes of 160 bytes = 80 halves, four per stride-2 tile, two per
// stride-1 tile), one tile ahead, double-buffered; out-of-image bytes are the TMA's zeros, which is what fp16 padding
// wants, so nothing is patched.  K is laid out as one group of 4 halves per staged row (c, ky): the 3 taps from the
// window's first column on plus one surplus half that meets a zero weight, i.e. a pixel's A row is 9 unaligned 8-byte
// reads.  Worker groups of 4 warps gather -> signal the MMA warp -> run the epilogue of the previous tile; accumulators
// are seeded with the f32 bias by tcgen05.st.
//
// Replaces, for this shape, the fp16 im2col + GEMM of shl_rvv_conv_im2col_gemm_fp16
// (source/thead_rvv/fp16/convolution_gemm_fp16.c); semantics shl_ref_conv2d_f32 on converted tensors
// (source/reference/convolution.c:20): f32 accumulation, tolerance 1e-3 relative (tests/test_gpu_parity.py).
#include "common.cuh"

namespace b200 {

// worker groups per CTA: TMEM holds groups x 2 accumulators x N columns <= 512
__host__ __device__ constexpr int f16_stem_groups(int nch) { return nch <= 2 ? 6 : 4; }
__host__ __device__ constexpr int f16_stem_threads(int nch) { return (f16_stem_groups(nch) * 4 + 1) * 32; }

struct StemF16Args {
    int n, h, w, o, oh, ow, cp_out;
    int sh, pt, pl;
    int ldw;               // weight row pitch in halves
    const __half *wt;      // [O][ldw], k = (ky, kx, c)
    __half *out;           // [n*oh*ow][cp_out]
    const float *bias;     // [O] or null
    int act;
    uint32_t idesc;
    uint32_t oh_inv;       // ceil(2^32 / oh)
};

__device__ __forceinline__ void f16_group_bar_sync(int g)
{
    asm volatile("bar.sync %0, 128;" ::"r"(g + 1) : "memory");
}

template <int SW, int NCH>
__global__ void __launch_bounds__(f16_stem_threads(NCH), 1) conv_stem_f16_tc_kernel(const __grid_constant__ CUtensorMap tmap, const StemF16Args a)
{
    constexpr int kF16StemGroups = f16_stem_groups(NCH);
    constexpr int kF16StemThreads = f16_stem_threads(NCH);
    constexpr int C = 3, KH = 3, KW = 3;
    constexpr int ROWS = KH * C;             // staged input rows per tile, (c, ky) order (the TMA box's)
    constexpr int HB = 160;                  // bytes of an image row per box (80 halves)
    constexpr int PB = SW == 2 ? 32 : 64;    // pixels per box
    constexpr int NB = 128 / PB;             // boxes per tile
    constexpr uint32_t BOX = (ROWS * HB + 127) / 128 * 128;
    constexpr uint32_t STG = NB * BOX;
    constexpr int KWP = 4;                   // halves per (c, ky) group: 3 taps + 1 that meets a zero weight
    constexpr int GW = KWP / 2;              // words per group
    constexpr int KG = ROWS * KWP;           // 36 halves
    constexpr int KP = (KG + 15) / 16 * 16;  // 48: whole k-steps of 16 halves
    constexpr int KSTEPS = KP / 16;
    constexpr int KCHUNKS = (KG * 2 + 15) / 16;  // 16-byte chunks of an A row that hold data
    constexpr int N = NCH * 16;
    constexpr uint32_t A_TILE = 128 * 128;
    constexpr uint32_t B_TILE = N * 128;
    // a pixel's last word read ends inside its box: ((PB - 1) * SW + 7 (alignment shift) + KWP) halves + the funnel's second word
    static_assert(((PB - 1) * SW + 7 + KWP) * 2 + 4 <= HB, "TMA box too narrow for the tile");

    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t *smem_a = smem;                                   // [group][A_TILE]
    uint8_t *smem_b = smem_a + kF16StemGroups * A_TILE;        // [N rows][128 B]
    uint8_t *smem_s = smem_b + B_TILE;                         // [group][2][NB][ROWS][HB]
    float *s_bias = reinterpret_cast<float *>(smem_s + kF16StemGroups * 2 * STG);
    uint64_t *a_full = reinterpret_cast<uint64_t *>(s_bias + N);     // [group][2]
    uint64_t *mma_done = a_full + kF16StemGroups * 2;
    uint64_t *in_full = mma_done + kF16StemGroups * 2;
    uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(in_full + kF16StemGroups * 2);

    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    pdl_launch_dependents();

    // ---- one-time setup: weights into the swizzled B tile (K order (c, ky, kx) in groups of four halves), bias, barriers, TMEM
    for (int i = tid; i < N * 8; i += kF16StemThreads) {
        const int chunk = i & 7, row = i >> 3;
        uint32_t wv[4] = {0, 0, 0, 0};
        if (row < a.o) {
#pragma unroll
            for (int e = 0; e < 8; e++) {
                const int k = chunk * 8 + e;                 // = (c * KH + ky) * KWP + kx
                const int grp = k / KWP, kx = k - grp * KWP;
                const int c = grp / KH, ky = grp - c * KH;
                if (grp < ROWS && kx < KW) {
                    const uint32_t hv = __half_as_ushort(a.wt[row * a.ldw + (ky * KW + kx) * C + c]);
                    wv[e >> 1] |= hv << (16 * (e & 1));
                }
            }
        }
        *reinterpret_cast<uint4 *>(smem_b + row * 128 + ((chunk ^ (row & 7)) << 4)) = make_uint4(wv[0], wv[1], wv[2], wv[3]);
    }
    for (int o = tid; o < N; o += kF16StemThreads) s_bias[o] = (o < a.o && a.bias) ? a.bias[o] : 0.f;
    if (tid == 0) {
        tma_prefetch_desc(&tmap);
        for (int i = 0; i < kF16StemGroups * 2; i++) {
            mbar_init(&a_full[i], 1);
            mbar_init(&mma_done[i], 1);
            mbar_init(&in_full[i], 1);
        }
        mbar_fence_init();
    }
    if (warp == kF16StemGroups * 4) {
        tmem_alloc(tmem_ptr, 512);
        tmem_relinquish();
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    const int nseg = (a.ow + 127) / 128;
    const int tiles = a.n * a.oh * nseg;
    auto tile_of = [&](int i, int g) { return (i * kF16StemGroups + g) * static_cast<int>(gridDim.x) + static_cast<int>(blockIdx.x); };

    if (warp == kF16StemGroups * 4) {
        // ===== MMA issuer =====
        if (elect_one()) {
            for (int i = 0;; i++) {
                bool any = false;
                for (int g = 0; g < kF16StemGroups; g++) {
                    if (tile_of(i, g) >= tiles) continue;
                    any = true;
                    const int s = i & 1;
                    mbar_wait(&a_full[g * 2 + s], (i >> 1) & 1);
                    tc_fence_after();
                    const uint32_t tmem_d = tmem_base + (g * 2 + s) * N;
                    const uint32_t a_addr = smem_u32(smem_a + g * A_TILE);
#pragma unroll
                    for (int ks = 0; ks < KSTEPS; ks++) {
                        const uint64_t adesc = umma_desc_sw128(a_addr) + 2 * ks;   // 16 halves = 32 bytes per k-step
                        const uint64_t bdesc = umma_desc_sw128(smem_u32(smem_b)) + 2 * ks;
                        tc_mma_f16(tmem_d, adesc, bdesc, a.idesc, 1u);  // accumulators are pre-seeded with the bias
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
        uint8_t *stg = smem_s + g * (2 * STG);

        auto seed = [&](uint32_t taddr, int ch) {
            uint32_t ib[16];
#pragma unroll
            for (int j4 = 0; j4 < 4; j4++) {
                const uint4 i4 = *reinterpret_cast<const uint4 *>(&s_bias[ch * 16 + j4 * 4]);
                ib[j4 * 4] = i4.x, ib[j4 * 4 + 1] = i4.y, ib[j4 * 4 + 2] = i4.z, ib[j4 * 4 + 3] = i4.w;
            }
            tmem_st_32x16(taddr, ib);
        };
        for (int ch = 0; ch < NCH; ch++) {
            seed(tlane + (g * 2 + 0) * N + ch * 16, ch);
            seed(tlane + (g * 2 + 1) * N + ch * 16, ch);
        }
        tmem_st_wait();
        tc_fence_before();

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
        // boxes of 80 halves from the tile's first input column rounded down to 8 halves (16 bytes); coordinates in bytes
        auto issue = [&](const TilePos &tp, int buf) {
            const int x0 = (tp.ox0 * SW - a.pl) & ~7;
            mbar_expect_tx(&in_full[g * 2 + buf], NB * ROWS * HB);
#pragma unroll
            for (int bx = 0; bx < NB; bx++)
                tma_load_4d(stg + buf * STG + bx * BOX, &tmap, &in_full[g * 2 + buf], (x0 + bx * PB * SW) * 2, tp.oy * a.sh - a.pt, 0, tp.b);
        };

        auto epilogue = [&](int i, const TilePos &tp) {
            const int s = i & 1;
            const int ox = tp.ox0 + r;
            const int p = (tp.b * a.oh + tp.oy) * a.ow + ox;
            const bool pix_ok = ox < a.ow;
            mbar_wait(&mma_done[g * 2 + s], (i >> 1) & 1);
            tc_fence_after();
            __half *dst = a.out + static_cast<size_t>(p) * a.cp_out;
#pragma unroll 1
            for (int ch = 0; ch < NCH; ch++) {
                uint32_t acc[16];
                const uint32_t taddr = tlane + (g * 2 + s) * N + ch * 16;
                tmem_ld_32x16(taddr, acc);
                tmem_ld_wait();
                seed(taddr, ch);  // re-seed for tile i + 2
                uint32_t pk[8];
#pragma unroll
                for (int q = 0; q < 8; q++) {
                    const int o = ch * 16 + q * 2;
                    float f0 = act_f(__uint_as_float(acc[2 * q]), a.act), f1 = act_f(__uint_as_float(acc[2 * q + 1]), a.act);
                    f0 = o < a.o ? f0 : 0.f, f1 = o + 1 < a.o ? f1 : 0.f;
                    const __half2 hv = __floats2half2_rn(f0, f1);
                    pk[q] = *reinterpret_cast<const uint32_t *>(&hv);
                }
                if (pix_ok && ch * 16 < a.cp_out) *reinterpret_cast<uint4 *>(dst + ch * 16) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                if (pix_ok && ch * 16 + 8 < a.cp_out) *reinterpret_cast<uint4 *>(dst + ch * 16 + 8) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
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
                const uint8_t *sb = stg + buf * STG;
                // the next tile's rows fly during this tile's gather and the previous tile's epilogue; their buffer was last
                // read by the gather of tile i - 1, which every thread of the group left through a barrier
                if (tile_of(i + 1, g) < tiles) {
                    tp_next = decode(tile_of(i + 1, g));
                    if (r == 0) issue(tp_next, buf ^ 1);
                }
                mbar_wait(&in_full[g * 2 + buf], (i >> 1) & 1);
                // the A tile is free once the MMAs of tile i - 1 have read it
                if (i > 0) mbar_wait(&mma_done[g * 2 + ((i - 1) & 1)], ((i - 1) >> 1) & 1);
                // ---- gather this thread's pixel: per staged row (c, ky) four halves from its window's first column on, as two
                // unaligned words (three aligned loads + two byte permutes; the sub-word offset is 0 or 2 bytes, the same on
                // every row)
                const int xs = tp_cur.ox0 * SW - a.pl;
                const int boff = ((r % PB) * SW + (xs - (xs & ~7))) * 2;
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
                // chunks past KCHUNKS are never written: whatever they hold multiplies the zero weights of k >= KG -- unless
                // it is a NaN / infinity pattern (0 * inf = NaN): chunk KCHUNKS (the rest of the last k-step) is zeroed too
#pragma unroll
                for (int j = 0; j < (KP * 2 + 15) / 16; j++) {
                    const uint4 v = j < KCHUNKS ? make_uint4(xw[j * 4], xw[j * 4 + 1], xw[j * 4 + 2], xw[j * 4 + 3]) : make_uint4(0, 0, 0, 0);
                    *reinterpret_cast<uint4 *>(arow + ((j ^ (r & 7)) << 4)) = v;
                }
                fence_proxy_async_smem();
            }
            f16_group_bar_sync(g);
            if (has && r == 0) mbar_arrive(&a_full[g * 2 + (i & 1)]);
            if (i > 0) epilogue(i - 1, tp_prev);
            if (!has) break;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == kF16StemGroups * 4) tmem_dealloc(tmem_base, 512);
}

template <int SW, int NCH>
static int stem_f16_launch(int grid, cudaStream_t s, const CUtensorMap &tm, const StemF16Args &a, int dev)
{
    constexpr int kF16StemGroups = f16_stem_groups(NCH);
    constexpr int kF16StemThreads = f16_stem_threads(NCH);
    constexpr int N = NCH * 16;
    constexpr int ROWS = 9, HB = 160;
    constexpr int NB = SW == 2 ? 4 : 2;
    constexpr int STG = NB * ((ROWS * HB + 127) / 128 * 128);
    const size_t smem = 1024 + static_cast<size_t>(kF16StemGroups) * 128 * 128 + N * 128 + kF16StemGroups * 2 * STG + N * 4 +
                        kF16StemGroups * 6 * 8 + 16;
    static bool attr[64] = {};
    if (dev >= 0 && dev < 64 && !attr[dev]) {
        B200_CUDA_CHECK(cudaFuncSetAttribute(conv_stem_f16_tc_kernel<SW, NCH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr[dev] = true;
    }
    B200_CUDA_CHECK(launch_kernel(conv_stem_f16_tc_kernel<SW, NCH>, dim3(grid), dim3(kF16StemThreads), smem, s, tm, a));
    return B200_OK;
}

}  // namespace b200

using namespace b200;

// called by b200_conv2d_direct_f16 (conv_direct.cu); returns B200_ERR_UNSUPPORTED when the shape is not the fp16 3x3 stem
// this kernel covers (the caller then runs conv_direct_f16_kernel)
int b200_conv_stem_f16_tc_launch(const b200_conv_direct_desc *d, void *stream)
{
    const int nch = (d->o + 15) / 16;
    // the image rows arrive by TMA: every row must start on a 16-byte boundary (base and row pitch of 2 * w bytes)
    if (d->c != 3 || d->kh != 3 || d->kw != 3 || d->dil_h != 1 || d->dil_w != 1 || nch > 4 || d->cp_out < d->o || d->cp_out % 8 ||
        (reinterpret_cast<uintptr_t>(d->out) & 15) || d->stride_h != d->stride_w || d->stride_w > 2 || d->w % 8 ||
        (reinterpret_cast<uintptr_t>(d->in) & 15) || d->pad_left > 7 || d->ldw % 2)
        return B200_ERR_UNSUPPORTED;
    const long long total = static_cast<long long>(d->n) * d->oh * d->ow;
    if (total >= (1ll << 31) - 128 || static_cast<long long>(d->n) * d->c * d->h * d->w >= (1ll << 30)) return B200_ERR_UNSUPPORTED;
    const long long tiles_ll = static_cast<long long>(d->n) * d->oh * ((d->ow + 127) / 128);
    if (tiles_ll >= (1ll << 31) / 8) return B200_ERR_UNSUPPORTED;
    StemF16Args a;
    a.n = d->n, a.h = d->h, a.w = d->w, a.o = d->o, a.oh = d->oh, a.ow = d->ow, a.cp_out = d->cp_out;
    a.sh = d->stride_h, a.pt = d->pad_top, a.pl = d->pad_left, a.ldw = d->ldw / 2;
    a.wt = static_cast<const __half *>(d->wt), a.out = static_cast<__half *>(d->out);
    a.bias = d->ep.badd, a.act = d->ep.act;
    const int nch_k = nch <= 2 ? nch : 4;
    a.idesc = umma_idesc(1 /*F32*/, 0 /*F16*/, 128, nch_k * 16);
    a.oh_inv = d->oh == 1 ? 0xFFFFFFFFu : static_cast<uint32_t>(((1ull << 32) + d->oh - 1) / d->oh);
    const int tiles = static_cast<int>(tiles_ll);
    const int grid = tiles < sm_count() ? tiles : sm_count();
    int dev = 0;
    B200_CUDA_CHECK(cudaGetDevice(&dev));
    cudaStream_t s = (cudaStream_t)stream;
    // the NCHW fp16 image as a 4-D BYTE tensor (row bytes, rows, channels, images); box = {160 bytes, 3 rows, 3 channels, 1}
    alignas(64) CUtensorMap tm;
    int rc = encode_tmap_nhwc_u8_nb(&tm, d->in, d->n, d->c, d->h, d->w * 2, 160, d->kh, d->c, 1);
    if (rc) return rc;
    if (d->stride_w == 2)
        rc = nch <= 1 ? stem_f16_launch<2, 1>(grid, s, tm, a, dev) : (nch == 2 ? stem_f16_launch<2, 2>(grid, s, tm, a, dev) : stem_f16_launch<2, 4>(grid, s, tm, a, dev));
    else
        rc = nch <= 1 ? stem_f16_launch<1, 1>(grid, s, tm, a, dev) : (nch == 2 ? stem_f16_launch<1, 2>(grid, s, tm, a, dev) : stem_f16_launch<1, 4>(grid, s, tm, a, dev));
    if (rc) return rc;
    B200_LAUNCH_CHECK();
    return B200_OK;
}
