// dwpw_fused.cu -- one MobileNet block in one kernel: int8 depthwise 3x3 (stride 1 / 2) whose
// requantised result never leaves the SM, feeding the pointwise 1x1 convolution on tcgen05.
//
//   HBM -> TMA halo boxes -> dp4a depthwise warps -> requantise in registers -> int8 written straight
//   into the 128B-swizzled A operand tile in shared memory -> one-thread tcgen05.mma (kind::i8)
//   against the RESIDENT pointwise weights -> TMEM -> seeded requantise epilogue -> swizzled
//   staging -> 4-D TMA store -> HBM
//
// so a block costs one read of the depthwise input and one write of the pointwise output; the
// intermediate tensor (as large as either) is never written or read back, and a network step has
// one launch per block instead of two.
//
// One persistent CTA per SM, 19 warps:
//   warp 0        TMA producer: one 4-D box per (tile, channel chunk) into a ring of halo stages;
//                 the pointwise weights once
//   warp 1        MMA issuer (one elected lane) + TMEM allocation
//   warps 2..17   workers.  Each runs the depthwise stage of tile t -- one output column x 4 channels per
//                 thread sliding down a band of tile rows (the TMA depthwise kernel's inner loop), output =
//                 A operand rows -- and then, while the tensor core multiplies tile t, the pointwise
//                 epilogue of tile t - 1 for its TMEM lane quadrant and a quarter of the columns
//                 (tcgen05.ld -> requantise -> staging, the GEMM's epilogue).  Both halves are bound by
//                 instruction issue, so every warp doing both keeps the two balanced whatever the ratio of
//                 channels to outputs, and at any moment some warps are in the dp4a-heavy phase and others
//                 in the float / table phase (first version: 8 depthwise + 8 epilogue warps -- the 8
//                 depthwise warps alone paced the kernel, slower than the two separate kernels).
//   warp 18       TMA-store warp
// A tile = TW x TH output pixels (<= 256 = two M128 blocks) x ALL channels; pixel (y, x) of the tile is
// A row / TMEM lane / staging row y * TW + x, which is also the order a 4-D TMA store box
// {channels, TW, TH, 1} expects, so partial tiles at the right / bottom edge are clipped by the
// store engine.  Pipelines: halo ring full/empty (TMA <-> depthwise), A tile full/empty double
// buffered (depthwise <-> MMA), TMEM stage full/empty double buffered (MMA <-> epilogue), staging
// full/empty double buffered (epilogue <-> store warp).
//
// Arithmetic: both stages are the contract of include/b200nn.h bit for bit -- the depthwise half is
// dw3x3_tma_kernel's (zero-point padding folded into per-border-class accumulator seeds), the
// pointwise half gemm_tc_kernel's (accumulators seeded with ibias through tcgen05.st).
//
// Replaces the pair shl_rvv_dwconv3x3s1_int8 / s2 -> shl_rvv_conv1x1s1_gemm_int8
// (source/thead_rvv/int8/depthwise_convolution_3x3_int8.c:31,244 ->
// source/thead_rvv/int8/convolution_1x1_int8.c:56) as it occurs 13 times in
// example/c906_mobilenetv1_f16.c; semantics shl_ref_depthwise_conv2d_quant + shl_ref_conv2d_quant
// (source/reference/convolution.c:416,370).
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "common.cuh"

namespace b200 {

constexpr int kFuWorkers = 16;  // warps that run the depthwise stage of tile t, then the pointwise epilogue of tile t - 1
constexpr int kFuWorkerThreads = kFuWorkers * 32;
constexpr int kFuDwThreads = kFuWorkerThreads;
constexpr int kFuThreads = (3 + kFuWorkers) * 32;  // + TMA producer, MMA issuer, store warp
constexpr int kFuMaxStages = 6;
constexpr int kFuAccStride = 256;  // TMEM columns per accumulator stage
constexpr size_t kFuSmemLimit = 226 * 1024;
constexpr int kFuFirstWorker = 2;
constexpr int kFuStoreWarp = kFuFirstWorker + kFuWorkers;

struct DwPwArgs {
    // depthwise geometry
    int n, cp, h, w, oh, ow, pt, pl;
    int tw, th, twi, thi;  // tile: output columns / rows, halo columns / rows
    int xbands, ybands;
    int dxb, dyb, db;      // the grid's stride over tiles as (x band, y band, image) digits
    int cc, words_log2;    // channels per halo chunk, log2(cc / 4)
    int cchunks;           // ceil(cp / cc)
    int rb, rpb;           // row bands per tile (one consumer thread walks one band), rows per band
    int stage_bytes, stage_stride, stages;
    int zp_in;
    const uint32_t *wrow;  // [3 (ky)][cp] words (w[ky][0][c], w[ky][1][c], w[ky][2][c], 0)
    EpiScalars dep;        // depthwise epilogue
    // pointwise
    int o, bn, k_blocks, kslices, mt;  // outputs, MMA N, 128-byte K blocks, 32-byte K slices, M128 blocks per tile
    uint32_t idesc;
    uint32_t a_buf_bytes;  // k_blocks * mt * 16 KB
    uint32_t stg_half_bytes;  // one staging unit: mt * 128 rows of one 128-column half of the tile (two units in a ring)
    int scols, nhalves;    // staging row bytes (16 / 32 / 64 / 128), halves per tile
    EpiScalars pep;        // pointwise epilogue
};

struct __align__(16) FuEpiParams {
    float mult[256];
    float badd[256];
    int32_t ibias[256];  // + kMagicI when the kernel converts through the magic constant
    uint8_t lut[256];
};

struct FuWalk {
    int xb, yb, b;
    __device__ __forceinline__ explicit FuWalk(const DwPwArgs &a)
    {
        uint32_t q = blockIdx.x;
        xb = q % a.xbands;
        q /= a.xbands;
        yb = q % a.ybands;
        b = q / a.ybands;
    }
    __device__ __forceinline__ void next(const DwPwArgs &a)
    {
        xb += a.dxb;
        if (xb >= a.xbands) xb -= a.xbands, yb++;
        yb += a.dyb;
        if (yb >= a.ybands) yb -= a.ybands, b++;
        b += a.db;
    }
};

__device__ __forceinline__ void tma_store_4d(const CUtensorMap *m, const void *smem_src, int c0, int c1, int c2, int c3)
{
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];\n" ::
                     "l"(reinterpret_cast<uint64_t>(m)),
                 "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}

__device__ __forceinline__ void fu_taps3(uint32_t a, uint32_t b, uint32_t c, uint32_t (&v)[4])
{
    const uint32_t lo = __byte_perm(a, b, 0x5140);  // (a.c0, b.c0, a.c1, b.c1)
    const uint32_t hi = __byte_perm(a, b, 0x7362);  // (a.c2, b.c2, a.c3, b.c3)
    v[0] = __byte_perm(lo, c, 0x4410);
    v[1] = __byte_perm(lo, c, 0x5532);
    v[2] = __byte_perm(hi, c, 0x6610);
    v[3] = __byte_perm(hi, c, 0x7732);
}

__device__ __forceinline__ void fu_dp4(int (&acc)[4], const uint32_t (&v)[4], const uint32_t (&w)[4])
{
#pragma unroll
    for (int e = 0; e < 4; e++) acc[e] = __dp4a(static_cast<int>(v[e]), static_cast<int>(w[e]), acc[e]);
}

template <int S, int DMODE, int PMODE, bool MAGIC>
__global__ void __launch_bounds__(kFuThreads, 1)
dwpw_kernel(const __grid_constant__ CUtensorMap tm_in, const __grid_constant__ CUtensorMap tm_w,
            const __grid_constant__ CUtensorMap tm_out, const DwPwArgs a)
{
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int cp = a.cp;
    uint8_t *smem_b = smem;                                          // resident pointwise weights
    uint8_t *smem_a = smem_b + a.k_blocks * a.bn * 128;              // two A tiles
    uint8_t *staging = smem_a + 2 * a.a_buf_bytes;                   // two output staging buffers
    uint8_t *halo = staging + 2 * a.stg_half_bytes;                  // halo ring
    uint32_t *s_wrow = reinterpret_cast<uint32_t *>(halo + a.stages * a.stage_stride);
    float *s_dmult = reinterpret_cast<float *>(s_wrow + 3 * cp);
    float *s_dbadd = s_dmult + cp;
    int *s_seed = reinterpret_cast<int *>(s_dbadd + cp);             // [16 classes][cp]
    FuEpiParams *epi = reinterpret_cast<FuEpiParams *>(s_seed + 16 * cp);
    uint8_t *s_dlut = reinterpret_cast<uint8_t *>(epi + 1);
    uint64_t *halo_full = reinterpret_cast<uint64_t *>(s_dlut + 256);
    uint64_t *halo_empty = halo_full + kFuMaxStages;
    uint64_t *a_full = halo_empty + kFuMaxStages;
    uint64_t *a_empty = a_full + 2;
    uint64_t *tmem_full = a_empty + 2;
    uint64_t *tmem_empty = tmem_full + 2;
    uint64_t *stg_full = tmem_empty + 2;
    uint64_t *stg_empty = stg_full + 2;
    uint64_t *b_bar = stg_empty + 2;
    uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(b_bar + 1);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    pdl_launch_dependents();

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tm_in);
        tma_prefetch_desc(&tm_w);
        tma_prefetch_desc(&tm_out);
        for (int i = 0; i < a.stages; i++) {
            mbar_init(&halo_full[i], 1);
            mbar_init(&halo_empty[i], kFuWorkers);
        }
        for (int i = 0; i < 2; i++) {
            mbar_init(&a_full[i], kFuWorkers);
            mbar_init(&a_empty[i], 1);
            mbar_init(&tmem_full[i], 1);
            mbar_init(&tmem_empty[i], kFuWorkers);
            mbar_init(&stg_full[i], kFuWorkers);
            mbar_init(&stg_empty[i], 1);
        }
        mbar_init(b_bar, 1);
        mbar_fence_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_ptr, 512);
        tmem_relinquish();
    }
    if (warp >= kFuFirstWorker && warp < kFuStoreWarp) {
        const int t = threadIdx.x - kFuFirstWorker * 32;
        if (t < 256 && a.pep.post_lut != nullptr) epi->lut[t] = static_cast<uint8_t>(a.pep.post_lut[t]);
        if (t >= 256 && a.dep.post_lut != nullptr) s_dlut[t - 256] = static_cast<uint8_t>(a.dep.post_lut[t - 256]);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    const int MT = a.mt;

    if (warp == 0) {
        // ===== TMA producer =====
        if (elect_one()) {
            mbar_expect_tx(b_bar, a.k_blocks * a.bn * 128);
            for (int kb = 0; kb < a.k_blocks; kb++) tma_load_2d(smem_b + kb * a.bn * 128, &tm_w, b_bar, kb * 128, 0);
            pdl_wait();  // the depthwise input is the predecessor's output
            int stage = 0;
            uint32_t phase = 0;
            for (FuWalk tw(a); tw.b < a.n; tw.next(a)) {
                const int x0 = tw.xb * a.tw * S - a.pl, y0 = tw.yb * a.th * S - a.pt;
                for (int c = 0; c < a.cchunks; c++) {
                    mbar_wait_parked(&halo_empty[stage], phase ^ 1);
                    mbar_expect_tx(&halo_full[stage], a.stage_bytes);
                    tma_load_4d(halo + static_cast<size_t>(stage) * a.stage_stride, &tm_in, &halo_full[stage], c * a.cc, x0,
                                y0, tw.b);
                    if (++stage == a.stages) stage = 0, phase ^= 1;
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (elect_one()) {
            mbar_wait_parked(b_bar, 0);
            int local = 0;
            for (FuWalk tw(a); tw.b < a.n; tw.next(a), local++) {
                const int buf = local & 1;
                const uint32_t ph = (local >> 1) & 1;
                mbar_wait_parked(&tmem_empty[buf], ph);  // seeded (ph + 1) times
                mbar_wait_parked(&a_full[buf], ph);
                tc_fence_after();
                const uint32_t a_base = smem_u32(smem_a + buf * a.a_buf_bytes);
                for (int m = 0; m < MT; m++) {
                    const uint32_t tmem_d = tmem_base + buf * kFuAccStride + m * a.bn;
                    for (int ks = 0; ks < a.kslices; ks++) {
                        const int kb = ks >> 2, k = ks & 3;
                        const uint64_t adesc = umma_desc_sw128(a_base + (kb * MT + m) * 16384) + 2 * k;
                        const uint64_t bdesc = umma_desc_sw128(smem_u32(smem_b + kb * a.bn * 128)) + 2 * k;
                        tc_mma_i8(tmem_d, adesc, bdesc, a.idesc, 1u);
                    }
                }
                tc_commit(&a_empty[buf]);    // the A tile may be overwritten once these MMAs have read it
                tc_commit(&tmem_full[buf]);  // accumulators complete -> epilogue
            }
        }
    } else if (warp == kFuStoreWarp) {
        // ===== TMA-store warp: one staging unit (a 128-column half of a tile) per store =====
        if (elect_one()) {
            pdl_wait();  // the output buffer may alias a tensor the predecessor still reads
            uint32_t unit = 0;
            for (FuWalk tw(a); tw.b < a.n; tw.next(a)) {
                for (int hf = 0; hf < a.nhalves; hf++, unit++) {
                    const int sb = unit & 1;
                    mbar_wait_parked(&stg_full[sb], (unit >> 1) & 1);
                    tma_store_4d(&tm_out, staging + sb * a.stg_half_bytes, hf * 128, tw.xb * a.tw, tw.yb * a.th, tw.b);
                    tma_store_commit();
                    tma_store_wait_read<0>();
                    mbar_arrive(&stg_empty[sb]);
                }
            }
            tma_store_wait<0>();
        }
    } else {
        // ===== workers: depthwise for tile t, then the pointwise epilogue of tile t - 1 =====
        const int wid = warp - kFuFirstWorker;
        const int quad = warp & 3;   // TMEM lane quadrant this warp may access
        const int part = wid >> 2;   // 0..3: which 16-column chunks of a tile's accumulators it drains
        const int tid = threadIdx.x - kFuFirstWorker * 32;
        // ---- pointwise epilogue state ----
        const EpiScalars &pe = a.pep;
        const int bn = a.bn;
        const int nsub = bn >> 4;
        const int pzp_m = pe.zp_out - kMagicI;
        const int plut_base = static_cast<int>(smem_u32(epi->lut));
        int plut_lo = kMagicI - pe.zp_out - 128 - plut_base;
        asm("mov.b32 %0, %0;" : "+r"(plut_lo));
        const bool p_has_lut = pe.post_lut != nullptr;
        if (tid < bn) {
            const bool ok = tid < a.o;
            epi->mult[tid] = ok ? pe.mult[tid] : 0.f;
            epi->badd[tid] = ok ? pe.badd[tid] : 0.f;
            epi->ibias[tid] = ((ok && pe.ibias) ? pe.ibias[tid] : 0) + (MAGIC ? kMagicI : 0);
        }
        // ---- depthwise state: constants of the whole layer into shared memory ----
        const int WORDS = 1 << a.words_log2;
        const int cw = tid & (WORDS - 1);
        const int xi = tid >> a.words_log2;
        const int x = xi % a.tw;     // output column inside the tile
        const int rbi = xi / a.tw;   // row band this thread walks
        const bool item_ok = rbi < a.rb;
        const EpiScalars &de = a.dep;
        const bool d_has_lut = de.post_lut != nullptr;
        const int dzp_m = de.zp_out - kMagicI;
        const int dlut_base = static_cast<int>(smem_u32(s_dlut));
        int dlut_lo = kMagicI - de.zp_out - 128 - dlut_base;
        asm("mov.b32 %0, %0;" : "+r"(dlut_lo));
        const bool top_pad = a.pt > 0;
        const bool bot_pad = (a.oh - 1) * S - a.pt + 2 >= a.h;
        for (int i = tid; i < 3 * cp; i += kFuWorkerThreads) s_wrow[i] = __ldg(a.wrow + i);
        for (int i = tid; i < cp; i += kFuWorkerThreads) {
            s_dmult[i] = __ldg(de.mult + i);
            s_dbadd[i] = __ldg(de.badd + i);
        }
        // accumulator seeds [row class][column class][channel]: ibias + kMagicI + zp_in * (sum of the weights of
        // the taps that fall into the padding), see dwconv3x3_tma.cu
        for (int i = tid; i < 16 * cp; i += kFuWorkerThreads) {
            const int c = i % cp, cls = i / cp;
            int padsum = 0;
#pragma unroll
            for (int ky = 0; ky < 3; ky++) {
                const uint32_t wv = __ldg(a.wrow + ky * cp + c);
                const bool rowpad = (ky == 0 && (cls & 4)) || (ky == 2 && (cls & 8));
#pragma unroll
                for (int kx = 0; kx < 3; kx++) {
                    const bool colpad = (kx == 0 && (cls & 1)) || (kx == 2 && (cls & 2));
                    if (rowpad || colpad) padsum += static_cast<int8_t>(wv >> (8 * kx));
                }
            }
            s_seed[i] = __ldg(de.ibias + c) + kMagicI + a.zp_in * padsum;
        }
        asm volatile("bar.sync 1, %0;" ::"n"(kFuWorkerThreads) : "memory");

        const uint32_t scols = a.scols;
        const uint32_t swz_mask = scols >= 128 ? 7u : (scols == 64 ? 3u : (scols == 32 ? 1u : 0u));
        const uint32_t tquad = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
        // seed both accumulator stages with ibias (+ magic): the MMAs always accumulate
        for (int sub = part; sub < nsub; sub += 4) {
            uint32_t ib[16];
#pragma unroll
            for (int j4 = 0; j4 < 4; j4++) {
                const int4 i4 = *reinterpret_cast<const int4 *>(&epi->ibias[sub * 16 + j4 * 4]);
                ib[j4 * 4 + 0] = i4.x, ib[j4 * 4 + 1] = i4.y, ib[j4 * 4 + 2] = i4.z, ib[j4 * 4 + 3] = i4.w;
            }
            for (int a2 = 0; a2 < 2; a2++)
                for (int m = 0; m < MT; m++) tmem_st_32x16(tquad + a2 * kFuAccStride + m * bn + sub * 16, ib);
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
            mbar_arrive(&tmem_empty[0]);
            mbar_arrive(&tmem_empty[1]);
        }

        uint32_t unit = 0;  // staging units handed to the store warp so far
        auto epilogue_tile = [&](int local) {
            const int buf = local & 1;
            const uint32_t ph = (local >> 1) & 1;
            mbar_wait_parked(&tmem_full[buf], ph);
            tc_fence_after();
            const uint32_t taddr = tquad + buf * kFuAccStride;
            for (int hf = 0; hf < a.nhalves; hf++, unit++) {
                const int sb = unit & 1;
                uint8_t *stg = staging + sb * a.stg_half_bytes;
                // the store issued from this staging unit two units ago must have finished reading it
                mbar_wait_parked(&stg_empty[sb], ((unit >> 1) & 1) ^ 1);
                const int sub_end = min(nsub, hf * 8 + 8);
                for (int sub = hf * 8 + part; sub < sub_end; sub += 4) {
                    uint64_t mu2[8], ba2[8];
                    uint32_t ib[16];
#pragma unroll
                    for (int j4 = 0; j4 < 4; j4++) {
                        const float4 m4 = *reinterpret_cast<const float4 *>(&epi->mult[sub * 16 + j4 * 4]);
                        const float4 b4 = *reinterpret_cast<const float4 *>(&epi->badd[sub * 16 + j4 * 4]);
                        const int4 i4 = *reinterpret_cast<const int4 *>(&epi->ibias[sub * 16 + j4 * 4]);
                        mu2[j4 * 2] = f2_pack(m4.x, m4.y), mu2[j4 * 2 + 1] = f2_pack(m4.z, m4.w);
                        ba2[j4 * 2] = f2_pack(b4.x, b4.y), ba2[j4 * 2 + 1] = f2_pack(b4.z, b4.w);
                        ib[j4 * 4 + 0] = i4.x, ib[j4 * 4 + 1] = i4.y, ib[j4 * 4 + 2] = i4.z, ib[j4 * 4 + 3] = i4.w;
                    }
                    const uint32_t colb = (static_cast<uint32_t>(sub) & 7u) * 16u;
#pragma unroll 1
                    for (int m = 0; m < MT; m++) {
                        uint32_t r[16];
                        tmem_ld_32x16(taddr + m * bn + sub * 16, r);
                        tmem_ld_wait();
                        tmem_st_32x16(taddr + m * bn + sub * 16, ib);  // re-seed for the tile after next
                        uint32_t packed[4];
#pragma unroll
                        for (int j4 = 0; j4 < 4; j4++) {
                            int t[4];
                            requant_pair<MAGIC>(r[j4 * 4 + 0], r[j4 * 4 + 1], mu2[j4 * 2], ba2[j4 * 2], t[0], t[1]);
                            requant_pair<MAGIC>(r[j4 * 4 + 2], r[j4 * 4 + 3], mu2[j4 * 2 + 1], ba2[j4 * 2 + 1], t[2], t[3]);
                            packed[j4] = finish4<PMODE>(t, pe, epi->lut, p_has_lut, pzp_m, plut_lo, plut_base);
                        }
                        uint32_t off = static_cast<uint32_t>(m * 128 + quad * 32 + lane) * scols + colb;
                        off ^= ((off >> 7) & swz_mask) << 4;
                        *reinterpret_cast<uint4 *>(stg + off) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
                    }
                }
                if (hf == a.nhalves - 1) {
                    // accumulators drained and re-seeded: the MMA warp may refill this TMEM stage
                    tmem_st_wait();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&tmem_empty[buf]);
                }
                // hand this warp's part of the staged unit to the store warp
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(&stg_full[sb]);
            }
        };

        const int TWI_WORDS = a.twi << a.words_log2;  // words per halo row
        const int y0 = rbi * a.rpb;                    // first tile row of this thread's band
        // byte offset of this thread's first tap word inside a halo stage, and of its first A operand row
        const uint32_t halo_off = static_cast<uint32_t>((y0 * S) * TWI_WORDS + ((x * S) << a.words_log2) + cw) * 4u;
        const uint32_t p0 = static_cast<uint32_t>(y0 * a.tw + x);
        const uint32_t a_row_step = static_cast<uint32_t>(a.tw) << 7;
        uint32_t wk[3][4];
        uint64_t mu[2], ba[2];
        auto load_chunk_consts = [&](int ch) {
#pragma unroll
            for (int ky = 0; ky < 3; ky++) {
                const uint4 wv = *reinterpret_cast<const uint4 *>(s_wrow + ky * cp + ch);
                wk[ky][0] = wv.x, wk[ky][1] = wv.y, wk[ky][2] = wv.z, wk[ky][3] = wv.w;
            }
            const float4 m4 = *reinterpret_cast<const float4 *>(s_dmult + ch);
            const float4 b4 = *reinterpret_cast<const float4 *>(s_dbadd + ch);
            mu[0] = f2_pack(m4.x, m4.y), mu[1] = f2_pack(m4.z, m4.w);
            ba[0] = f2_pack(b4.x, b4.y), ba[1] = f2_pack(b4.z, b4.w);
        };
        int stage = 0;
        uint32_t phase = 0;
        int local = 0;
        for (FuWalk walk(a); walk.b < a.n; walk.next(a), local++) {
            const int buf = local & 1;
            const int ox = walk.xb * a.tw + x;
            const int oy0 = walk.yb * a.th;
            const int rows_out = min(a.rpb, min(a.th, a.oh - oy0) - y0);  // rows of this thread's band inside the image
            const bool work = item_ok && ox < a.ow && rows_out > 0;
            const int colc = ((ox * S - a.pl < 0) ? 1 : 0) | ((ox * S - a.pl + 2 >= a.w) ? 2 : 0);
            const int y_bot = bot_pad ? a.oh - 1 - oy0 - y0 : -1;  // band row that is the image's last row
            const uint32_t a_tile = smem_u32(smem_a + buf * a.a_buf_bytes);
            // the MMAs that read this A buffer two tiles ago must have retired
            mbar_wait_parked(&a_empty[buf], ((local >> 1) & 1) ^ 1);
            for (int c = 0; c < a.cchunks; c++) {
                mbar_wait_parked(&halo_full[stage], phase);
                const int ch = c * a.cc + cw * 4;  // first of this thread's four channels
                if (work && ch < cp) {
                    load_chunk_consts(ch);
                    const uint32_t hbase = smem_u32(halo + static_cast<size_t>(stage) * a.stage_stride) + halo_off;
                    const uint32_t pix = 4u << a.words_log2;     // bytes between horizontally adjacent pixels
                    const uint32_t rowb = static_cast<uint32_t>(TWI_WORDS) * 4u;
                    uint32_t rowp = hbase;                       // running address of the next input row's first tap
                    auto taps = [&](uint32_t (&v)[4]) {
                        uint32_t t0, t1, t2;
                        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(t0) : "r"(rowp));
                        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(t1) : "r"(rowp + pix));
                        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(t2) : "r"(rowp + 2 * pix));
                        rowp += rowb;
                        fu_taps3(t0, t1, t2, v);
                    };
                    // seeds: interior rows from registers; the image's first / last row (padded ky = 0 / 2) from the table
                    const uint32_t seed_mid = smem_u32(s_seed + colc * cp + ch);
                    int sm0, sm1, sm2, sm3;
                    asm volatile("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(sm0), "=r"(sm1), "=r"(sm2), "=r"(sm3) : "r"(seed_mid));
                    auto first_at = [&](int (&acc)[4], const uint32_t (&v)[4], uint32_t seed_addr) {
                        int s0, s1, s2, s3;
                        asm volatile("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];"
                                     : "=r"(s0), "=r"(s1), "=r"(s2), "=r"(s3)
                                     : "r"(seed_addr));
                        acc[0] = __dp4a(static_cast<int>(v[0]), static_cast<int>(wk[0][0]), s0);
                        acc[1] = __dp4a(static_cast<int>(v[1]), static_cast<int>(wk[0][1]), s1);
                        acc[2] = __dp4a(static_cast<int>(v[2]), static_cast<int>(wk[0][2]), s2);
                        acc[3] = __dp4a(static_cast<int>(v[3]), static_cast<int>(wk[0][3]), s3);
                    };
                    auto first = [&](int (&acc)[4], const uint32_t (&v)[4], int yo) {
                        if (yo == y_bot) {
                            first_at(acc, v, seed_mid + 8 * cp * 4);
                        } else {
                            acc[0] = __dp4a(static_cast<int>(v[0]), static_cast<int>(wk[0][0]), sm0);
                            acc[1] = __dp4a(static_cast<int>(v[1]), static_cast<int>(wk[0][1]), sm1);
                            acc[2] = __dp4a(static_cast<int>(v[2]), static_cast<int>(wk[0][2]), sm2);
                            acc[3] = __dp4a(static_cast<int>(v[3]), static_cast<int>(wk[0][3]), sm3);
                        }
                    };
                    const uint32_t seed_row0 =
                        seed_mid + (((top_pad && oy0 + y0 == 0) ? 4 : 0) | (y_bot == 0 ? 8 : 0)) * cp * 4;
                    // A operand row of pixel (y0 + yy, x): p = (y0 + yy) * TW + x; 128-byte rows, 16-byte
                    // chunks XOR-swizzled with (p & 7); K block kb = ch / 128 holds MT * 128 rows
                    uint32_t a_lin = a_tile + static_cast<uint32_t>(ch >> 7) * (MT * 16384) + (ch & 15) + (p0 << 7);
                    const uint32_t c16 = (ch >> 4) & 7;
                    uint32_t p = p0;
                    auto store = [&](const int (&acc)[4]) {
                        int t[4];
                        requant_pair<true>(acc[0], acc[1], mu[0], ba[0], t[0], t[1]);
                        requant_pair<true>(acc[2], acc[3], mu[1], ba[1], t[2], t[3]);
                        const uint32_t q4 = finish4<DMODE>(t, de, s_dlut, d_has_lut, dzp_m, dlut_lo, dlut_base);
                        const uint32_t addr = a_lin + (((c16 ^ p) & 7) << 4);
                        asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(q4) : "memory");
                        p += a.tw;
                        a_lin += a_row_step;
                    };
                    int accA[4], accB[4], accC[4];
                    uint32_t v[4];
                    if (S == 1) {
                        taps(v);
                        first_at(accA, v, seed_row0);
                        taps(v);
                        fu_dp4(accA, v, wk[1]), first(accB, v, 1);
                        for (int y = 0;; y += 3) {
                            taps(v);
                            fu_dp4(accA, v, wk[2]), fu_dp4(accB, v, wk[1]), first(accC, v, y + 2);
                            store(accA);
                            if (y + 1 >= rows_out) break;
                            taps(v);
                            fu_dp4(accB, v, wk[2]), fu_dp4(accC, v, wk[1]), first(accA, v, y + 3);
                            store(accB);
                            if (y + 2 >= rows_out) break;
                            taps(v);
                            fu_dp4(accC, v, wk[2]), fu_dp4(accA, v, wk[1]), first(accB, v, y + 4);
                            store(accC);
                            if (y + 3 >= rows_out) break;
                        }
                    } else {
                        taps(v);
                        first_at(accA, v, seed_row0);
                        for (int y = 0;; y += 2) {
                            taps(v);
                            fu_dp4(accA, v, wk[1]);
                            taps(v);
                            fu_dp4(accA, v, wk[2]), first(accB, v, y + 1);
                            store(accA);
                            if (y + 1 >= rows_out) break;
                            taps(v);
                            fu_dp4(accB, v, wk[1]);
                            taps(v);
                            fu_dp4(accB, v, wk[2]), first(accA, v, y + 2);
                            store(accB);
                            if (y + 2 >= rows_out) break;
                        }
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&halo_empty[stage]);
                if (++stage == a.stages) stage = 0, phase ^= 1;
            }
            // this warp's part of the A tile is written: generic-proxy stores before the tensor core's
            // async-proxy reads
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&a_full[buf]);
            // while the tensor core works on this tile, drain the previous one
            if (local > 0) epilogue_tile(local - 1);
        }
        if (local > 0) epilogue_tile(local - 1);
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}

struct FuPlan {
    int tw, th, mt, cc, rb, rpb, stages;
    int scols, nhalves;
    size_t smem;
};

static size_t fu_smem_bytes(const FuPlan &p, int cp, int bn, int k_blocks, int S)
{
    const int twi = S * (p.tw - 1) + 3, thi = S * (p.th - 1) + 3;
    const size_t stage = (static_cast<size_t>(twi) * thi * p.cc + 127) & ~static_cast<size_t>(127);
    const size_t stg_half = static_cast<size_t>(p.mt) * 128 * p.scols;  // multiple of 2 KB
    return 1024 + static_cast<size_t>(k_blocks) * bn * 128 + 2 * static_cast<size_t>(k_blocks) * p.mt * 16384 + 2 * stg_half +
           p.stages * stage + static_cast<size_t>(cp) * (3 + 2 + 16) * 4 + sizeof(FuEpiParams) + 256 +
           (2 * kFuMaxStages + 13) * sizeof(uint64_t) + 16;
}

// Tile shape, channel chunk and ring depth for one layer; false when the pair does not fit the kernel.
// Every (TW, TH) with TW * TH <= 256 pixels is scored by an instruction-count model of the two issue-bound
// halves -- depthwise: ~13 instructions per output, of which the input-row fetch + byte permutes (a fifth) grow
// with the input rows a thread reads per output row, divided by the share of consumer threads and of tile
// columns / rows that do real work; pointwise epilogue: ~5.6 per output over the share of the M128 lanes that
// are real pixels -- with a mild preference for thin halos and at least three halo stages.
static bool fu_plan(const b200_dwpw_desc *d, FuPlan *out)
{
    const b200_dwconv_desc &w = d->dw;
    const int S = w.stride_h;
    const int bn = (d->o + 15) / 16 * 16;
    if (bn > 256) return false;
    const int k_blocks = (w.cp + 127) / 128;
    const int mt_max = bn <= 128 ? 2 : 1;
    double best = -1;
    FuPlan bp = {};
    // developer override: SHL_B200_DWPW_TILE="tw,th" pins the tile shape (tools/run_pair.py sweeps it)
    int force_tw = 0, force_th = 0;
    if (const char *e = getenv("SHL_B200_DWPW_TILE")) sscanf(e, "%d,%d", &force_tw, &force_th);
    for (int tw = 4; tw <= 64; tw++) {
        if (tw > 4 && tw > w.ow) break;
        for (int th = 2; th <= 64; th++) {
            if (force_tw && (tw != force_tw || th != force_th)) continue;
            const int P = tw * th;
            const int mt = (P + 127) / 128;
            if (mt > mt_max) break;
            if (th > 2 && th > w.oh) break;
            if (S * (tw - 1) + 3 > 256 || S * (th - 1) + 3 > 256) continue;
            FuPlan p = {};
            p.tw = tw, p.th = th, p.mt = mt;
            // channel chunk: the widest whose items (columns x words) fit the consumer threads
            int cc = 128;
            while (cc > 16 && (tw * (cc / 4) > kFuDwThreads || cc / 2 >= w.cp)) cc /= 2;
            if (tw * (cc / 4) > kFuDwThreads) continue;
            p.cc = cc;
            int rb = kFuDwThreads / (tw * (cc / 4));
            while (rb > 1 && (th + rb - 1) / rb < 4) rb--;
            p.rpb = (th + rb - 1) / rb;
            p.rb = (th + p.rpb - 1) / p.rpb;
            p.scols = bn <= 16 ? 16 : (bn <= 32 ? 32 : (bn <= 64 ? 64 : 128));
            p.nhalves = (bn + 127) / 128;
            p.stages = kFuMaxStages;
            while (p.stages > 2 && fu_smem_bytes(p, w.cp, bn, k_blocks, S) > kFuSmemLimit) p.stages--;
            p.smem = fu_smem_bytes(p, w.cp, bn, k_blocks, S);
            if (p.smem > kFuSmemLimit) continue;
            const int xb = (w.ow + tw - 1) / tw, yb = (w.oh + th - 1) / th;
            const double col_util = static_cast<double>(w.ow) / (xb * tw), row_util = static_cast<double>(w.oh) / (yb * th);
            const int cchunks = (w.cp + cc - 1) / cc;
            const double busy = static_cast<double>(tw * (cc / 4) * p.rb) / kFuDwThreads * w.cp / (cchunks * cc);
            const double rows_norm = static_cast<double>(S * (p.rpb - 1) + 3) / (S * p.rpb);  // input rows per row needed
            // + the per-tile prologue of a thread (tile decode, constants, addresses: ~200 instructions per 4 channels)
            const double dw_cost = w.c * (13.0 * (0.2 * rows_norm + 0.8) + 50.0 / p.rpb) / (col_util * row_util * busy);
            const double lane_util = static_cast<double>(w.ow) * w.oh / (static_cast<double>(xb) * yb * mt * 128.0);
            const double pw_cost = d->o * (5.6 + 1.5 / mt) / lane_util;  // + the per-column parameters, loaded once per M128 pair
            const double halo = static_cast<double>(S * tw) * (S * th) / ((S * (tw - 1) + 3.0) * (S * (th - 1) + 3.0));
            double score = 1e4 / (dw_cost + pw_cost) * sqrt(sqrt(halo));
            if (p.stages < 3) score *= 0.9;
            if (score > best + 1e-9) best = score, bp = p;
        }
    }
    if (best < 0) return false;
    *out = bp;
    return true;
}

template <int S, int DMODE, int PMODE, bool MAGIC>
static int fu_launch(int grid, size_t smem, cudaStream_t stream, const CUtensorMap &ti, const CUtensorMap &tw,
                     const CUtensorMap &to, const DwPwArgs &args, int dev)
{
    static bool attr_set[64] = {};
    if (dev >= 0 && dev < 64 && !attr_set[dev]) {
        B200_CUDA_CHECK(cudaFuncSetAttribute(dwpw_kernel<S, DMODE, PMODE, MAGIC>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)kFuSmemLimit));
        attr_set[dev] = true;
    }
    B200_CUDA_CHECK(launch_kernel(dwpw_kernel<S, DMODE, PMODE, MAGIC>, dim3(grid), dim3(kFuThreads), smem, stream, ti, tw, to, args));
    return B200_OK;
}

}  // namespace b200

using namespace b200;

static bool dwpw_shape_ok(const b200_dwpw_desc *d)
{
    const b200_dwconv_desc &w = d->dw;
    return w.dtype == B200_I8 && w.wt_row3 && w.kh == 3 && w.kw == 3 && w.dil_h == 1 && w.dil_w == 1 &&
           w.stride_h == w.stride_w && (w.stride_h == 1 || w.stride_h == 2) && w.pad_top <= 1 && w.pad_left <= 1 &&
           (w.oh - 1) * w.stride_h - w.pad_top + 2 <= w.h && (w.ow - 1) * w.stride_w - w.pad_left + 2 <= w.w &&
           w.n > 0 && w.c > 0 && w.cp >= w.c && w.cp % 16 == 0 && d->o > 0 && d->ldw >= w.c && d->ldw % 16 == 0 &&
           d->ldo >= d->o && d->ldo % 16 == 0 && w.ep.mult && w.ep.badd && w.ep.ibias && d->ep.mult && d->ep.badd &&
           w.in && d->w && d->out;
}

extern "C" int b200_dwpw_supported(const b200_dwpw_desc *d)
{
    if (!d || !dwpw_shape_ok(d)) return 0;
    FuPlan p;
    return fu_plan(d, &p) ? 1 : 0;
}

// the tile plan as text (tools / tests): "tw th mt cc rb rpb stages scols nhalves smem"
extern "C" int b200_dwpw_plan_describe(const b200_dwpw_desc *d, char *buf, int buflen)
{
    FuPlan p;
    if (!d || !buf || buflen <= 0 || !dwpw_shape_ok(d) || !fu_plan(d, &p)) return 0;
    return snprintf(buf, buflen, "tw=%d th=%d mt=%d cc=%d rb=%d rpb=%d stages=%d scols=%d nhalves=%d smem=%zu", p.tw, p.th,
                    p.mt, p.cc, p.rb, p.rpb, p.stages, p.scols, p.nhalves, p.smem);
}

extern "C" int b200_dwpw_fused(const b200_dwpw_desc *d, void *stream)
{
    if (!d || !dwpw_shape_ok(d)) {
        set_error("b200_dwpw_fused: descriptor outside the fused kernel's domain (int8 3x3 depthwise, stride 1 / 2, pads <= 1)");
        return B200_ERR_UNSUPPORTED;
    }
    FuPlan p;
    if (!fu_plan(d, &p)) {
        set_error("b200_dwpw_fused: no tile plan fits shared memory (c=%d o=%d)", d->dw.c, d->o);
        return B200_ERR_UNSUPPORTED;
    }
    const b200_dwconv_desc &w = d->dw;
    const int S = w.stride_h;
    DwPwArgs a = {};
    a.n = w.n, a.cp = w.cp, a.h = w.h, a.w = w.w, a.oh = w.oh, a.ow = w.ow, a.pt = w.pad_top, a.pl = w.pad_left;
    a.tw = p.tw, a.th = p.th, a.twi = S * (p.tw - 1) + 3, a.thi = S * (p.th - 1) + 3;
    a.xbands = (w.ow + p.tw - 1) / p.tw, a.ybands = (w.oh + p.th - 1) / p.th;
    a.cc = p.cc;
    a.words_log2 = p.cc == 128 ? 5 : (p.cc == 64 ? 4 : (p.cc == 32 ? 3 : 2));
    a.cchunks = (w.cp + p.cc - 1) / p.cc;
    a.rb = p.rb, a.rpb = p.rpb;
    a.stage_bytes = a.twi * a.thi * p.cc;
    a.stage_stride = (a.stage_bytes + 127) & ~127;
    a.stages = p.stages;
    a.zp_in = w.zp_in;
    a.wrow = static_cast<const uint32_t *>(w.wt_row3);
    a.dep = make_epi(w.ep);
    a.o = d->o;
    a.bn = (d->o + 15) / 16 * 16;
    a.k_blocks = (w.cp + 127) / 128;
    a.kslices = (w.c + 31) / 32;
    a.mt = p.mt;
    a.idesc = umma_idesc(2 /*S32*/, 1 /*S8*/, 128, a.bn);
    a.a_buf_bytes = static_cast<uint32_t>(a.k_blocks) * p.mt * 16384;
    a.scols = p.scols, a.nhalves = p.nhalves;
    a.stg_half_bytes = static_cast<uint32_t>(p.mt) * 128 * p.scols;
    a.pep = make_epi(d->ep);

    const long long tiles = static_cast<long long>(w.n) * a.xbands * a.ybands;
    if (tiles >= (1ll << 31)) {
        set_error("b200_dwpw_fused: %lld tiles exceed the kernel's 32-bit tile index", tiles);
        return B200_ERR_UNSUPPORTED;
    }
    const int grid = static_cast<int>(tiles < sm_count() ? tiles : sm_count());
    {
        int q = grid;
        a.dxb = q % a.xbands;
        q /= a.xbands;
        a.dyb = q % a.ybands;
        a.db = q / a.ybands;
    }

    alignas(64) CUtensorMap ti, tw, to;
    int rc = encode_tmap_nhwc_u8_ex(&ti, w.in, w.n, w.h, w.w, w.cp, p.cc, a.twi, a.thi, 1, 0);
    if (rc) return rc;
    // pointwise weights [o][ldw], K-major: one 128-byte K block x bn rows per box (rows >= o and bytes >= c zero-filled)
    rc = encode_tmap_2d(&tw, 1, d->w, w.c, d->o, static_cast<uint64_t>(d->ldw), 128, a.bn);
    if (rc) return rc;
    // output tile: {scols channels, TW, TH, 1} pixels; clipped at the image edges and at the row pitch
    rc = encode_tmap_nhwc_u8_ex(&to, d->out, w.n, w.oh, w.ow, d->ldo, p.scols, p.tw, p.th, 1, p.scols >= 32 ? p.scols : 0);
    if (rc) return rc;

    int dev = 0;
    B200_CUDA_CHECK(cudaGetDevice(&dev));
    cudaStream_t s = (cudaStream_t)stream;
    // |acc + ibias| <= K * 2 * 128 * 127 < 2^22 lets the pointwise epilogue convert through the magic constant
    const bool magic = w.c <= 128;
    // a clamp followed by a table is folded into the table on the host (b200_opt); the kernel knows
    // "table only" and the generic clamp
    const int dmode = (w.ep.post_lut && w.ep.act == B200_ACT_NONE) ? EPI_LUT : EPI_GENERIC;
    const int pmode = (d->ep.post_lut && d->ep.act == B200_ACT_NONE) ? EPI_LUT : EPI_GENERIC;
#define B200_FU_CASE(SS, DM, PM)                                                                        \
    rc = magic ? fu_launch<SS, DM, PM, true>(grid, p.smem, s, ti, tw, to, a, dev)                       \
               : fu_launch<SS, DM, PM, false>(grid, p.smem, s, ti, tw, to, a, dev)
    if (S == 1) {
        if (dmode == EPI_LUT && pmode == EPI_LUT) B200_FU_CASE(1, EPI_LUT, EPI_LUT);
        else if (dmode == EPI_LUT) B200_FU_CASE(1, EPI_LUT, EPI_GENERIC);
        else if (pmode == EPI_LUT) B200_FU_CASE(1, EPI_GENERIC, EPI_LUT);
        else B200_FU_CASE(1, EPI_GENERIC, EPI_GENERIC);
    } else {
        if (dmode == EPI_LUT && pmode == EPI_LUT) B200_FU_CASE(2, EPI_LUT, EPI_LUT);
        else if (dmode == EPI_LUT) B200_FU_CASE(2, EPI_LUT, EPI_GENERIC);
        else if (pmode == EPI_LUT) B200_FU_CASE(2, EPI_GENERIC, EPI_LUT);
        else B200_FU_CASE(2, EPI_GENERIC, EPI_GENERIC);
    }
#undef B200_FU_CASE
    if (rc) return rc;
    B200_LAUNCH_CHECK();
    return B200_OK;
}
