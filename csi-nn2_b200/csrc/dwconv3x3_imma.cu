// dwconv3x3_imma.cu -- int8 depthwise 3x3 (stride 1 / 2, dilation 1) on pixel-major tensors with the taps
// accumulated by warp-level tensor-core MMAs (mma.sync.m16n8k32 s8, SASS IMMA.16832) over a TMA-fed halo tile.
//
// Why: the dp4a kernel (dwconv3x3_tma.cu) spends 13 issue slots per output and is bound by instruction issue, not by
// HBM.  5.25 of those slots are its tap arithmetic: three shared loads, six byte permutes that transpose pixel-major
// words into per-channel tap words, twelve dp4a -- per four outputs.  A depthwise convolution is a GEMM with a
// DIAGONAL weight matrix per tap: D[pixel][c] += sum_k A[pixel][k] * B[k][c] with k running over (two taps) x (the 16
// channels of a chunk) and B[(tap, k)][c] = w[tap][c] * (k == c).  15/16 of that MMA's multiplies hit zeros, but the
// warp-level tensor path is otherwise idle (tools/probes/imma_rate.cu: 2048 MAC per clock and SM, issuing beside the integer
// pipe), the A fragment of two taps is ONE ldmatrix.x4 of 2 x 16 pixels x 16 channels straight out of the halo tile (no
// transposition: the pixel-major layout IS the row-major A operand), and one instruction covers 16 pixels x 8 channels
// of two taps: 10 MMAs + 2 ldmatrix per 256 outputs = 0.4 issue slots per output for the taps (taps (ky, 0), (ky, 1) of an
// input row share an MMA; the kx = 2 column is paired vertically: (0, 2) + (1, 2), then (2, 2) with a zero half).  What is
// left is the requantise epilogue, the same code as everywhere (common.cuh).
//
// Tile: THI x TWI input pixels x CC channels (CC = 128 / 64 / 32 = the TMA swizzle span, so a pixel's 16-byte chunks
// are XOR-ed with its index and eight consecutive pixels of one chunk fall into eight different bank groups:
// ldmatrix runs conflict-free).  TWI is a multiple of 8 pixels, so the swizzle term of a (column, chunk) pair is the
// same on every row and a lane's ldmatrix address advances by a constant per row.  A warp owns one strip of 16
// output columns x 16 channels and slides down the tile: every input row is loaded once and feeds the three output rows
// it belongs to through rotating accumulator sets, as in the dp4a kernel.  Output column n of the two MMAs per fragment is mapped to the channels so that a thread ends up with four
// ADJACENT channels of two pixels: one 32-bit store per pixel.
//
// Zero-point padding: the TMA zero-fills.  Padded COLUMNS: the difference zp_in * (weights of the padded taps) sits in
// per-(column class, channel) accumulator seeds, which enter as the C operand of an output row's first MMA (as in
// dwconv3x3_tma.cu).  Padded ROWS (at most one above / below the image per tile): the consumers overwrite the image's
// columns of that tile row with zp_in before the first ldmatrix -- a row class per output row would cost a predicated
// seed reload in every row step.
//
// Measured (B200, batch 256): bit-exact; faster than the dp4a kernel for stride 2 on the larger maps,
// slower for stride 1 -- the m16n8k16 shape holds the tensor pipe exactly as long as m16n8k32 (a first version on k16 ran
// it at 69 % and lost everywhere), and with k32 the kernel sits at 46 % tensor / 55 % issue with the epilogue unchanged.
//
// Replaces shl_rvv_dwconv3x3s1_int8 / shl_rvv_dwconv3x3s2_int8 (source/thead_rvv/int8/depthwise_convolution_3x3_int8.c:31,244);
// semantics shl_ref_depthwise_conv2d_quant (source/reference/convolution.c:416).
#include <stdlib.h>

#include "common.cuh"

namespace b200 {

constexpr int kImStagesMax = 3;
constexpr int kImWarps = 8;  // consumer warps per CTA = strips x 16-channel chunks of a tile

struct DwImmaArgs {
    int n, c, cp, h, w, oh, ow, pt, pl;
    int th;        // output rows per tile
    int sp;        // output columns a strip owns (it computes 16)
    int row_bytes; // TWI * CC
    int stages;
    int ybands, xbands, cchunks;
    int dyb, db;  // the grid's stride over (row band, image) as two digits
    int stage_bytes, stage_stride;
    const uint32_t *wrow;  // [3 (ky)][cp] words: (w[ky][0][c], w[ky][1][c], w[ky][2][c], 0)
    int8_t *out;
    int zp_in;
    EpiScalars ep;
};

// This CTA's walk over the tile space: the channel chunk AND the column band are blockIdx.x % (cchunks * xbands) for good
// (the grid is a multiple of that), so everything that depends on them -- per-channel constants, B fragments, seeds, the
// threads' output columns and their border classes -- is set up once; (row band, image) advance by a fixed stride, applied as
// two digits with a carry.
struct ImmaWalk {
    int cc, xb, yb, b;
    __device__ __forceinline__ explicit ImmaWalk(const DwImmaArgs &a)
    {
        const uint32_t lanes = a.cchunks * a.xbands;
        const uint32_t l = blockIdx.x % lanes;
        cc = l % a.cchunks;
        xb = l / a.cchunks;
        const uint32_t q = blockIdx.x / lanes;
        yb = q % a.ybands;
        b = q / a.ybands;
    }
    __device__ __forceinline__ void next(const DwImmaArgs &a)
    {
        yb += a.dyb;
        if (yb >= a.ybands) yb -= a.ybands, b++;
        b += a.db;
    }
};

__device__ __forceinline__ void ldsm_x4(uint32_t &r0, uint32_t &r1, uint32_t &r2, uint32_t &r3, uint32_t addr)
{
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x2(uint32_t &r0, uint32_t &r1, uint32_t addr)
{
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0, %1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr));
}
// D = A (16 pixels x 32 = two taps x 16 channels, s8) * B (32 x 8, s8) + C.  (The k16 shape occupies the tensor pipe as
// long as this one -- measured: a first version of this kernel on m16n8k16 ran the pipe at 69 % and lost to dp4a.)
// B travels as one 64-bit value (unpacked inside the asm block, which costs nothing): the register allocator then keeps
// the (tap, zero) pairs resident instead of re-making them from a zero register before every MMA
__device__ __forceinline__ void imma16832(int (&d)[4], const int (&c)[4], const uint32_t (&a)[4], uint64_t b)
{
    asm("{\n\t.reg .b32 blo, bhi;\n\tmov.b64 {blo, bhi}, %8;\n\t"
        "mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {blo, bhi}, {%9, %10, %11, %12};\n\t}"
        : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "l"(b), "r"(c[0]), "r"(c[1]), "r"(c[2]), "r"(c[3]));
}

template <int S, int CC, int MODE>
__global__ void __launch_bounds__(kImWarps * 32 + 32, 2) dw3x3_imma_kernel(const __grid_constant__ CUtensorMap tmap, const DwImmaArgs a)
{
    constexpr int NCH = CC / 16;          // 16-channel chunks of a tile
    constexpr int kConsumers = kImWarps * 32;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // the swizzle pattern is a function of the absolute shared-memory address: slots start on 1024-byte boundaries
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    __shared__ uint64_t full_bar[kImStagesMax], empty_bar[kImStagesMax];
    __shared__ uint8_t s_lut[256];
    __shared__ __align__(16) int s_seed[4 * CC];  // [column class][channel]

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    pdl_launch_dependents();
    if (tid == 0) {
        tma_prefetch_desc(&tmap);
        for (int i = 0; i < a.stages; i++) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], kImWarps);
        }
        mbar_fence_init();
    }
    if (a.ep.post_lut != nullptr && tid < 256) s_lut[tid] = static_cast<uint8_t>(a.ep.post_lut[tid]);
    __syncthreads();

    if (warp == kImWarps) {
        // ===== TMA producer =====
        if (elect_one()) {
            pdl_wait();  // the input is the predecessor's output
            int stage = 0;
            uint32_t phase = 0;
            ImmaWalk tw(a);
            for (; tw.b < a.n; tw.next(a)) {
                mbar_wait(&empty_bar[stage], phase ^ 1);
                mbar_expect_tx(&full_bar[stage], a.stage_bytes);
                tma_load_4d(smem + static_cast<size_t>(stage) * a.stage_stride, &tmap, &full_bar[stage], tw.cc * CC,
                            tw.xb * (a.sp * (kImWarps / NCH)) * S - a.pl, tw.yb * a.th * S - a.pt, tw.b);
                if (++stage == a.stages) {
                    stage = 0;
                    phase ^= 1;
                }
            }
        }
        return;
    }

    // ===== consumers =====
    const int chunk = warp % NCH, strip = warp / NCH;
    const int g = lane >> 2, q = lane & 3;
    const bool has_lut = a.ep.post_lut != nullptr;
    const int zp_m = a.ep.zp_out - kMagicI;
    const int lut_base = static_cast<int>(smem_u32(s_lut));
    int lut_lo = kMagicI - a.ep.zp_out - 128 - lut_base;
    asm("mov.b32 %0, %0;" : "+r"(lut_lo));
    ImmaWalk walk(a);
    const int cc = walk.cc;
    const uint32_t row_bytes = a.row_bytes;

    // ldmatrix lane addresses on row 0 of a tile (relative to the slot): lane l supplies row (l & 7) of matrix (l >> 3).
    // F load: matrices (kx 0, pixels 0-7), (kx 0, pixels 8-15), (kx 1, 0-7), (kx 1, 8-15) = the A operand of the MMA over the
    // taps kx 0 and 1 of one input row.  G load: (kx 2, pixels 0-7), (kx 2, 8-15) of this row and of the NEXT row = the A
    // operand over the taps (ky, 2), (ky + 1, 2).
    auto swz = [&](int x) {
        const uint32_t lin = static_cast<uint32_t>(x) * CC + chunk * 16;
        return lin ^ (((lin >> 7) & (CC / 16 - 1)) << 4);
    };
    const int m = lane >> 3;
    const int xl = strip * a.sp * S + S * ((lane & 7) + 8 * (m & 1));
    const uint32_t offF = swz(xl + (m >> 1));
    const uint32_t offG = swz(xl + 2) + (m >> 1) * row_bytes;

    // B fragments: column n = g of MMA nh holds channel 4 * (g >> 1) + 2 * nh + (g & 1) of the chunk, so that this
    // thread's accumulators (columns 2q, 2q + 1 of both MMAs) are the four adjacent channels 4q .. 4q + 3.
    // K halves: Bf[ky] = (w[ky][0], w[ky][1]); Bg01 = (w[0][2], w[1][2]); Bg12 = (w[1][2], w[2][2]); Bg2 = (w[2][2], 0)
    uint64_t Bf[3][2], Bg01[2], Bg2[2];
    uint32_t zero;  // opaque to the compiler: the (w, 0) pairs stay resident instead of being re-made before every MMA
    asm volatile("mov.u32 %0, 0;" : "=r"(zero));
    {
        const int cgrp = cc * CC + chunk * 16 + 4 * (g >> 1);
#pragma unroll
        for (int nh = 0; nh < 2; nh++) {
            const int j = 2 * nh + (g & 1);
            uint32_t w2[3];
#pragma unroll
            for (int ky = 0; ky < 3; ky++) {
                const uint32_t wv = (cgrp + j < a.cp && q == (g >> 1)) ? __ldg(a.wrow + ky * a.cp + cgrp + j) : 0u;
                Bf[ky][nh] = f2_pack_bits((wv & 0xFFu) << (8 * j), ((wv >> 8) & 0xFFu) << (8 * j));
                w2[ky] = ((wv >> 16) & 0xFFu) << (8 * j);
            }
            Bg01[nh] = f2_pack_bits(w2[0], w2[1]);
            Bg2[nh] = f2_pack_bits(w2[2], zero);
        }
    }
    const int ch = cc * CC + chunk * 16 + 4 * q;  // first of this thread's four output channels
    const bool ch_ok = ch < a.cp;
    uint64_t mu[2], ba[2];
    {
        const int chs = ch_ok ? ch : 0;
        const float4 m4 = __ldg(reinterpret_cast<const float4 *>(a.ep.mult + chs));
        const float4 b4 = __ldg(reinterpret_cast<const float4 *>(a.ep.badd + chs));
        mu[0] = f2_pack(m4.x, m4.y), mu[1] = f2_pack(m4.z, m4.w);
        ba[0] = f2_pack(b4.x, b4.y), ba[1] = f2_pack(b4.z, b4.w);
        // accumulator seeds per column class (1 = kx 0 hangs over the left edge, 2 = kx 2 over the right edge): ibias +
        // kMagicI + zp_in * (the weights of the column's padded taps, all three rows).  Padded ROWS are patched to zp_in in
        // the tile itself (below), over the image's columns only, so the two corrections never overlap.
        for (int i = tid; i < 4 * CC; i += kConsumers) {
            const int c = i % CC, cls = i / CC;
            const int cg = cc * CC + c;
            int v = 0;
            if (cg < a.cp) {
                int padsum = 0;
#pragma unroll
                for (int ky = 0; ky < 3; ky++) {
                    const uint32_t wv = __ldg(a.wrow + ky * a.cp + cg);
                    if (cls & 1) padsum += static_cast<int8_t>(wv);
                    if (cls & 2) padsum += static_cast<int8_t>(wv >> 16);
                }
                v = __ldg(a.ep.ibias + cg) + kMagicI + a.zp_in * padsum;
            }
            s_seed[i] = v;
        }
        asm volatile("bar.sync 1, %0;" ::"n"(kConsumers) : "memory");
    }
    const size_t orow = static_cast<size_t>(a.ow) * a.cp;
    const uint32_t zpw = 0x01010101u * static_cast<uint32_t>(a.zp_in & 0xFF);
    const int xband = a.sp * (kImWarps / NCH);  // output columns of a tile
    const int xb = walk.xb;
    // this thread's two output pixels: columns g and g + 8 of the strip (fixed for the CTA's life, like the chunk)
    const int ox0 = xb * xband + strip * a.sp + g;
    const int ox1 = ox0 + 8;
    const bool ok0 = ch_ok && g < a.sp && ox0 < a.ow;
    const bool ok1 = ch_ok && g + 8 < a.sp && ox1 < a.ow;
    // seeds in the layout of the two MMAs' C operands: sq[nh] = channels 2nh, 2nh + 1 of pixel g, then of pixel g + 8
    int sq[2][4];
    {
        const int col0 = ((ox0 * S - a.pl < 0) ? 1 : 0) | ((ox0 * S - a.pl + 2 >= a.w) ? 2 : 0);
        const int col1 = ((ox1 * S - a.pl < 0) ? 1 : 0) | ((ox1 * S - a.pl + 2 >= a.w) ? 2 : 0);
        const int *s0 = s_seed + col0 * CC + chunk * 16 + 4 * q, *s1 = s_seed + col1 * CC + chunk * 16 + 4 * q;
        sq[0][0] = s0[0], sq[0][1] = s0[1], sq[1][0] = s0[2], sq[1][1] = s0[3];
        sq[0][2] = s1[0], sq[0][3] = s1[1], sq[1][2] = s1[2], sq[1][3] = s1[3];
    }
    int8_t *const pcol = a.out + static_cast<size_t>(ox0) * a.cp + ch;
    // the image's columns inside a tile row, as 16-byte chunks (for the zero-point patch of padded rows)
    const int x_start = xb * xband * S - a.pl;            // image column of tile column 0
    const int tx0 = max(0, -x_start);
    const int per_row = (min(static_cast<int>(row_bytes / CC), a.w - x_start) - tx0) * (CC / 16);
    int stage = 0;
    uint32_t phase = 0;

    for (; walk.b < a.n; walk.next(a)) {
        const int oy0 = walk.yb * a.th;
        const int rows_out = min(a.th, a.oh - oy0);
        int8_t *po0 = pcol + (static_cast<size_t>(walk.b) * a.oh + oy0) * orow;
        int8_t *po1 = po0 + 8 * a.cp;

        pdl_wait();  // the output buffer may alias a tensor the predecessor still reads
        mbar_wait(&full_bar[stage], phase);
        uint8_t *tile_p = smem + static_cast<size_t>(stage) * a.stage_stride;
        const uint32_t tile = smem_u32(tile_p);

        // rows of the tile above / below the image were zero-filled by the TMA; the contract wants zp_in there, over the
        // image's columns (a pixel's CC bytes are CC bytes whatever the swizzle).  At most one row each side (host-checked).
        {
            const int iy0 = oy0 * S - a.pt;                       // image row of tile row 0
            const int rows_in = S * (rows_out - 1) + 3;           // tile rows the valid outputs read
            const bool top = iy0 < 0, bot = iy0 + rows_in > a.h;
            if ((top || bot) && zpw != 0u) {
                if (per_row > 0) {
                    const uint4 z4 = make_uint4(zpw, zpw, zpw, zpw);
                    for (int i = tid; i < 2 * per_row; i += kConsumers) {
                        const int which = i >= per_row;
                        if (which ? !bot : !top) continue;
                        const int r = which ? a.h - iy0 : 0;      // the first row past the image / the row above it
                        const int k = which ? i - per_row : i;
                        *reinterpret_cast<uint4 *>(tile_p + static_cast<size_t>(r) * row_bytes + (tx0 * CC) + k * 16) = z4;
                    }
                }
                asm volatile("bar.sync 1, %0;" ::"n"(kConsumers) : "memory");
            }
        }

        uint32_t aF = tile + offF, aG = tile + offG;
        uint32_t Af[4], Ag[4];
        auto load_f = [&]() {
            ldsm_x4(Af[0], Af[1], Af[2], Af[3], aF);
            aF += row_bytes;
        };
        auto load_g = [&]() {
            ldsm_x4(Ag[0], Ag[1], Ag[2], Ag[3], aG);
            aG += S * row_bytes;
        };
        auto mm_f = [&](int (&acc)[2][4], const int ky) {
#pragma unroll
            for (int nh = 0; nh < 2; nh++) imma16832(acc[nh], acc[nh], Af, Bf[ky][nh]);
        };
        // a fresh output row: its seeds are the C operand of the MMA over the taps (0, 0), (0, 1); then (0, 2), (1, 2)
        auto first = [&](int (&acc)[2][4]) {
#pragma unroll
            for (int nh = 0; nh < 2; nh++) {
                imma16832(acc[nh], sq[nh], Af, Bf[0][nh]);
                imma16832(acc[nh], acc[nh], Ag, Bg01[nh]);
            }
        };
        // the last input row of an output row: taps (2, 0), (2, 1) and (2, 2)
        auto last = [&](int (&acc)[2][4]) {
#pragma unroll
            for (int nh = 0; nh < 2; nh++) {
                imma16832(acc[nh], acc[nh], Af, Bf[2][nh]);
                imma16832(acc[nh], acc[nh], Ag, Bg2[nh]);
            }
        };
        auto store = [&](const int (&acc)[2][4]) {
            int t0[4], t1[4];
            requant_pair<true>(acc[0][0], acc[0][1], mu[0], ba[0], t0[0], t0[1]);
            requant_pair<true>(acc[1][0], acc[1][1], mu[1], ba[1], t0[2], t0[3]);
            requant_pair<true>(acc[0][2], acc[0][3], mu[0], ba[0], t1[0], t1[1]);
            requant_pair<true>(acc[1][2], acc[1][3], mu[1], ba[1], t1[2], t1[3]);
            const uint32_t w0 = finish4<MODE>(t0, a.ep, s_lut, has_lut, zp_m, lut_lo, lut_base);
            const uint32_t w1 = finish4<MODE>(t1, a.ep, s_lut, has_lut, zp_m, lut_lo, lut_base);
            if (ok0) *reinterpret_cast<uint32_t *>(po0) = w0;
            if (ok1) *reinterpret_cast<uint32_t *>(po1) = w1;
            po0 += orow, po1 += orow;
        };

        // one input row of the stride-1 walk: last row of `lst`, middle row of `mid`, first row of `nw`.  The six MMAs over
        // the F operand are independent; the four over G each continue an accumulator started at least four MMAs earlier
        // (a dependent MMA issued back to back waits out the whole tensor latency).
        auto step1 = [&](int (&lst)[2][4], int (&mid)[2][4], int (&nw)[2][4]) {
            load_f(), load_g();
#pragma unroll
            for (int nh = 0; nh < 2; nh++) imma16832(lst[nh], lst[nh], Af, Bf[2][nh]);
#pragma unroll
            for (int nh = 0; nh < 2; nh++) imma16832(mid[nh], mid[nh], Af, Bf[1][nh]);
#pragma unroll
            for (int nh = 0; nh < 2; nh++) imma16832(nw[nh], sq[nh], Af, Bf[0][nh]);
#pragma unroll
            for (int nh = 0; nh < 2; nh++) imma16832(lst[nh], lst[nh], Ag, Bg2[nh]);
#pragma unroll
            for (int nh = 0; nh < 2; nh++) imma16832(nw[nh], nw[nh], Ag, Bg01[nh]);
            store(lst);
        };
        int accA[2][4], accB[2][4], accC[2][4];
        if (S == 1) {
            // input row r: first row of output row r, middle row of r - 1, last row of r - 2
            load_f(), load_g();
            first(accA);
            load_f(), load_g();
            mm_f(accA, 1), first(accB);
            for (int y = 0;; y += 3) {
                step1(accA, accB, accC);
                if (y + 1 >= rows_out) break;
                step1(accB, accC, accA);
                if (y + 2 >= rows_out) break;
                step1(accC, accA, accB);
                if (y + 3 >= rows_out) break;
            }
        } else {
            // output row y reads input rows 2y (first), 2y + 1 (middle), 2y + 2 (last = first of row y + 1)
            load_f(), load_g();
            first(accA);
            for (int y = 0;; y += 2) {
                load_f();
                mm_f(accA, 1);
                load_f(), load_g();
                last(accA), first(accB);
                store(accA);
                if (y + 1 >= rows_out) break;
                load_f();
                mm_f(accB, 1);
                load_f(), load_g();
                last(accB), first(accA);
                store(accB);
                if (y + 2 >= rows_out) break;
            }
        }
        (void)accC;

        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[stage]);
        if (++stage == a.stages) {
            stage = 0;
            phase ^= 1;
        }
    }
}

template <int S, int CC>
static int launch_imma(int mode, int grid, size_t smem, cudaStream_t s, const CUtensorMap &tm, const DwImmaArgs &a, int dev)
{
#define B200_DWI_CASE(M)                                                                                   \
    case M: {                                                                                              \
        static bool attr[64] = {};                                                                         \
        if (dev >= 0 && dev < 64 && !attr[dev]) {                                                          \
            B200_CUDA_CHECK(cudaFuncSetAttribute(dw3x3_imma_kernel<S, CC, M>,                              \
                                                 cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); \
            attr[dev] = true;                                                                              \
        }                                                                                                  \
        B200_CUDA_CHECK(launch_kernel(dw3x3_imma_kernel<S, CC, M>, dim3(grid), dim3(kImWarps * 32 + 32), smem, s, tm, a)); \
        break;                                                                                             \
    }
    switch (mode) {
        B200_DWI_CASE(EPI_PLAIN)
        B200_DWI_CASE(EPI_RELU)
        B200_DWI_CASE(EPI_RELU6)
        B200_DWI_CASE(EPI_LUT)
        default:
            B200_DWI_CASE(EPI_GENERIC)
    }
#undef B200_DWI_CASE
    return B200_OK;
}

}  // namespace b200

using namespace b200;

// called by b200_dwconv2d (dwconv.cu) for int8 3x3, dilation 1, stride 1 / 2, pads <= 1; *handled = 0 when the shape is
// left to the dp4a kernel (channel count not a multiple of 32, maps narrower than a useful strip)
int b200_dwconv3x3_imma_launch(const b200_dwconv_desc *d, const void *wrow, void *stream, int *handled)
{
    *handled = 0;
    // SHL_B200_DW_IMMA: 0 = never, 2 = wherever the kernel applies, default (1) = where it measured faster than the dp4a kernel
    // on B200 (batch 256): stride 2 on maps that leave at least 28 output columns (64 x 112^2 -> 56^2: 55.3 vs 63.5 us,
    // 128 x 56^2 -> 28^2: 35.0 vs 37.8 us).  Stride 1 loses (128 x 56^2: 86 vs 76 us, 512 x 14^2: 28.7 vs 26.6 us): ten MMAs
    // per 256 outputs keep the tensor pipe 46 % busy and the same requantise epilogue still has to issue beside them.
    const int enabled = getenv("SHL_B200_DW_IMMA") ? atoi(getenv("SHL_B200_DW_IMMA")) : 1;
    if (!enabled) return B200_OK;
    const int S = d->stride_h;
    if (enabled == 1 && !(S == 2 && d->ow >= 28 && d->cp <= 64)) return B200_OK;
    const int CC = d->cp % 128 == 0 ? 128 : (d->cp % 64 == 0 ? 64 : (d->cp % 32 == 0 ? 32 : 0));
    if (!CC || d->ow < 12) return B200_OK;
    const int ns = kImWarps / (CC / 16);  // strips per tile
    // columns per strip: 16, or 14 when that wastes fewer of the 16 computed columns (the 14 / 28 / 56 / 112 wide maps)
    auto waste = [&](int sp) { const int xb = (d->ow + sp * ns - 1) / (sp * ns); return (double)d->ow / (xb * ns * 16.0); };
    const int sp = waste(14) > waste(16) + 1e-9 ? 14 : 16;
    const int twi = (S * (sp * ns - 1) + 3 + 7) & ~7;
    if (twi > 256) return B200_OK;
    const int row_bytes = twi * CC;
    static const int slot_kb = getenv("SHL_B200_DWI_SLOT_KB") ? atoi(getenv("SHL_B200_DWI_SLOT_KB")) : 0;
    static const int stages_env = getenv("SHL_B200_DWI_STAGES") ? atoi(getenv("SHL_B200_DWI_STAGES")) : 0;
    static const int ctas_per_sm = getenv("SHL_B200_DWI_CTAS") ? atoi(getenv("SHL_B200_DWI_CTAS")) : 2;
    int stages = stages_env ? stages_env : (S == 1 ? 3 : 2);
    if (stages > kImStagesMax) stages = kImStagesMax;
    if (stages < 2) stages = 2;
    const int slot = (slot_kb ? slot_kb : (S == 1 ? 32 : 46)) * 1024;
    int thi_max = slot / row_bytes;
    if (thi_max > 256) thi_max = 256;
    int th = S == 1 ? thi_max - 2 : (thi_max - 1) / 2;
    if (th < 1) return B200_OK;
    if (th > d->oh) th = d->oh;
    int ybands = (d->oh + th - 1) / th;
    th = (d->oh + ybands - 1) / ybands;
    const int xbands = (d->ow + sp * ns - 1) / (sp * ns);
    const int cchunks = d->cp / CC;
    {
        const long long per_band = static_cast<long long>(d->n) * xbands * cchunks;
        while (per_band * ybands < sm_count() && th > 4) {
            th = (th + 1) / 2;
            ybands = (d->oh + th - 1) / th;
            th = (d->oh + ybands - 1) / ybands;
        }
    }
    const int thi = S * (th - 1) + 3;

    DwImmaArgs a;
    a.n = d->n, a.c = d->c, a.cp = d->cp, a.h = d->h, a.w = d->w, a.oh = d->oh, a.ow = d->ow;
    a.pt = d->pad_top, a.pl = d->pad_left, a.th = th, a.sp = sp, a.row_bytes = row_bytes, a.stages = stages;
    a.ybands = ybands, a.xbands = xbands, a.cchunks = cchunks;
    a.stage_bytes = thi * row_bytes;
    a.stage_stride = (a.stage_bytes + 1023) & ~1023;
    a.wrow = static_cast<const uint32_t *>(wrow);
    a.out = static_cast<int8_t *>(d->out);
    a.zp_in = d->zp_in;
    a.ep = make_epi(d->ep);

    const long long tiles = static_cast<long long>(d->n) * ybands * xbands * cchunks;
    if (tiles >= (1ll << 31) || static_cast<long long>(d->n) * d->oh * d->ow * d->cp >= (1ll << 32)) return B200_OK;

    alignas(64) CUtensorMap tm;
    int rc = encode_tmap_nhwc_u8_ex(&tm, d->in, d->n, d->h, d->w, d->cp, CC, twi, thi, 1, CC);
    if (rc) return rc;

    const int lanes = cchunks * xbands;
    long long cap = static_cast<long long>(sm_count()) * ctas_per_sm;
    if (cap > lanes) cap -= cap % lanes;
    if (cap < lanes) cap = lanes;
    const int grid = static_cast<int>(tiles < cap ? tiles : cap);  // tiles is a multiple of lanes
    {
        const int qd = grid / lanes;
        a.dyb = qd % ybands;
        a.db = qd / ybands;
    }
    // reads past a slot, all multiplied by zero weights or feeding outputs that are never stored, but they must stay inside
    // the allocation: the G load of a tile's last row touches the row after it, the strips' idle columns up to 3 pixels more
    const size_t smem = static_cast<size_t>(stages) * a.stage_stride + 1024 + row_bytes + 1024;
    int mode;
    if (d->ep.post_lut)
        mode = d->ep.act == B200_ACT_NONE ? EPI_LUT : EPI_GENERIC;
    else
        mode = d->ep.act == B200_ACT_NONE ? EPI_PLAIN : (d->ep.act == B200_ACT_RELU ? EPI_RELU : EPI_RELU6);
    int dev = 0;
    B200_CUDA_CHECK(cudaGetDevice(&dev));
    cudaStream_t s = (cudaStream_t)stream;
    if (S == 1)
        rc = CC == 128 ? launch_imma<1, 128>(mode, grid, smem, s, tm, a, dev)
                       : (CC == 64 ? launch_imma<1, 64>(mode, grid, smem, s, tm, a, dev) : launch_imma<1, 32>(mode, grid, smem, s, tm, a, dev));
    else
        rc = CC == 128 ? launch_imma<2, 128>(mode, grid, smem, s, tm, a, dev)
                       : (CC == 64 ? launch_imma<2, 64>(mode, grid, smem, s, tm, a, dev) : launch_imma<2, 32>(mode, grid, smem, s, tm, a, dev));
    if (rc) return rc;
    B200_LAUNCH_CHECK();
    *handled = 1;
    return B200_OK;
}
