// common.cuh -- sm_100a PTX wrappers and shared helpers for the b200nn kernels.
// Everything here is hand-written inline PTX (mbarrier, TMA, tcgen05/TMEM); no CUTLASS.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "b200nn.h"

namespace b200 {

// ---- host side error plumbing -------------------------------------------------
void set_error(const char *fmt, ...);
void count_launch(int n = 1);
#define B200_CUDA_CHECK(expr)                                                              \
    do {                                                                                   \
        cudaError_t _e = (expr);                                                           \
        if (_e != cudaSuccess) {                                                           \
            b200::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr,                  \
                            cudaGetErrorString(_e));                                       \
            return B200_ERR_CUDA;                                                          \
        }                                                                                  \
    } while (0)
#define B200_LAUNCH_CHECK()                                                                \
    do {                                                                                   \
        cudaError_t _e = cudaGetLastError();                                               \
        if (_e != cudaSuccess) {                                                           \
            b200::set_error("%s:%d: kernel launch -> %s", __FILE__, __LINE__,              \
                            cudaGetErrorString(_e));                                       \
            return B200_ERR_CUDA;                                                          \
        }                                                                                  \
        b200::count_launch();                                                              \
    } while (0)

int sm_count();
bool pdl_enabled();  // runtime.cu: programmatic dependent launch when SHL_B200_PDL is set (opt-in, see there)

// Every kernel of the library is launched through this: with the programmatic-stream-serialization
// attribute (SHL_B200_PDL=1) a kernel may be scheduled while its predecessor in the stream (or CUDA graph) is still
// draining, run its prologue (barrier init, TMEM allocation, weight / table staging -- constants
// written at session setup) and then block in pdl_wait() until the predecessor has completed and
// its writes are visible.  Kernels call pdl_launch_dependents() first thing, so the chain never
// serialises on launch latency.  All accesses to activations (reads and writes: the arena reuses
// buffers) come after pdl_wait().
template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel_cluster(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                         int cluster_x, Args &&...args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute at[2];
    int n = 0;
    if (pdl_enabled()) {
        at[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[n].val.programmaticStreamSerializationAllowed = 1;
        n++;
    }
    if (cluster_x > 1) {  // thread-block cluster along x: CTAs 2j, 2j+1 share TMA multicast loads
        at[n].id = cudaLaunchAttributeClusterDimension;
        at[n].val.clusterDim.x = cluster_x;
        at[n].val.clusterDim.y = 1;
        at[n].val.clusterDim.z = 1;
        n++;
    }
    cfg.attrs = at;
    cfg.numAttrs = n;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                 Args &&...args)
{
    return launch_kernel_cluster(kern, grid, block, smem, stream, 1, static_cast<Args &&>(args)...);
}
// 2-D tiled tensor map, SWIZZLE_128B, zero fill out of bounds (runtime.cu)
int encode_tmap_nhwc_u8(CUtensorMap *map, const void *base, int n, int h, int w, int cp, int box_c,
                        int box_w, int box_h);
int encode_tmap_nhwc_u8_nb(CUtensorMap *map, const void *base, int n, int h, int w, int cp, int box_c,
                           int box_w, int box_h, int box_n);  // the same with box_n images per box
int encode_tmap_nhwc_u8_sw128(CUtensorMap *map, const void *base, int n, int h, int w, int cp, int box_w, int box_h,
                              int box_n);
int encode_tmap_nhwc_u8_ex(CUtensorMap *map, const void *base, int n, int h, int w, int cp, int box_c, int box_w,
                           int box_h, int box_n, int swizzle_bytes);
int encode_tmap_im2col_u8(CUtensorMap *map, const void *base, int n, int h, int w, int c, int cp, int lower_w, int lower_h,
                          int upper_w, int upper_h, int stride_w, int stride_h, int chans, int pixels, int swizzle_bytes);
int encode_tmap_2d(CUtensorMap *map, int elem_bytes, const void *base, uint64_t inner,
                   uint64_t outer, uint64_t pitch_bytes, uint32_t box_inner, uint32_t box_outer,
                   int swizzle_bytes = 128);

// ---- device helpers -------------------------------------------------------------
__device__ __forceinline__ void pdl_launch_dependents()
{
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
__device__ __forceinline__ void pdl_wait()
{
    asm volatile("griddepcontrol.wait;" ::: "memory");
}

__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ bool elect_one()
{
    uint32_t pred = 0;
    asm volatile(
        "{\n\t"
        ".reg .pred P;\n\t"
        "elect.sync _|P, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t"
        "}\n"
        : "=r"(pred));
    return pred != 0;
}

// mbarrier ------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// the same wait with a suspend-time hint: a failed try_wait parks the thread in hardware (it is woken when the
// phase completes) instead of returning after ~20 cycles, so a waiting warp stops taking issue slots from the
// warps that work (ncu on the fused block kernel: the three single-thread service warps alone spent 14 % of
// the SM's issue slots in SYNCS / BRA / YIELD spin loops)
__device__ __forceinline__ void mbar_wait_parked(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, %2;\n\t"
        "@P1 bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity), "r"(0x989680)
        : "memory");
}

// TMA -----------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap *m)
{
    asm volatile("prefetch.tensormap [%0];\n" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void *smem_dst, const CUtensorMap *m, uint64_t *bar,
                                            int c0, int c1)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4}], [%2];\n" ::"r"(smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
// the same load delivered to the same shared-memory offset (and mbarrier) of every CTA of the cluster
// whose bit is set in cta_mask: one L2 read feeds several SMs
__device__ __forceinline__ void tma_load_2d_multicast(void *smem_dst, const CUtensorMap *m, uint64_t *bar, int c0,
                                                      int c1, uint16_t cta_mask)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
        " [%0], [%1, {%3, %4}], [%2], %5;\n" ::"r"(smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(cta_mask)
        : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void tma_load_4d(void *smem_dst, const CUtensorMap *m, uint64_t *bar,
                                            int c0, int c1, int c2, int c3)
{
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5, %6}], [%2];\n" ::"r"(smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
// im2col-mode load (tensor map from encode_tmap_im2col_u8): `pixels` output pixels starting at base pixel
// (w, h, n) x the map's channel count starting at c, of the filter tap at offsets (woff, hoff)
__device__ __forceinline__ void tma_load_im2col_4d(void *smem_dst, const CUtensorMap *m, uint64_t *bar, int c, int w, int h,
                                                   int n, uint16_t woff, uint16_t hoff)
{
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};\n" ::"r"(smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c), "r"(w), "r"(h), "r"(n), "h"(woff), "h"(hoff)
        : "memory");
}
// the same im2col load delivered to the same shared-memory offset (and mbarrier) of every CTA of the cluster in cta_mask
__device__ __forceinline__ void tma_load_im2col_4d_multicast(void *smem_dst, const CUtensorMap *m, uint64_t *bar, int c, int w,
                                                             int h, int n, uint16_t woff, uint16_t hoff, uint16_t cta_mask)
{
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes.multicast::cluster"
        " [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8}, %9;\n" ::"r"(smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c), "r"(w), "r"(h), "r"(n), "h"(woff), "h"(hoff),
        "h"(cta_mask)
        : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap *m, const void *smem_src, int c0,
                                             int c1)
{
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];\n" ::
                     "l"(reinterpret_cast<uint64_t>(m)),
                 "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tma_store_commit()
{
    asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
}
template <int N>
__device__ __forceinline__ void tma_store_wait_read()
{
    asm volatile("cp.async.bulk.wait_group.read %0;\n" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void tma_store_wait()
{
    asm volatile("cp.async.bulk.wait_group %0;\n" ::"n"(N) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem()
{
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
}

// tcgen05 / TMEM ------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t *dst_smem, uint32_t ncols)
{
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(
                     smem_u32(dst_smem)),
                 "r"(ncols)
                 : "memory");
}
__device__ __forceinline__ void tmem_relinquish()
{
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols)
{
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(taddr), "r"(ncols)
                 : "memory");
}
__device__ __forceinline__ void tc_fence_before()
{
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after()
{
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
}
// arrives on `bar` once all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void tc_commit(uint64_t *bar)
{
    asm volatile(
        "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(
            smem_u32(bar))
        : "memory");
}
// the same arrive on the mbarrier at this shared-memory offset in every CTA of cta_mask
__device__ __forceinline__ void tc_commit_multicast(uint64_t *bar, uint16_t cta_mask)
{
    asm volatile(
        "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n" ::"r"(
            smem_u32(bar)),
        "h"(cta_mask)
        : "memory");
}
__device__ __forceinline__ void tc_mma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                          uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tc_mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                           uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// 32 lanes x 32 consecutive 32-bit columns -> 32 registers per thread (thread = lane)
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
          "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
          "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]),
          "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]),
          "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&r)[16])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
          "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
          "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait()
{
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
}
// 16 registers per thread -> 32 lanes x 16 consecutive 32-bit columns (thread = lane): used to seed
// an accumulator with its per-column integer bias before the MMAs accumulate on top of it
__device__ __forceinline__ void tmem_st_32x16(uint32_t taddr, const uint32_t (&r)[16])
{
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};\n" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait()
{
    asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
}

// UMMA shared-memory matrix descriptor, K-major operand, SWIZZLE_128B, rows of 128 bytes:
// 8-row x 128-byte swizzle atoms stacked every 1024 bytes (SBO); LBO unused for swizzled K-major.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr)
{
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);  // start address, bits [0,14)
    d |= static_cast<uint64_t>(1) << 16;                     // LBO (ignored), bits [16,30)
    d |= static_cast<uint64_t>(1024 >> 4) << 32;             // SBO = 1024 B, bits [32,46)
    d |= static_cast<uint64_t>(1) << 46;                     // descriptor version (sm_100)
    d |= static_cast<uint64_t>(2) << 61;                     // layout: SWIZZLE_128B
    return d;
}

// the same for SWIZZLE_64B, rows of 64 bytes: 8-row x 64-byte atoms stacked every 512 bytes
__device__ __forceinline__ uint64_t umma_desc_sw64(uint32_t smem_addr)
{
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
    d |= static_cast<uint64_t>(1) << 16;
    d |= static_cast<uint64_t>(512 >> 4) << 32;              // SBO = 512 B
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(4) << 61;                     // layout: SWIZZLE_64B
    return d;
}

// UMMA instruction descriptor (32-bit), dense, K-major A and B.
//   c_fmt: 1 = F32, 2 = S32;  ab_fmt: kind::i8 -> 1 = S8;  kind::f16 -> 0 = F16, 1 = BF16
__host__ __device__ constexpr uint32_t umma_idesc(uint32_t c_fmt, uint32_t ab_fmt, uint32_t m,
                                                  uint32_t n)
{
    return (c_fmt << 4) | (ab_fmt << 7) | (ab_fmt << 10) | ((n >> 3) << 17) | ((m >> 4) << 24);
}

// ---- quantised epilogue (device copy of b200_epilogue) ---------------------------
struct EpiScalars {
    const float *mult;
    const float *badd;
    const int32_t *ibias;
    const int8_t *post_lut;
    int32_t zp_out, act, q6;
};
inline EpiScalars make_epi(const b200_epilogue &e)
{
    EpiScalars s;
    s.mult = e.mult;
    s.badd = e.badd;
    s.ibias = e.ibias;
    s.post_lut = e.post_lut;
    s.zp_out = e.zp_out;
    s.act = e.act;
    s.q6 = e.q6;
    return s;
}

__device__ __forceinline__ int clamp_i8(int v) { return max(-128, min(127, v)); }

// int8 epilogue for one accumulator (ibias already added); contract in b200nn.h.
// __float2int_rn is cvt.rni (round half even) and saturates, the host bounds |f| < 2^22.
__device__ __forceinline__ int requant_i8(int acc, float mult, float badd, int zp_out, int act,
                                          int q6)
{
    float f = fmaf(static_cast<float>(acc), mult, badd);
    int q = clamp_i8(__float2int_rn(f) + zp_out);
    if (act != B200_ACT_NONE) q = max(q, zp_out);
    if (act == B200_ACT_RELU6) q = min(q, q6);
    return q;
}

// ---- fast epilogue pieces ---------------------------------------------------------------
// Same contract as requant_i8, fewer instructions: the epilogues are issue-bound on the
// memory-bound layers, so every op per output counts.
//
// (q0,q1,q2,q3) -> 4 saturated int8 in one word: two cvt.pack.sat (I2IP) instead of eight
// min/max plus shifts and ors.  d = {c.lo16, sat(a), sat(b)} with b in byte 0.
__device__ __forceinline__ uint32_t pack4_sat_i8(int q0, int q1, int q2, int q3)
{
    uint32_t hi, d;
    asm("cvt.pack.sat.s8.s32.b32 %0, %1, %2, %3;" : "=r"(hi) : "r"(q3), "r"(q2), "r"(0));
    asm("cvt.pack.sat.s8.s32.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(q1), "r"(q0), "r"(hi));
    return d;
}

// four requantised values through the post table (shared memory, indexed q + 128) into one word
__device__ __forceinline__ uint32_t lut4_i8(int q0, int q1, int q2, int q3, const uint8_t *lut)
{
    const uint32_t b0 = lut[clamp_i8(q0) + 128], b1 = lut[clamp_i8(q1) + 128];
    const uint32_t b2 = lut[clamp_i8(q2) + 128], b3 = lut[clamp_i8(q3) + 128];
    return __byte_perm(__byte_perm(b0, b1, 0x0040), __byte_perm(b2, b3, 0x0040), 0x5410);
}

// Exact int <-> float through the 1.5 * 2^23 magic constant, valid while |value| < 2^22 (the
// host guarantees it: b200_opt/quant.c bounds |acc * mult + bias| and the depthwise / small-K
// accumulators): an accumulator that was INITIALISED with kMagicI + ibias is turned into
// float(acc) by one FADD, and round-half-even back to int is one FADD + one IADD.
constexpr int kMagicI = 0x4B400000;
constexpr float kMagicF = 12582912.0f;
__device__ __forceinline__ float magic_to_float(int biased_acc)
{
    return __fsub_rn(__int_as_float(biased_acc), kMagicF);
}
__device__ __forceinline__ int magic_round(float f, int zp_minus_magic)
{
    return __float_as_int(__fadd_rn(f, kMagicF)) + zp_minus_magic;
}

// Packed f32x2 arithmetic (FADD2 / FFMA2 on sm_100): one issue slot for two lanes' worth of
// epilogue math.  IEEE round-to-nearest per element, so results equal the scalar sequence.
__device__ __forceinline__ uint64_t f2_pack(float lo, float hi)
{
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ uint64_t f2_pack_bits(uint32_t lo, uint32_t hi)
{
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi));
    return r;
}
__device__ __forceinline__ void f2_unpack_bits(uint64_t v, int &lo, int &hi)
{
    asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t f2_add(uint64_t a, uint64_t b)
{
    uint64_t r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ uint64_t f2_fma(uint64_t a, uint64_t b, uint64_t c)
{
    uint64_t r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
// two accumulators -> kMagicI + round_half_even(fma(float(acc), mult, badd)) each, as int bits.
// MAGIC: the accumulators carry + kMagicI already (they were seeded with it), so float(acc) is
// one packed FADD; otherwise two I2F.
template <bool MAGIC>
__device__ __forceinline__ void requant_pair(uint32_t a0, uint32_t a1, uint64_t mult2, uint64_t badd2, int &t0,
                                             int &t1)
{
    uint64_t x;
    if (MAGIC)
        x = f2_add(f2_pack_bits(a0, a1), f2_pack(-kMagicF, -kMagicF));
    else
        x = f2_pack(static_cast<float>(static_cast<int>(a0)), static_cast<float>(static_cast<int>(a1)));
    x = f2_fma(x, mult2, badd2);
    x = f2_add(x, f2_pack(kMagicF, kMagicF));
    f2_unpack_bits(x, t0, t1);
}

// epilogue specialisations
enum { EPI_PLAIN = 0, EPI_RELU = 1, EPI_RELU6 = 2, EPI_LUT = 3, EPI_GENERIC = 4 };

// four int8 outputs from four rounded values t = kMagicI + round_half_even(f) -> one packed word.
// EPI_LUT: lut_lo arrives minus the table's shared-memory address (lut_base), so the clamp of the
// table index and the address addition are the same two instructions.
template <int MODE>
__device__ __forceinline__ uint32_t finish4(const int (&t)[4], const EpiScalars &ep, const uint8_t *lut,
                                            bool has_lut, int zp_m, int lut_lo, int lut_base)
{
    int q[4];
#pragma unroll
    for (int e = 0; e < 4; e++) {
        if (MODE == EPI_LUT) {
            q[e] = min(max(t[e] - lut_lo, lut_base), lut_base + 255);  // &lut[clamp(q, -128, 127) + 128]
        } else {
            q[e] = t[e] + zp_m;
            if (MODE == EPI_RELU || MODE == EPI_RELU6) q[e] = max(q[e], ep.zp_out);
            if (MODE == EPI_RELU6) q[e] = min(q[e], ep.q6);
            if (MODE == EPI_GENERIC) {
                if (ep.act != B200_ACT_NONE) q[e] = max(q[e], ep.zp_out);
                if (ep.act == B200_ACT_RELU6) q[e] = min(q[e], ep.q6);
            }
        }
    }
    if (MODE == EPI_LUT) {
        uint32_t b0, b1, b2, b3;
        asm("ld.shared.u8 %0, [%1];" : "=r"(b0) : "r"(q[0]));
        asm("ld.shared.u8 %0, [%1];" : "=r"(b1) : "r"(q[1]));
        asm("ld.shared.u8 %0, [%1];" : "=r"(b2) : "r"(q[2]));
        asm("ld.shared.u8 %0, [%1];" : "=r"(b3) : "r"(q[3]));
        return __byte_perm(__byte_perm(b0, b1, 0x0040), __byte_perm(b2, b3, 0x0040), 0x5410);
    }
    if (MODE == EPI_GENERIC && has_lut) return lut4_i8(q[0], q[1], q[2], q[3], lut);
    return pack4_sat_i8(q[0], q[1], q[2], q[3]);
}

__device__ __forceinline__ uint32_t pack4_i8(int a, int b, int c, int d)
{
    return (static_cast<uint32_t>(a) & 0xFF) | ((static_cast<uint32_t>(b) & 0xFF) << 8) |
           ((static_cast<uint32_t>(c) & 0xFF) << 16) | ((static_cast<uint32_t>(d) & 0xFF) << 24);
}

__device__ __forceinline__ float act_f(float v, int act)
{
    if (act != B200_ACT_NONE) v = v > 0.f ? v : 0.f;
    if (act == B200_ACT_RELU6) v = fminf(v, 6.f);
    return v;
}

// exact restatement of source/nn2/utils.c:550 float_to_int8_base on the device:
// IEEE division, round half even, clamp.  Used by add / pool / softmax.
__device__ __forceinline__ int quant_i8_exact(float x, float s, int zp)
{
    float t = __fadd_rn(rintf(__fdiv_rn(x, s)), static_cast<float>(zp));
    t = fminf(fmaxf(t, -128.f), 127.f);
    return static_cast<int>(t);
}
__device__ __forceinline__ float dequant_i8(int q, float s, int zp)
{
    return __fmul_rn(__fsub_rn(static_cast<float>(q), static_cast<float>(zp)), s);
}

}  // namespace b200
