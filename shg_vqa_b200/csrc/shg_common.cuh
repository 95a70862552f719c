// Common device/host helpers for libshgvqa (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/shgvqa.h"

#if defined(__CUDA_ARCH__) && !defined(__CUDA_ARCH_FEAT_SM100_ALL)
#error "libshgvqa is written for sm_100a only: compile with -gencode arch=compute_100a,code=sm_100a"
#endif

namespace shg {

void set_error(const char* fmt, ...);
int check_launch(const char* what);

#define SHG_REQUIRE(cond, code, ...)      \
    do {                                  \
        if (!(cond)) {                    \
            shg::set_error(__VA_ARGS__);  \
            return (code);                \
        }                                 \
    } while (0)

constexpr int kNumSMs = 148;

// Every kernel of the library is launched through shg::launch (cudaLaunchKernelEx on the caller's stream).
template <class... KArgs, class... Args>
inline cudaError_t launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                          Args&&... args) {
    static_assert(sizeof...(KArgs) == sizeof...(Args), "argument count mismatch");
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid, cfg.blockDim = block, cfg.dynamicSmemBytes = smem, cfg.stream = stream;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}"
        : "=r"(pred));
    return pred != 0;
}

__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Arrive without release semantics: used where the only thing handed over is TMEM that this thread has already
// finished reading (tcgen05.wait::ld + tcgen05.fence::before_thread_sync) — a releasing arrive would also wait for the
// thread's outstanding global stores (MEMBAR), which stalls the epilogue for a full HBM round trip per tile.
__device__ __forceinline__ void mbar_arrive_relaxed(uint64_t* bar) {
    asm volatile("mbarrier.arrive.relaxed.cta.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred P;\n"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P, [%0], %1;\n\t"
        "@P bra DONE_%=;\n\t"
        "bra WAIT_%=;\n"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ---------------------------------------------------------------- TMA
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::
            "r"(smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}

// TMA store of a [box] tile staged in shared memory (128B-swizzled like the loads); clips at the tensor edge.
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, const void* src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                     reinterpret_cast<uint64_t>(m)),
                 "r"(smem_u32(src)), "r"(c0), "r"(c1)
                 : "memory");
}
// TMA reduce-add of a staged [box] fp32 tile into global memory: the adds happen in L2 on whole lines; clips at the edge.
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* m, const void* src, int c0, int c1) {
    asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                     reinterpret_cast<uint64_t>(m)),
                 "r"(smem_u32(src)), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// all but the newest N bulk groups of this thread have finished READING their shared-memory source
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}

// ---------------------------------------------------------------- tcgen05 / TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
                 "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], bf16 in / fp32 accumulate, one CTA.
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// registers -> 16 consecutive 32-bit columns of this thread's TMEM lane (the mirror of tmem_ld16)
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
        "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
        "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
        : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(v[0]),
                 "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// D[tmem] (+)= A[tmem] * B[smem]: the A operand (M = 128 rows in the lanes, K-major, two bf16 per 32-bit column, so one
// 16-deep k-step is 8 columns) is read from tensor memory — written there by tmem_st16 — instead of a swizzled
// shared-memory tile.  A from TMEM cannot be transposed (idesc a_major = K).
__device__ __forceinline__ void umma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
        "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

// Instruction descriptor, kind::f16, BF16 x BF16 -> F32 (cute/arch/mma_sm100_desc.hpp bit layout):
// c_format[4,6)=1 (F32) a_format[7,10)=1 (BF16) b_format[10,13)=1, a_major bit15, b_major bit16 (1 = MN-major),
// n_dim[17,23)=N>>3, m_dim[24,29)=M>>4.
__host__ __device__ constexpr uint32_t make_idesc_bf16(int m, int n, int a_mn_major, int b_mn_major) {
    return (1u << 4) | (1u << 7) | (1u << 10) | (uint32_t(a_mn_major) << 15) | (uint32_t(b_mn_major) << 16) |
           (uint32_t(n >> 3) << 17) | (uint32_t(m >> 4) << 24);
}

// Shared-memory matrix descriptor (SWIZZLE_128B, version 1).  Offsets in bytes.
//   K-major  : rows of 128 B (64 bf16 along K); 8-row groups SBO apart (1024 B when dense); LBO unused.
//   MN-major : rows of 128 B (64 bf16 along MN); 8 k-rows per group, groups SBO (=1024 B) apart along K,
//              64-wide MN atoms LBO apart.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= uint64_t((smem_addr & 0x3FFFF) >> 4);
    d |= uint64_t((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= uint64_t((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= uint64_t(1) << 46;  // version = 1 (Blackwell)
    d |= uint64_t(2) << 61;  // SWIZZLE_128B
    return d;
}

// ---------------------------------------------------------------- CTA pairs (cta_group::2)
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// address of the same shared-memory object in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t smem_addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
                 "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// TMA load issued by either CTA of a pair; the transaction bytes land on the LEADER CTA's mbarrier (peer bit cleared)
__device__ __forceinline__ void tma_load_2d_pair(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], "
        "[%2];" ::"r"(smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(c0), "r"(c1)
        : "memory");
}
// the same with a multicast mask: the box lands at the same offset in every CTA of `mask` and completes bytes on the
// barrier of each destination's pair leader
__device__ __forceinline__ void tma_load_2d_pair_mc(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                                                    uint16_t mask) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster "
        "[%0], [%1, {%3, %4}], [%2], %5;" ::"r"(smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(c0), "r"(c1), "h"(mask)
        : "memory");
}
// commit of a pair's MMAs arriving on the barrier at this offset in every CTA of `mask` (cluster ranks)
__device__ __forceinline__ void tc_commit_mask(uint64_t* bar, uint16_t mask) {
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
            smem_u32(bar)),
        "h"(mask)
        : "memory");
}
// commit of the pair's MMAs, arriving on the barrier at this offset in BOTH CTAs
__device__ __forceinline__ void tc_commit_pair(uint64_t* bar) {
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
            smem_u32(bar)),
        "h"((uint16_t)3)
        : "memory");
}
// D[tmem of both CTAs] (+)= A * B across the CTA pair: M = 256 (128 rows from each CTA's smem), B's N split in halves
__device__ __forceinline__ void umma_bf16_pair(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                               uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

// ---------------------------------------------------------------- misc math / packing
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t u) {
    __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&u);
    return __bfloat1622float2(v);
}
// 2^x and 1/x on the SFU, flush-to-zero: one MUFU each.
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// Exact-erf GELU (lxrt/modeling_capsbert.py:127-133), x * 0.5 * (1 + erf(x / sqrt 2)), with erf from Abramowitz &
// Stegun 7.1.26 (|error| <= 1.5e-7 — four orders below the bf16 resolution of the stored result) on two SFU ops and
// seven FMAs instead of erff's ~25 instructions: on a K = 768 projection the activation otherwise costs the SM as many
// issue slots as the tensor cores need for the whole tile.  `e` returns exp(-x^2/2) for the derivative.
__device__ __forceinline__ float gelu_phi(float x, float& e) {
    const float z = fabsf(x) * 0.70710678118654752440f;
    const float k = rcp_approx(fmaf(0.3275911f, z, 1.0f));
    float poly = fmaf(1.061405429f, k, -1.453152027f);
    poly = fmaf(poly, k, 1.421413741f);
    poly = fmaf(poly, k, -0.284496736f);
    poly = fmaf(poly, k, 0.254829592f);
    poly *= k;
    e = ex2_approx(z * z * -1.4426950408889634f);
    const float erf_abs = fmaf(-poly, e, 1.0f);
    return fmaf(copysignf(erf_abs, x), 0.5f, 0.5f);          // Phi(x)
}
__device__ __forceinline__ float gelu_erf(float x) {
    float e;
    return x * gelu_phi(x, e);
}
__device__ __forceinline__ float gelu_erf_grad(float x) {
    // d/dx [x Phi(x)] = Phi(x) + x phi(x),  phi(x) = exp(-x^2/2) / sqrt(2 pi)
    float e;
    const float phi_cdf = gelu_phi(x, e);
    return fmaf(x * 0.39894228040143267794f, e, phi_cdf);
}

// Batched forms for the GEMM epilogues.  Written stage by stage over 8 elements, with the two SFU operations as
// `asm volatile` so that ptxas keeps the stages apart: left to itself it serialises the per-element chains through
// three registers (MUFU latency exposed 2x per element, measured +37 us on the 12576x3072x768 FFN projection).
__device__ __forceinline__ float ex2_approx_v(float x) {
    float y;
    asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rcp_approx_v(float x) {
    float y;
    asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// phi[i] = Phi(x[i]); e[i] = exp(-x[i]^2 / 2)
__device__ __forceinline__ void gelu_phi_x8(const float* x, float* phi, float* e) {
    float z[8], k[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) z[i] = fabsf(x[i]) * 0.70710678118654752440f;
#pragma unroll
    for (int i = 0; i < 8; ++i) k[i] = rcp_approx_v(fmaf(0.3275911f, z[i], 1.0f));
#pragma unroll
    for (int i = 0; i < 8; ++i) e[i] = ex2_approx_v(z[i] * z[i] * -1.4426950408889634f);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        float poly = fmaf(1.061405429f, k[i], -1.453152027f);
        poly = fmaf(poly, k[i], 1.421413741f);
        poly = fmaf(poly, k[i], -0.284496736f);
        poly = fmaf(poly, k[i], 0.254829592f);
        poly *= k[i];
        const float erf_abs = fmaf(-poly, e[i], 1.0f);
        phi[i] = fmaf(copysignf(erf_abs, x[i]), 0.5f, 0.5f);
    }
}
template <int N>
__device__ __forceinline__ void gelu_erf_batch(float (&v)[N]) {
#pragma unroll
    for (int b = 0; b < N; b += 8) {
        float phi[8], e[8];
        gelu_phi_x8(v + b, phi, e);
#pragma unroll
        for (int i = 0; i < 8; ++i) v[b + i] *= phi[i];
    }
}
// acc[i] *= gelu'(x[i])
template <int N>
__device__ __forceinline__ void gelu_erf_grad_batch(float (&acc)[N], const float (&x)[N]) {
#pragma unroll
    for (int b = 0; b < N; b += 8) {
        float phi[8], e[8];
        gelu_phi_x8(x + b, phi, e);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[b + i] *= fmaf(x[b + i] * 0.39894228040143267794f, e[i], phi[i]);
    }
}

// ---------------------------------------------------------------- dropout (training mode)
// Counter-based Bernoulli mask: element `idx` of call site `salt` at step seed `seed` is kept iff a 16-bit slice of
// a 32-bit integer hash of (seed, salt, idx >> 1) is >= p * 65536.  Forward and backward kernels regenerate the
// same mask from the same three numbers — nothing is stored.  (The reference draws from torch's Philox stream,
// nn.Dropout at lxrt/modeling_capsbert.py:311,339,381,429,483 and lxrt/transformer.py:196-204; the streams cannot
// match, the distribution does: iid keep probability 1-p, survivors scaled by 1/(1-p).)
struct DropCfg {
    float p;                 // 0 = off
    uint32_t salt;
    const uint32_t* seed;    // device scalar, advanced every step outside the CUDA graph
};
__device__ __forceinline__ uint32_t mix32(uint32_t x) {
    x ^= x >> 16;
    x *= 0x7feb352du;
    x ^= x >> 15;
    x *= 0x846ca68bu;
    x ^= x >> 16;
    return x;
}
__device__ __forceinline__ uint32_t drop_key(uint32_t seed, uint32_t salt) { return mix32(seed ^ (salt * 0x9E3779B9u)); }
__device__ __forceinline__ uint32_t drop_thresh(float p) { return (uint32_t)(p * 65536.0f + 0.5f); }
// keep-bits of the element pair (2*pair, 2*pair+1): bit 0 / bit 1
__device__ __forceinline__ uint32_t drop_pair(uint32_t key, uint32_t pair, uint32_t thresh) {
    const uint32_t h = mix32(key + pair * 0x85ebca6bu);
    return ((h & 0xffffu) >= thresh ? 1u : 0u) | ((h >> 16) >= thresh ? 2u : 0u);
}
// AND-mask for a packed bf16x2 of that pair (dropped lanes zeroed)
__device__ __forceinline__ uint32_t drop_pair_mask(uint32_t key, uint32_t pair, uint32_t thresh) {
    const uint32_t h = mix32(key + pair * 0x85ebca6bu);
    return ((h & 0xffffu) >= thresh ? 0x0000ffffu : 0u) | ((h >> 16) >= thresh ? 0xffff0000u : 0u);
}
__device__ __forceinline__ bool drop_keep(uint32_t key, uint64_t idx, uint32_t thresh) {
    return (drop_pair(key, (uint32_t)(idx >> 1), thresh) >> (idx & 1)) & 1u;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// ---------------------------------------------------------------- host: TMA descriptor cache
// 2-D .. 5-D bf16 (or fp32) tensor maps with 128-byte swizzle (out-of-bounds elements read as zero), cached by (ptr, dims, strides, box).
constexpr int kTmapBf16 = 0, kTmapF32 = 1;
const CUtensorMap* get_tmap(const void* ptr, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                            const uint32_t* box, int dtype);
inline const CUtensorMap* get_tmap_bf16(const void* ptr, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                                        const uint32_t* box) {
    return get_tmap(ptr, rank, dims, strides_bytes, box, kTmapBf16);
}

}  // namespace shg
