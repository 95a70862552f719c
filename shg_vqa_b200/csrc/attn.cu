// Batched multi-head attention (head dim 64) on tcgen05: forward.
//
// One CTA per (128-query tile, head, batch element).  Every key/value row of the head fits in shared memory
// (Lk <= 400 on this path: 40 text, 393 visual, 128/48 queries, 177 hypergraph tokens), so there is no online
// softmax: S = Q.K^T is produced whole in TMEM (<= 416 fp32 columns), a pair of the 256 softmax threads owns one
// query row (tcgen05.ld 32x32b — no shuffles), applies scale + mask, exponentiates (exp2), and writes P as bf16 into
// 128B-swizzled shared memory (aliasing the dead Q/K tiles) as the A operand of O = P.V, with V read as an
// MN-major B operand exactly where TMA dropped it.  Masks: none | additive key mask (BERT's -10000 padding,
// lxrt/modeling_capsbert.py:1834) | block-causal by frame (-inf for later frames, lxrt/entry.py:114-121).
#include <stdlib.h>

#include "shg_common.cuh"

namespace shg {

constexpr int kHeadDim = 64;
constexpr int kAttnThreads = 512;   // 16 warps: four threads per query row, each a quarter of the key chunks

struct AttnParams {
    __nv_bfloat16* o;
    int64_t ldo;
    float* o_f32;           // optional fp32 copy of O (consumed by the backward's delta)
    int64_t ldof;
    float* lse;
    const float* key_mask;  // [batch, lk] additive, or null
    int batch, heads, lq, lk;
    int lk16, lk32;         // lk rounded up to 16 / 32
    int n_kbox;             // 64-row K/V boxes
    int v_off;              // byte offset of V in smem
    int tmem_cols, o_col;
    float scale;
    int mask_mode, q_per_frame, k_per_frame;
    DropCfg drop;           // dropout on the probabilities (training mode)
    uint32_t* drop_bits;    // optional out: keep-bit words [batch, heads, lq, lk32/32]
};

// Dropout on the probabilities: element (b, h, q, k) of a call site is pair (row_pair0(b,h,q) + k/2), lane k & 1 of
// the counter-based mask in shg_common.cuh; the forward and both backward passes index it the same way.
__device__ __forceinline__ uint32_t attn_row_pair0(int b, int heads, int h, int lq, int q, int lk) {
    return (uint32_t)(((size_t)b * heads + h) * lq + q) * (uint32_t)(((lk + 31) & ~31) >> 1);
}

// 512 threads: thread (row, quad) owns query row `row` of the tile and one quarter of the key chunks (16 warps keep
// four in flight per scheduler: the softmax is latency-bound — tcgen05.ld, ex2, pack, st.shared chains).
template <bool HAS_KMASK, bool DROP>
__global__ void __launch_bounds__(kAttnThreads)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tma_q, const __grid_constant__ CUtensorMap tma_k,
                const __grid_constant__ CUtensorMap tma_v, const AttnParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sQ = smem;                 // 128 x 128 B
    uint8_t* sK = smem + 16384;         // n_kbox x (64 x 128 B)
    uint8_t* sP = smem;                 // aliases Q/K once S is complete: blocks of 128 x 128 B
    uint8_t* sV = smem + p.v_off;
    float* s_kmask = reinterpret_cast<float*>(sV + p.n_kbox * 8192);   // [448] additive key mask * log2(e)
    float* s_red = s_kmask + 448;                                       // [4][128] row max, then row sum
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_red + 512);
    uint64_t* bar_qk = bars;
    uint64_t* bar_v = bars + 1;
    uint64_t* bar_s = bars + 2;
    uint64_t* bar_o = bars + 3;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int q0 = blockIdx.x * 128, h = blockIdx.y, b = blockIdx.z;

    if (tid == 0) {
        mbar_init(bar_qk, 1);
        mbar_init(bar_v, 1);
        mbar_init(bar_s, 1);
        mbar_init(bar_o, 1);
        fence_barrier_init();
        tma_prefetch_desc(&tma_q);
        tma_prefetch_desc(&tma_k);
        tma_prefetch_desc(&tma_v);
    }
    if (warp == 0) tmem_alloc(tmem_slot, p.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (tid == 0) {
        mbar_expect_tx(bar_qk, 16384 + p.n_kbox * 8192);
        tma_load_3d(sQ, &tma_q, bar_qk, h * kHeadDim, q0, b);
        for (int i = 0; i < p.n_kbox; ++i) tma_load_3d(sK + i * 8192, &tma_k, bar_qk, h * kHeadDim, i * 64, b);
        mbar_expect_tx(bar_v, p.n_kbox * 8192);
        for (int i = 0; i < p.n_kbox; ++i) tma_load_3d(sV + i * 8192, &tma_v, bar_v, h * kHeadDim, i * 64, b);
        // S = Q . K^T   (M = 128, N = lk16 in chunks of <= 256, K = 64)
        mbar_wait(bar_qk, 0);
        tc_fence_after();
        const uint32_t aq = smem_u32(sQ), ak = smem_u32(sK);
        for (int n_off = 0; n_off < p.lk16; n_off += 256) {
            const int n = min(256, p.lk16 - n_off);
            const uint32_t idesc = make_idesc_bf16(128, n, 0, 0);
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
                umma_bf16(tmem + n_off, make_smem_desc(aq + ks * 32, 16, 1024),
                          make_smem_desc(ak + n_off * 128 + ks * 32, 16, 1024), idesc, ks > 0);
        }
        tc_commit(bar_s);
    }
    __syncwarp();
    if (HAS_KMASK) {
        for (int k = tid; k < p.lk32; k += kAttnThreads)
            s_kmask[k] = k < p.lk ? p.key_mask[(size_t)b * p.lk + k] * 1.4426950408889634f : 0.f;
        __syncthreads();
    }

    // ---------------------------------------------------------------- softmax in the log2 domain
    const int row = (warp & 3) * 32 + lane, quad = warp >> 2;   // warp w may touch TMEM lanes [32(w%4), +32)
    const int q = q0 + row;
    const uint32_t t_row = tmem + (uint32_t((warp & 3) * 32) << 16);
    // keys < k_lim are visible to this row (block-causal: frames <= the query's frame)
    const int k_lim = p.mask_mode == 2 ? min(p.lk, (q / p.q_per_frame + 1) * p.k_per_frame) : p.lk;
    const float c1 = p.scale * 1.4426950408889634f;
    const int n_ch = p.lk32 >> 5;
    // warps whose 32 query rows lie beyond lq (last tile of 393 = 3 x 128 + 9) skip the softmax: their P rows only
    // feed O rows that are never stored
    const bool rows_live = q0 + (warp & 3) * 32 < p.lq;
    const int ch0 = (n_ch * quad) >> 2, ch1 = rows_live ? (n_ch * (quad + 1)) >> 2 : ch0;

    mbar_wait(bar_s, 0);
    tc_fence_after();
    float mx = -INFINITY;
    for (int ch = ch0; ch < ch1; ++ch) {
        const int c = ch << 5;
        uint32_t raw[32];
        tmem_ld32(t_row + c, raw);
        tmem_ld_wait();
        if (HAS_KMASK) {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const float t = fmaf(__uint_as_float(raw[i]), c1, s_kmask[c + i]);
                mx = fmaxf(mx, c + i < k_lim ? t : -INFINITY);
            }
        } else if (c + 32 <= k_lim) {
#pragma unroll
            for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(raw[i]));
        } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) mx = fmaxf(mx, c + i < k_lim ? __uint_as_float(raw[i]) : -INFINITY);
        }
    }
    if (!HAS_KMASK) mx *= c1;    // scale > 0: the max commutes with the scaling
    s_red[quad * 128 + row] = mx;
    __syncthreads();
    mx = fmaxf(fmaxf(s_red[row], s_red[128 + row]), fmaxf(s_red[256 + row], s_red[384 + row]));
    if (mx == -INFINITY) mx = 0.f;
    __syncthreads();             // s_red is reused for the sums below
    float sum = 0.f;
    uint32_t dkey = 0, dth = 0, row_pair0 = 0;
    if (DROP) {
        dkey = drop_key(__ldg(p.drop.seed), p.drop.salt), dth = drop_thresh(p.drop.p);
        row_pair0 = attn_row_pair0(b, p.heads, h, p.lq, q, p.lk);
    }
    for (int ch = ch0; ch < ch1; ++ch) {
        const int c = ch << 5;
        uint32_t raw[32];
        tmem_ld32(t_row + c, raw);
        tmem_ld_wait();
        float pv[32];
        if (c + 32 <= k_lim) {          // whole chunk visible: no per-element predicate
#pragma unroll
            for (int i = 0; i < 32; ++i)
                pv[i] = ex2_approx(HAS_KMASK ? fmaf(__uint_as_float(raw[i]), c1, s_kmask[c + i] - mx)
                                             : fmaf(__uint_as_float(raw[i]), c1, -mx));
        } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const float t = HAS_KMASK ? fmaf(__uint_as_float(raw[i]), c1, s_kmask[c + i] - mx)
                                          : fmaf(__uint_as_float(raw[i]), c1, -mx);
                pv[i] = c + i < k_lim ? ex2_approx(t) : 0.f;
            }
        }
        uint8_t* blk = sP + (c >> 6) * 16384 + row * 128;
        const int j0 = (c & 32) ? 4 : 0;
        uint32_t keep_bits = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            uint4 u;
            u.x = pack_bf16x2(pv[8 * j], pv[8 * j + 1]), u.y = pack_bf16x2(pv[8 * j + 2], pv[8 * j + 3]);
            u.z = pack_bf16x2(pv[8 * j + 4], pv[8 * j + 5]), u.w = pack_bf16x2(pv[8 * j + 6], pv[8 * j + 7]);
            // The row is normalised by the sum of the ROUNDED probabilities the tensor core consumes, so that the
            // stored O is an exact convex combination of V rows (sum_j P_ij = 1): the backward's
            // delta = rowsum(dO o O_f32) then cancels against dP to second order (see attn_bwd_dq_kernel).
            sum += (__uint_as_float(u.x << 16) + __uint_as_float(u.x & 0xffff0000u)) +
                   (__uint_as_float(u.y << 16) + __uint_as_float(u.y & 0xffff0000u)) +
                   (__uint_as_float(u.z << 16) + __uint_as_float(u.z & 0xffff0000u)) +
                   (__uint_as_float(u.w << 16) + __uint_as_float(u.w & 0xffff0000u));
            if (DROP) {
                // softmax is normalised over ALL keys; the dropped ones are zeroed only in the P.V operand and
                // the survivors' 1/(1-p) is folded into the final row scaling
                const uint32_t pb = row_pair0 + ((c + 8 * j) >> 1);
                const uint32_t k0 = drop_pair(dkey, pb, dth), k1 = drop_pair(dkey, pb + 1, dth);
                const uint32_t k2 = drop_pair(dkey, pb + 2, dth), k3 = drop_pair(dkey, pb + 3, dth);
                u.x &= (k0 & 1u) * 0xffffu + (k0 >> 1) * 0xffff0000u, u.y &= (k1 & 1u) * 0xffffu + (k1 >> 1) * 0xffff0000u;
                u.z &= (k2 & 1u) * 0xffffu + (k2 >> 1) * 0xffff0000u, u.w &= (k3 & 1u) * 0xffffu + (k3 >> 1) * 0xffff0000u;
                keep_bits |= (k0 | (k1 << 2) | (k2 << 4) | (k3 << 6)) << (8 * j);
            }
            *reinterpret_cast<uint4*>(blk + (((j0 + j) ^ (row & 7)) << 4)) = u;
        }
        if (DROP && p.drop_bits && q < p.lq)
            p.drop_bits[(((size_t)b * p.heads + h) * p.lq + q) * (p.lk32 >> 5) + (c >> 5)] = keep_bits;
    }
    s_red[quad * 128 + row] = sum;
    fence_proxy_async_smem();   // generic-proxy writes of P -> visible to the tensor core's async proxy
    tc_fence_before();
    __syncthreads();
    sum = (s_red[row] + s_red[128 + row]) + (s_red[256 + row] + s_red[384 + row]);

    if (tid == 0) {
        tc_fence_after();
        mbar_wait(bar_v, 0);
        // O = P . V   (M = 128, N = 64, K = lk16): A = P (K-major blocks), B = V (MN-major, 8-row groups 1024 B apart)
        const uint32_t ap = smem_u32(sP), av = smem_u32(sV);
        const uint32_t idesc = make_idesc_bf16(128, kHeadDim, 0, 1);
        for (int k = 0; k < p.lk16; k += 16)
            umma_bf16(tmem + p.o_col, make_smem_desc(ap + (k >> 6) * 16384 + ((k & 63) >> 4) * 32, 16, 1024),
                      make_smem_desc(av + k * 128, 8192, 1024), idesc, k > 0);
        tc_commit(bar_o);
    }
    __syncwarp();
    mbar_wait(bar_o, 0);
    tc_fence_after();
    const float inv = DROP ? 1.0f / (sum * (1.0f - p.drop.p)) : 1.0f / sum;
    if (rows_live) {
        uint32_t raw[16];
        tmem_ld16(t_row + p.o_col + quad * 16, raw);
        tmem_ld_wait();
        if (q < p.lq) {
            __nv_bfloat16* dst = p.o + ((size_t)b * p.lq + q) * p.ldo + h * kHeadDim + quad * 16;
#pragma unroll
            for (int i = 0; i < 16; i += 8) {
                uint4 u;
                u.x = pack_bf16x2(__uint_as_float(raw[i]) * inv, __uint_as_float(raw[i + 1]) * inv);
                u.y = pack_bf16x2(__uint_as_float(raw[i + 2]) * inv, __uint_as_float(raw[i + 3]) * inv);
                u.z = pack_bf16x2(__uint_as_float(raw[i + 4]) * inv, __uint_as_float(raw[i + 5]) * inv);
                u.w = pack_bf16x2(__uint_as_float(raw[i + 6]) * inv, __uint_as_float(raw[i + 7]) * inv);
                *reinterpret_cast<uint4*>(dst + i) = u;
            }
            if (p.o_f32) {
                float* df = p.o_f32 + ((size_t)b * p.lq + q) * p.ldof + h * kHeadDim + quad * 16;
#pragma unroll
                for (int i = 0; i < 16; i += 4)
                    *reinterpret_cast<float4*>(df + i) =
                        make_float4(__uint_as_float(raw[i]) * inv, __uint_as_float(raw[i + 1]) * inv,
                                    __uint_as_float(raw[i + 2]) * inv, __uint_as_float(raw[i + 3]) * inv);
            }
        }
    }
    // natural-log log-sum-exp of the scaled, masked scores
    if (p.lse && quad == 0 && q < p.lq)
        p.lse[((size_t)b * p.heads + h) * p.lq + q] = (mx + log2f(sum)) * 0.6931471805599453f;
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem, p.tmem_cols);
    }
}

// =====================================================================================================
// Forward for long key sequences (Lk > 256: the 393 visual tokens): chunked, two CTAs per SM.
//
// The kernel above keeps the whole score row in TMEM (416 + 64 columns -> a 512-column allocation) and Q, K, V of the
// head in shared memory (130 KB): ONE CTA per SM, whose phases — operand load, QK^T, max pass, exponential pass, PV,
// store — run strictly one after the other (ncu r01c: tensor pipe 6 % active, issue slots 33 % used: latency-bound).
// Here keys are processed in chunks of 128 with an online softmax, so a CTA needs 192 TMEM columns (S chunk + O
// chunk -> a 256-column allocation) and 85 KB of shared memory (Q, one K chunk, one V chunk, one P chunk): TWO CTAs are
// resident per SM and fill each other's bubbles.  Inside a CTA a ninth warp issues every TMA load and tcgen05.mma, and
// issues S_{c+1} = Q.K_{c+1}^T BEFORE O_c = P_c.V_c, so the softmax of chunk c+1 also overlaps the PV product of
// chunk c.  O is accumulated in REGISTERS (each PV product is a fresh [128 x 64] tile read back with tcgen05.ld):
//   a_c = a_{c-1} * 2^(M_{c-1} - M_c) + P_c.V_c,   l_c = l_{c-1} * 2^(M_{c-1} - M_c) + rowsum(P_c),   out = a / l
// with M_c the running row maximum (log2 domain), P_c = bf16(2^(s - M_c)) and rowsum taken over the ROUNDED values
// the tensor core consumes, so the output stays an exact convex combination of V rows (see the kernel above).
//   warps 0..7 : softmax — thread = (query row, half of the chunk's 128 key columns; half of O's 64 columns)
//   warp  8    : TMA + MMA issuer (one lane)
// =====================================================================================================
constexpr int kFwd2Threads = 288;
constexpr int kFwd2Chunk = 128;

struct Fwd2Bars {
    uint64_t q, k, v, s, sfree, p, pv;
};

template <bool HAS_KMASK, bool DROP>
__global__ void __launch_bounds__(kFwd2Threads, 2)
attn_fwd2_kernel(const __grid_constant__ CUtensorMap tma_q, const __grid_constant__ CUtensorMap tma_k,
                 const __grid_constant__ CUtensorMap tma_v, const AttnParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sQ = smem;                 // 128 x 128 B
    uint8_t* sK = smem + 16384;         // one chunk: 128 keys x 128 B
    uint8_t* sV = smem + 32768;         // one chunk
    float* s_kmask = reinterpret_cast<float*>(smem + 49152);          // [448] additive key mask * log2(e)
    float* s_red = s_kmask + 448;                                      // [2][128]
    Fwd2Bars* bars = reinterpret_cast<Fwd2Bars*>(s_red + 256);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int q0 = blockIdx.x * 128, h = blockIdx.y, b = blockIdx.z;
    const int n_ch = (p.lk + kFwd2Chunk - 1) / kFwd2Chunk;

    if (tid == 0) {
        mbar_init(&bars->q, 1);
        mbar_init(&bars->k, 1);
        mbar_init(&bars->v, 1);
        mbar_init(&bars->s, 1);
        mbar_init(&bars->sfree, 8);
        mbar_init(&bars->p, 8);
        mbar_init(&bars->pv, 1);
        fence_barrier_init();
        tma_prefetch_desc(&tma_q);
        tma_prefetch_desc(&tma_k);
        tma_prefetch_desc(&tma_v);
    }
    if (warp == 8) tmem_alloc(tmem_slot, 256);
    if (HAS_KMASK)
        for (int k = tid; k < p.lk32; k += kFwd2Threads)
            s_kmask[k] = k < p.lk ? p.key_mask[(size_t)b * p.lk + k] * 1.4426950408889634f : 0.f;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    // S chunk 128 columns | O tile 64 | P chunk (bf16 pairs) 64: P goes registers -> tcgen05.st -> TMEM and is the TMEM
    // A operand of the PV product — no swizzled shared-memory tile, no proxy fence, no A-operand shared-memory reads
    const uint32_t t_s = tmem, t_o = tmem + kFwd2Chunk, t_p = tmem + kFwd2Chunk + 64;

    if (warp == 8) {
        // ---------------------------------------------------------------- TMA + MMA issuer
        if (lane == 0) {
            const uint32_t aq = smem_u32(sQ), ak = smem_u32(sK), av = smem_u32(sV);
            const uint32_t idesc_o = make_idesc_bf16(128, kHeadDim, 0, 1);
            mbar_expect_tx(&bars->q, 16384);
            tma_load_3d(sQ, &tma_q, &bars->q, h * kHeadDim, q0, b);
            mbar_expect_tx(&bars->k, 16384);
            tma_load_3d(sK, &tma_k, &bars->k, h * kHeadDim, 0, b);
            mbar_expect_tx(&bars->v, 16384);
            tma_load_3d(sV, &tma_v, &bars->v, h * kHeadDim, 0, b);
            // O_j = P_j . V_j (fresh tile), then V / P buffers are free: fetch the next V chunk
            auto issue_pv = [&](int j) {
                const int n16 = min(kFwd2Chunk, p.lk16 - j * kFwd2Chunk);
                mbar_wait(&bars->p, j & 1);
                mbar_wait(&bars->v, j & 1);
                tc_fence_after();
                for (int k = 0; k < n16; k += 16)
                    umma_bf16_ts(t_o, t_p + (k >> 1), make_smem_desc(av + k * 128, 8192, 1024), idesc_o, k > 0);
                tc_commit(&bars->pv);
                mbar_wait(&bars->pv, j & 1);
                if (j + 1 < n_ch) {
                    mbar_expect_tx(&bars->v, 16384);
                    tma_load_3d(sV, &tma_v, &bars->v, h * kHeadDim, (j + 1) * kFwd2Chunk, b);
                }
            };
            mbar_wait(&bars->q, 0);
            for (int c = 0; c < n_ch; ++c) {
                const int n16 = min(kFwd2Chunk, p.lk16 - c * kFwd2Chunk);
                mbar_wait(&bars->k, c & 1);
                if (c > 0) mbar_wait(&bars->sfree, (c - 1) & 1);       // every softmax warp is done with S_{c-1}
                tc_fence_after();
                const uint32_t idesc_s = make_idesc_bf16(128, n16, 0, 0);
#pragma unroll
                for (int ks = 0; ks < 4; ++ks)
                    umma_bf16(t_s, make_smem_desc(aq + ks * 32, 16, 1024), make_smem_desc(ak + ks * 32, 16, 1024),
                              idesc_s, ks > 0);
                tc_commit(&bars->s);
                if (c > 0) issue_pv(c - 1);          // after S_c: the softmax of chunk c runs under this product
                mbar_wait(&bars->s, c & 1);           // S_c complete: the K buffer is free
                if (c + 1 < n_ch) {
                    mbar_expect_tx(&bars->k, 16384);
                    tma_load_3d(sK, &tma_k, &bars->k, h * kHeadDim, (c + 1) * kFwd2Chunk, b);
                }
            }
            issue_pv(n_ch - 1);
        }
    } else {
        // ---------------------------------------------------------------- softmax warps
        const int quarter = warp & 3, half = warp >> 2;
        const int row = quarter * 32 + lane;
        const int q = q0 + row;
        const uint32_t t_row = (uint32_t(quarter * 32) << 16);
        const bool rows_live = q0 + quarter * 32 < p.lq;       // warps past Lq (393 = 3 x 128 + 9) only keep the barriers going
        const int k_lim = p.mask_mode == 2 ? min(p.lk, (q / p.q_per_frame + 1) * p.k_per_frame) : p.lk;
        const float c1 = p.scale * 1.4426950408889634f;
        uint32_t dkey = 0, dth = 0, row_pair0 = 0;
        if (DROP) {
            dkey = drop_key(__ldg(p.drop.seed), p.drop.salt), dth = drop_thresh(p.drop.p);
            row_pair0 = attn_row_pair0(b, p.heads, h, p.lq, q, p.lk);
        }
        if (!rows_live) {
            // rows past Lq are never written below; the PV product still reads them: keep them finite (zero), so that no
            // NaN is left behind in TMEM for whoever allocates these columns next
            uint32_t z[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) z[j] = 0u;
            tmem_st16(t_p + t_row + half * 32, z);
            tmem_st16(t_p + t_row + half * 32 + 16, z);
            tmem_st_wait();
        }
        float m_run = -INFINITY, l_run = 0.f, alpha_prev = 0.f;
        float o_acc[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) o_acc[i] = 0.f;
        for (int c = 0; c < n_ch; ++c) {
            const int kc0 = c * kFwd2Chunk;                       // first key of the chunk
            const int n16 = min(kFwd2Chunk, p.lk16 - kc0);        // columns the MMA produced
            const int col0 = half * 64;                           // my columns of the chunk: [col0, col0 + 64)
            const int n_mine = rows_live ? max(0, min(64, n16 - col0)) : 0;   // of those, how many exist (multiple of 16)
            // dropout keep bits of my 64 columns, hashed BEFORE the wait for the scores: pure ALU work that fills issue
            // slots the warp would otherwise spend stalled on the barrier (the hash is ~60 % of the pass's instructions)
            uint32_t keep_w[2] = {0xffffffffu, 0xffffffffu};
            if (DROP) {
#pragma unroll
                for (int w = 0; w < 2; ++w) {
                    uint32_t bits = 0;
                    if (w * 32 < n_mine) {
                        const uint32_t pb = row_pair0 + ((kc0 + col0 + w * 32) >> 1);
#pragma unroll
                        for (int i = 0; i < 16; ++i) bits |= drop_pair(dkey, pb + i, dth) << (2 * i);
                    }
                    keep_w[w] = bits;
                }
            }
            mbar_wait(&bars->s, c & 1);
            tc_fence_after();
            // pass A: row maximum over my columns
            float mx = -INFINITY;
            for (int g = 0; g < n_mine; g += 32) {
                uint32_t raw[32];
                tmem_ld32(tmem + t_row + col0 + g, raw);
                tmem_ld_wait();
                const int k0 = kc0 + col0 + g;
                if (!HAS_KMASK && k0 + 32 <= k_lim) {        // whole group visible, no mask: max commutes with the scale
                    float m2 = -INFINITY;
#pragma unroll
                    for (int i = 0; i < 32; ++i) m2 = fmaxf(m2, __uint_as_float(raw[i]));
                    mx = fmaxf(mx, m2 * c1);
                } else {
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        float t = __uint_as_float(raw[i]) * c1;
                        if (HAS_KMASK) t += s_kmask[min(k0 + i, 447)];
                        mx = fmaxf(mx, k0 + i < k_lim ? t : -INFINITY);
                    }
                }
            }
            s_red[half * 128 + row] = mx;
            asm volatile("bar.sync 1, 256;" ::: "memory");
            mx = fmaxf(s_red[row], s_red[128 + row]);
            const float m_new = fmaxf(m_run, mx);
            const float m_use = m_new == -INFINITY ? 0.f : m_new;
            const float alpha = m_run == -INFINITY ? 0.f : ex2_approx(m_run - m_use);
            // O_{c-1} (relative to M_{c-1}) joins the accumulator; the P / V buffers and the O tile are then free
            if (c > 0) {
                mbar_wait(&bars->pv, (c - 1) & 1);
                tc_fence_after();
                if (rows_live) {
                    uint32_t ro[32];
                    tmem_ld32(t_o + t_row + half * 32, ro);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 32; ++i) o_acc[i] = fmaf(o_acc[i], alpha_prev, __uint_as_float(ro[i]));
                }
            }
            alpha_prev = alpha;
            // pass B: probabilities of my columns -> bf16 -> P block `half` (row = 128 B, 16-byte units swizzled)
            float sum = 0.f;
            for (int g = 0; g < n_mine; g += 32) {
                uint32_t raw[32];
                tmem_ld32(tmem + t_row + col0 + g, raw);
                tmem_ld_wait();
                float pv[32];
                const int k0 = kc0 + col0 + g;
                if (k0 + 32 <= k_lim) {                   // whole group visible: no per-element predicate
#pragma unroll
                    for (int i = 0; i < 32; ++i)
                        pv[i] = ex2_approx(HAS_KMASK ? fmaf(__uint_as_float(raw[i]), c1, s_kmask[min(k0 + i, 447)] - m_use)
                                                     : fmaf(__uint_as_float(raw[i]), c1, -m_use));
                } else {
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        float t = fmaf(__uint_as_float(raw[i]), c1, -m_use);
                        if (HAS_KMASK) t += s_kmask[min(k0 + i, 447)];
                        pv[i] = k0 + i < k_lim ? ex2_approx(t) : 0.f;
                    }
                }
                const uint32_t keep_bits = keep_w[g >> 5];
                uint32_t pk[16];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    uint4 u;
                    u.x = pack_bf16x2(pv[8 * j], pv[8 * j + 1]), u.y = pack_bf16x2(pv[8 * j + 2], pv[8 * j + 3]);
                    u.z = pack_bf16x2(pv[8 * j + 4], pv[8 * j + 5]), u.w = pack_bf16x2(pv[8 * j + 6], pv[8 * j + 7]);
                    sum += (__uint_as_float(u.x << 16) + __uint_as_float(u.x & 0xffff0000u)) +
                           (__uint_as_float(u.y << 16) + __uint_as_float(u.y & 0xffff0000u)) +
                           (__uint_as_float(u.z << 16) + __uint_as_float(u.z & 0xffff0000u)) +
                           (__uint_as_float(u.w << 16) + __uint_as_float(u.w & 0xffff0000u));
                    if (DROP) {
                        // bits 2i, 2i+1 of the word = keep flags of the packed pair i
                        const uint32_t kb = keep_bits >> (8 * j);
                        u.x &= (kb & 1u) * 0xffffu + ((kb >> 1) & 1u) * 0xffff0000u;
                        u.y &= ((kb >> 2) & 1u) * 0xffffu + ((kb >> 3) & 1u) * 0xffff0000u;
                        u.z &= ((kb >> 4) & 1u) * 0xffffu + ((kb >> 5) & 1u) * 0xffff0000u;
                        u.w &= ((kb >> 6) & 1u) * 0xffffu + ((kb >> 7) & 1u) * 0xffff0000u;
                    }
                    pk[4 * j] = u.x, pk[4 * j + 1] = u.y, pk[4 * j + 2] = u.z, pk[4 * j + 3] = u.w;
                }
                tmem_st16(t_p + t_row + half * 32 + (g >> 1), pk);
                if (DROP && p.drop_bits && q < p.lq)
                    p.drop_bits[(((size_t)b * p.heads + h) * p.lq + q) * (p.lk32 >> 5) + ((kc0 + col0 + g) >> 5)] = keep_bits;
            }
            l_run = fmaf(l_run, alpha, sum);
            m_run = m_new;
            // S_c is consumed and P_c is stored (both TMEM): hand both to the issuer
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(&bars->sfree);
                mbar_arrive(&bars->p);
            }
        }
        // last PV product
        mbar_wait(&bars->pv, (n_ch - 1) & 1);
        tc_fence_after();
        if (rows_live) {
            uint32_t ro[32];
            tmem_ld32(t_o + t_row + half * 32, ro);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) o_acc[i] = fmaf(o_acc[i], alpha_prev, __uint_as_float(ro[i]));
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");          // s_red: everyone has read the last chunk's maxima
        s_red[half * 128 + row] = l_run;
        asm volatile("bar.sync 1, 256;" ::: "memory");
        const float sum = s_red[row] + s_red[128 + row];
        const float inv = DROP ? 1.0f / (sum * (1.0f - p.drop.p)) : 1.0f / sum;
        if (rows_live && q < p.lq) {
            __nv_bfloat16* dst = p.o + ((size_t)b * p.lq + q) * p.ldo + h * kHeadDim + half * 32;
#pragma unroll
            for (int i = 0; i < 32; i += 8) {
                uint4 u;
                u.x = pack_bf16x2(o_acc[i] * inv, o_acc[i + 1] * inv), u.y = pack_bf16x2(o_acc[i + 2] * inv, o_acc[i + 3] * inv);
                u.z = pack_bf16x2(o_acc[i + 4] * inv, o_acc[i + 5] * inv), u.w = pack_bf16x2(o_acc[i + 6] * inv, o_acc[i + 7] * inv);
                *reinterpret_cast<uint4*>(dst + i) = u;
            }
            if (p.o_f32) {
                float* df = p.o_f32 + ((size_t)b * p.lq + q) * p.ldof + h * kHeadDim + half * 32;
#pragma unroll
                for (int i = 0; i < 32; i += 4)
                    *reinterpret_cast<float4*>(df + i) =
                        make_float4(o_acc[i] * inv, o_acc[i + 1] * inv, o_acc[i + 2] * inv, o_acc[i + 3] * inv);
            }
            // natural-log log-sum-exp of the scaled, masked scores
            if (p.lse && half == 0)
                p.lse[((size_t)b * p.heads + h) * p.lq + q] = (m_run + log2f(sum)) * 0.6931471805599453f;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 8) {
        tc_fence_after();
        tmem_dealloc(tmem, 256);
    }
}

// =====================================================================================================
// Backward.  Two passes, each recomputing the probabilities from Q, K and the saved log-sum-exp:
//   dq pass  : CTA = (128-query tile, head, batch); loops over 128-key chunks;   dQ  = scale * dS . K
//   dkv pass : CTA = (128-key tile,  head, batch); loops over 128-query chunks;  dV  = P^T . dO,
//                                                                                dK  = scale * dS^T . Q
// with dS = P o (dP - delta), dP = dO . V^T, delta = rowsum(dO o O).  Per chunk: two tcgen05 products into
// TMEM (S and dP, or their transposes), 256 threads turn them into bf16 P / dS tiles in 128B-swizzled shared
// memory (thread = one row, a quarter of the columns), and those tiles feed the accumulating products as K-major A
// operands while the TMA-loaded K / Q / dO tiles are consumed as MN-major B operands in place.
// =====================================================================================================
constexpr int kBwdThreads = 512;
constexpr int kBwdTileBytes = 196608;   // operand / P / dS tiles of either backward kernel (3-stage streamed operands)   // 16 warps: four threads per tile row, one 32-column chunk each
constexpr float kLog2e = 1.4426950408889634f;

struct AttnBwdParams {
    const float* o_f32;        // fp32 forward output [batch*lq, ldof] or null (then delta comes from a pre-sweep)
    int64_t ldof;
    const __nv_bfloat16* o;    // forward output  [batch*lq, ldo]
    const __nv_bfloat16* d_o;  // its gradient    [batch*lq, lddo]
    int64_t ldo, lddo;
    const float* lse;          // [batch, heads, lq]
    float* delta;              // [batch, heads, lq]  (written by the dq pass, read by the dkv pass)
    __nv_bfloat16* dq;
    __nv_bfloat16* dk;
    __nv_bfloat16* dv;
    int64_t lddq, lddk, lddv;
    const float* key_mask;
    int batch, heads, lq, lk;
    float scale;
    int mask_mode, q_per_frame, k_per_frame;
    DropCfg drop;
    const uint32_t* drop_bits;   // the forward's keep-bit words [batch, heads, lq, lk32/32], or null (re-hash)
};

// 32 consecutive bf16 of row `row` at column c (multiple of 32) of a [128 x 128] K-major tile made of two
// [128 x 64] 128B-swizzled blocks.
__device__ __forceinline__ void store_tile_chunk32(uint8_t* tile, int row, int c, const float (&v)[32]) {
    uint8_t* blk = tile + (c >> 6) * 16384 + row * 128;
    const int j0 = (c & 32) ? 4 : 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        uint4 u;
        u.x = pack_bf16x2(v[8 * j], v[8 * j + 1]), u.y = pack_bf16x2(v[8 * j + 2], v[8 * j + 3]);
        u.z = pack_bf16x2(v[8 * j + 4], v[8 * j + 5]), u.w = pack_bf16x2(v[8 * j + 6], v[8 * j + 7]);
        *reinterpret_cast<uint4*>(blk + (((j0 + j) ^ (row & 7)) << 4)) = u;
    }
}

// same, from values already packed as bf16 pairs (pk[i] = elements 2i, 2i+1)
__device__ __forceinline__ void store_tile_chunk32_packed(uint8_t* tile, int row, int c, const uint32_t (&pk)[16]) {
    uint8_t* blk = tile + (c >> 6) * 16384 + row * 128;
    const int j0 = (c & 32) ? 4 : 0;
#pragma unroll
    for (int j = 0; j < 4; ++j)
        *reinterpret_cast<uint4*>(blk + (((j0 + j) ^ (row & 7)) << 4)) =
            make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
}

// K-major A descriptor for k-step ks (16 columns) of such a tile; MN-major B descriptor for rows [16ks,16ks+16)
// of a [128 x 64] TMA tile.
__device__ __forceinline__ uint64_t tile_a_desc(uint32_t tile, int ks) {
    return make_smem_desc(tile + (ks >> 2) * 16384 + (ks & 3) * 32, 16, 1024);
}
__device__ __forceinline__ uint64_t rows_b_desc(uint32_t tile, int ks) {
    return make_smem_desc(tile + ks * 2048, 8192, 1024);
}


// ---------------------------------------------------------------------------------------------------------
// Per-thread math of one 32-column group, specialised at compile time so that the common case (no key mask,
// every column visible) is 4 issue slots per element: FFMA, MUFU.EX2, FFMA, FMUL (+ the bf16 pack).
// ---------------------------------------------------------------------------------------------------------
// dq pass: ds_e = p_e * (dP_e * scale - delta * scale),  p_e = 2^(s_e * c1 [+ kmask_e] - lse)
// DROP: 0 = no dropout, 1 = mask regenerated from the counter hash, 2 = mask read from the forward's keep-bit word
template <bool KMASK, int DROP, bool VIS, int N = 32>
__device__ __forceinline__ void dq_group_math(const uint32_t (&rs)[N], const uint32_t (&rp)[N], uint32_t (&ds)[N / 2],
                                              float c1, float neg_lse, const float* km, int n_vis, float scale_keep,
                                              float nds, uint32_t dkey, uint32_t pair0, uint32_t dth, uint32_t bits) {
#pragma unroll
    for (int i = 0; i < N; i += 2) {
        uint32_t dmask = 0xffffffffu;
        if (DROP == 1) dmask = drop_pair_mask(dkey, pair0 + (i >> 1), dth);
        float v2[2];
#pragma unroll
        for (int t = 0; t < 2; ++t) {
            const int e = i + t;
            float x = fmaf(__uint_as_float(rs[e]), c1, neg_lse);
            if (KMASK) x += km[e];
            const float pr = ex2_approx(x);
            // o = dropout(P) V: the gradient reaches P_ij only through a kept entry, scaled by 1/(1-p) (in scale_keep)
            float g = fmaf(__uint_as_float(rp[e]), scale_keep, nds);
            if (DROP == 1) g = (dmask & (t ? 0xffff0000u : 0x0000ffffu)) ? g : nds;
            if (DROP == 2) g = (bits >> e) & 1u ? g : nds;
            const float v = pr * g;
            v2[t] = (VIS || e < n_vis) ? v : 0.f;
        }
        ds[i >> 1] = pack_bf16x2(v2[0], v2[1]);
    }
}

// merged pass (single query tile): ds_e as in the dq pass plus pd_e = dropout(p)_e / (1-p), the operand of dV
template <bool KMASK, int DROP, bool VIS>
__device__ __forceinline__ void dqkv_group_math(const uint32_t (&rs)[32], const uint32_t (&rp)[32], uint32_t (&ds)[16],
                                                uint32_t (&pd)[16], float c1, float neg_lse, const float* km, int n_vis,
                                                float scale_keep, float nds, float inv_keep, uint32_t dkey, uint32_t pair0,
                                                uint32_t dth, uint32_t bits) {
#pragma unroll
    for (int i = 0; i < 32; i += 2) {
        uint32_t dmask = 0xffffffffu;
        if (DROP == 1) dmask = drop_pair_mask(dkey, pair0 + (i >> 1), dth);
        float v2[2], p2[2];
#pragma unroll
        for (int t = 0; t < 2; ++t) {
            const int e = i + t;
            float x = fmaf(__uint_as_float(rs[e]), c1, neg_lse);
            if (KMASK) x += km[e];
            float pr = ex2_approx(x);
            if (!VIS) pr = e < n_vis ? pr : 0.f;
            float g = fmaf(__uint_as_float(rp[e]), scale_keep, nds);
            bool keep = true;
            if (DROP == 1) keep = (dmask & (t ? 0xffff0000u : 0x0000ffffu)) != 0;
            if (DROP == 2) keep = ((bits >> e) & 1u) != 0;
            if (DROP) g = keep ? g : nds;
            v2[t] = pr * g;
            p2[t] = DROP ? (keep ? pr * inv_keep : 0.f) : pr;
        }
        ds[i >> 1] = pack_bf16x2(v2[0], v2[1]);
        pd[i >> 1] = pack_bf16x2(p2[0], p2[1]);
    }
}

// dkv pass (thread = key row, 32 queries): p_e as above with per-query lse, pd_e = dropout(p)_e (-> dV), ds_e (-> dK)
template <int DROP, bool VIS, int N = 32>
__device__ __forceinline__ void dkv_group_math(const uint32_t (&rs)[N], const uint32_t (&rp)[N], uint32_t (&pr)[N / 2],
                                               uint32_t (&ds)[N / 2], float c1, float kbias, const float* lse, const float* del,
                                               int e_first, float scale_keep, float inv_keep, uint32_t dkey,
                                               uint32_t pair0, uint32_t pair_stride, uint32_t dth, uint32_t lane_mask,
                                               const uint32_t* bits, int bit) {
#pragma unroll
    for (int e = 0; e < N; e += 4) {
        const float4 l4 = *reinterpret_cast<const float4*>(lse + e);
        const float4 d4 = *reinterpret_cast<const float4*>(del + e);
        uint32_t bw[4] = {0, 0, 0, 0};
        if (DROP == 2) {
            const uint4 w4 = *reinterpret_cast<const uint4*>(bits + e);     // keep words of queries e..e+3 (this key's word)
            bw[0] = w4.x, bw[1] = w4.y, bw[2] = w4.z, bw[3] = w4.w;
        }
        const float lv[4] = {l4.x, l4.y, l4.z, l4.w}, dv4[4] = {d4.x, d4.y, d4.z, d4.w};
        float po[4], so[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const float x = fmaf(__uint_as_float(rs[e + t]), c1, kbias - lv[t]);
            float pe = ex2_approx(x);
            if (!VIS) pe = (e + t >= e_first) ? pe : 0.f;
            float g = fmaf(__uint_as_float(rp[e + t]), scale_keep, -dv4[t]), pd = pe;
            if (DROP) {
                const bool keep = DROP == 2 ? ((bw[t] >> bit) & 1u) != 0
                                            : (drop_pair_mask(dkey, pair0 + (uint32_t)(e + t) * pair_stride, dth) & lane_mask) != 0;
                g = keep ? g : -dv4[t];
                pd = keep ? pe * inv_keep : 0.f;          // dV = dropout(P)^T dO
            }
            po[t] = pd;
            so[t] = pe * g;
        }
        pr[e >> 1] = pack_bf16x2(po[0], po[1]), pr[(e >> 1) + 1] = pack_bf16x2(po[2], po[3]);
        ds[e >> 1] = pack_bf16x2(so[0], so[1]), ds[(e >> 1) + 1] = pack_bf16x2(so[2], so[3]);
    }
}

// MERGE (problems with ONE query tile, lq <= 128, o_f32 given): the same pass also produces dK and dV of every key
// chunk — dV_j = dropout(P)_j^T dO and dK_j = dS_j^T Q are complete after a single query tile, so the separate dkv
// pass (which recomputes S, dP and the exponentials per key tile) is not launched.  The bf16 P / dS tiles written for
// the K-major dQ product are read as MN-major A operands (M = keys) by the two extra products: same bytes, no
// transposed copy.  dK_j / dV_j leave TMEM at the start of the next iteration.
template <int DROP, bool MERGE>
__global__ void __launch_bounds__(kBwdThreads)
attn_bwd_dq_kernel(const __grid_constant__ CUtensorMap tma_q, const __grid_constant__ CUtensorMap tma_do,
                   const __grid_constant__ CUtensorMap tma_k, const __grid_constant__ CUtensorMap tma_v,
                   const AttnBwdParams p) {
    // delta_i = sum_j P_ij dP_ij must be consistent with the dP that build dS to fp32 accuracy: when the value rows of
    // a head are nearly identical (the decoder states of this model are) dP - delta cancels to ~1e-3 of dP, and the
    // textbook delta = rowsum(dO o O) on the bf16-rounded O is then pure rounding noise.  Two exact routes:
    //   * o_f32 given (the forward's fp32 copy of O): delta = rowsum(dO o O_f32), one sweep over the key chunks;
    //   * otherwise: a pre-sweep accumulates sum(p) and sum(p * dP) per row, delta = their ratio (2 sweeps).
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sQ = smem;                       // 16 KB
    uint8_t* sdO = smem + 16384;              // 16 KB
    uint8_t* sK = smem + 32768;               // 3 x 16 KB
    uint8_t* sV = smem + 81920;               // 3 x 16 KB
    uint8_t* sdS = smem + 131072;             // 2 x 32 KB (MERGE: one dS tile, then one P tile)
    uint8_t* sPd = smem + 131072 + 32768;
    float* s_part = reinterpret_cast<float*>(smem + kBwdTileBytes);   // [4 quarters][128 rows][2], then [512] key mask
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_part + 1024 + 512);
    uint64_t* bar_q = bars;          // Q + dO landed
    uint64_t* bar_kv = bars + 1;     // [3]
    uint64_t* bar_s = bars + 4;      // S and dP of the current chunk complete
    uint64_t* bar_dq = bars + 5;     // [2] dQ product retired (frees its dS tile and K/V buffer)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 7);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int q0 = blockIdx.x * 128, h = blockIdx.y, b = blockIdx.z;
    const int n_kc = (p.lk + 127) / 128;
    const int n_pre = p.o_f32 ? 0 : n_kc;      // iterations of the delta pre-sweep
    const int n_it = n_pre + n_kc;

    if (tid == 0) {
        mbar_init(bar_q, 1);
        mbar_init(bar_kv, 1), mbar_init(bar_kv + 1, 1), mbar_init(bar_kv + 2, 1);
        mbar_init(bar_s, 1);
        mbar_init(bar_dq, 1), mbar_init(bar_dq + 1, 1);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t t_s = tmem, t_dp = tmem + 128, t_dq = tmem + 256, t_dv = tmem + 320, t_dk = tmem + 384;
    const uint32_t idesc_dq = make_idesc_bf16(128, kHeadDim, 0, 1);
    const uint32_t idesc_kv = make_idesc_bf16(128, kHeadDim, 1, 1);     // A = P^T / dS^T (MN-major), B = dO / Q rows
    if (MERGE) {
        // rows of the P / dS tiles that no thread writes (query rows beyond lq) are read by the dK / dV products
        // against zero-filled dO / Q rows: keep them finite
        for (int i = tid; i < 65536 / 16; i += kBwdThreads) reinterpret_cast<uint4*>(sdS)[i] = make_uint4(0, 0, 0, 0);
        fence_proxy_async_smem();
        __syncthreads();
    }
    // keys of chunk j that exist, rounded up to the MMA granule: the products of the last chunk are only that wide
    auto chunk_n16 = [&](int j) { return (min(128, p.lk - j * 128) + 15) & ~15; };

    auto issue_s = [&](int jj) {   // S = Q K_j^T ; dP = dO V_j^T   (operands of iteration jj sit in buffer jj % 3)
        const uint32_t aq = smem_u32(sQ), ado = smem_u32(sdO);
        const uint32_t bk = smem_u32(sK + (jj % 3) * 16384), bv = smem_u32(sV + (jj % 3) * 16384);
        const uint32_t idesc_s = make_idesc_bf16(128, chunk_n16(jj < n_pre ? jj : jj - n_pre), 0, 0);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
            umma_bf16(t_s, make_smem_desc(aq + ks * 32, 16, 1024), make_smem_desc(bk + ks * 32, 16, 1024), idesc_s, ks > 0);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
            umma_bf16(t_dp, make_smem_desc(ado + ks * 32, 16, 1024), make_smem_desc(bv + ks * 32, 16, 1024), idesc_s, ks > 0);
        tc_commit(bar_s);
    };
    auto load_kv = [&](int jj) {
        const int j = jj < n_pre ? jj : jj - n_pre, st = jj % 3;
        mbar_expect_tx(bar_kv + st, 32768);
        tma_load_3d(sK + st * 16384, &tma_k, bar_kv + st, h * kHeadDim, j * 128, b);
        tma_load_3d(sV + st * 16384, &tma_v, bar_kv + st, h * kHeadDim, j * 128, b);
    };

    if (tid == 0) {
        mbar_expect_tx(bar_q, 32768);
        tma_load_3d(sQ, &tma_q, bar_q, h * kHeadDim, q0, b);
        tma_load_3d(sdO, &tma_do, bar_q, h * kHeadDim, q0, b);
        load_kv(0);
        if (n_it > 1) load_kv(1);
        if (n_it > 2) load_kv(2);
        mbar_wait(bar_q, 0);
        mbar_wait(bar_kv, 0);
        tc_fence_after();
        issue_s(0);
    }
    __syncwarp();

    const int row = (warp & 3) * 32 + lane, quad = warp >> 2;   // quad: which 32-column chunk of the 128
    const int q = q0 + row;
    const uint32_t t_row = uint32_t((warp & 3) * 32) << 16;
    const float* kmask = p.mask_mode == 1 ? p.key_mask + (size_t)b * p.lk : nullptr;
    const int k_lim = p.mask_mode == 2 ? min(p.lk, (q / p.q_per_frame + 1) * p.k_per_frame) : p.lk;
    const float c1 = p.scale * kLog2e;
    const float my_lse = q < p.lq ? p.lse[((size_t)b * p.heads + h) * p.lq + q] * kLog2e : INFINITY;
    float acc_p = 0.f, acc_pd = 0.f, neg_delta_scaled = 0.f;
    uint32_t dkey = 0, dth = 0, row_pair0 = 0;
    float inv_keep = 1.f;
    if (DROP) {
        dkey = drop_key(__ldg(p.drop.seed), p.drop.salt), dth = drop_thresh(p.drop.p);
        row_pair0 = attn_row_pair0(b, p.heads, h, p.lq, q, p.lk), inv_keep = 1.f / (1.f - p.drop.p);
    }
    // additive key mask of this batch element, pre-scaled by log2(e), staged once in shared memory
    float* s_kmask = s_part + 1024;
    if (kmask) {
        for (int k = tid; k < 512; k += kBwdThreads) s_kmask[k] = k < p.lk ? kmask[k] * kLog2e : 0.f;
        __syncthreads();
    }

    if (p.o_f32) {
        // delta = rowsum(dO o O_f32): each of the row's four threads takes 16 of the 64 head columns
        float part = 0.f;
        if (q < p.lq) {
            const float4* po = reinterpret_cast<const float4*>(p.o_f32 + ((size_t)b * p.lq + q) * p.ldof + h * kHeadDim + quad * 16);
            const uint4* pd = reinterpret_cast<const uint4*>(p.d_o + ((size_t)b * p.lq + q) * p.lddo + h * kHeadDim + quad * 16);
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const uint4 g = __ldg(pd + i);
                const float4 o0 = __ldg(po + 2 * i), o1 = __ldg(po + 2 * i + 1);
                float2 y;
                y = unpack_bf16x2(g.x), part = fmaf(o0.x, y.x, fmaf(o0.y, y.y, part));
                y = unpack_bf16x2(g.y), part = fmaf(o0.z, y.x, fmaf(o0.w, y.y, part));
                y = unpack_bf16x2(g.z), part = fmaf(o1.x, y.x, fmaf(o1.y, y.y, part));
                y = unpack_bf16x2(g.w), part = fmaf(o1.z, y.x, fmaf(o1.w, y.y, part));
            }
        }
        s_part[quad * 128 + row] = part;
        __syncthreads();
        const float my_delta = (s_part[row] + s_part[128 + row]) + (s_part[256 + row] + s_part[384 + row]);
        neg_delta_scaled = -my_delta * p.scale;
        if (quad == 0 && q < p.lq) p.delta[((size_t)b * p.heads + h) * p.lq + q] = my_delta;
        __syncthreads();
    }

    // Per key chunk: S/dP (TMEM) -> registers -> [the next chunk's S/dP products are issued here, so the tensor core
    // works through them while the threads do the exponentials] -> dS tile -> dQ += dS K.
    // Warps whose 32 query rows lie beyond lq (the last query tile: 393 = 3 x 128 + 9) and column groups beyond the
    // chunk's keys do no math at all: their dS entries feed only dQ rows that are never stored / k-steps never issued.
    const bool rows_live = q0 + (warp & 3) * 32 < p.lq;
    const float scale_keep = p.scale * inv_keep;
    // MERGE: dK / dV of key chunk j (TMEM lanes = keys) -> global; thread = (key row, 16 of the 64 head columns)
    auto store_dkv = [&](int j) {
        const int key = j * 128 + row;
        uint32_t rk[16], rv[16];
        tmem_ld16(t_dk + t_row + quad * 16, rk);
        tmem_ld16(t_dv + t_row + quad * 16, rv);
        tmem_ld_wait();
        if (key < p.lk) {
            __nv_bfloat16* dk = p.dk + ((size_t)b * p.lk + key) * p.lddk + h * kHeadDim + quad * 16;
            __nv_bfloat16* dv = p.dv + ((size_t)b * p.lk + key) * p.lddv + h * kHeadDim + quad * 16;
#pragma unroll
            for (int i = 0; i < 16; i += 8) {
                uint4 u;
                u.x = pack_bf16x2(__uint_as_float(rk[i]), __uint_as_float(rk[i + 1]));
                u.y = pack_bf16x2(__uint_as_float(rk[i + 2]), __uint_as_float(rk[i + 3]));
                u.z = pack_bf16x2(__uint_as_float(rk[i + 4]), __uint_as_float(rk[i + 5]));
                u.w = pack_bf16x2(__uint_as_float(rk[i + 6]), __uint_as_float(rk[i + 7]));
                *reinterpret_cast<uint4*>(dk + i) = u;
                u.x = pack_bf16x2(__uint_as_float(rv[i]), __uint_as_float(rv[i + 1]));
                u.y = pack_bf16x2(__uint_as_float(rv[i + 2]), __uint_as_float(rv[i + 3]));
                u.z = pack_bf16x2(__uint_as_float(rv[i + 4]), __uint_as_float(rv[i + 5]));
                u.w = pack_bf16x2(__uint_as_float(rv[i + 6]), __uint_as_float(rv[i + 7]));
                *reinterpret_cast<uint4*>(dv + i) = u;
            }
        }
    };
    for (int jj = 0; jj < n_it; ++jj) {
        const bool sweep2 = jj >= n_pre;
        const int j = sweep2 ? jj - n_pre : jj;     // key chunk
        const int j2 = jj - n_pre;                  // index within the main sweep
        const int n16 = chunk_n16(j);
        const int c = quad * 32;
        const bool live = rows_live && c < n16;     // warp-uniform
        uint32_t keep_word = 0;
        if (DROP == 2 && live && q < p.lq)          // requested ahead of the wait for the score products
            keep_word = __ldg(p.drop_bits + (((size_t)b * p.heads + h) * p.lq + q) * (size_t)((p.lk + 31) >> 5) + j * 4 + quad);
        mbar_wait(bar_s, jj & 1);
        tc_fence_after();
        uint32_t rs[32], rp[32];
        if (live) {
            tmem_ld32(t_s + t_row + c, rs);
            tmem_ld32(t_dp + t_row + c, rp);
            tmem_ld_wait();
        }
        if (MERGE && jj >= 1) {
            // products of the previous chunk retired: its dK / dV leave TMEM, its P / dS tiles and K/V buffer are free
            mbar_wait(bar_dq + ((jj - 1) & 1), ((jj - 1) >> 1) & 1);
            tc_fence_after();
            store_dkv(jj - 1);
        }
        tc_fence_before();
        __syncthreads();                            // every thread holds its S/dP values: the TMEM tiles are free
        if (tid == 0) {
            tc_fence_after();
            if (jj + 1 < n_it) {
                mbar_wait(bar_kv + (jj + 1) % 3, ((jj + 1) / 3) & 1);
                tc_fence_after();
                issue_s(jj + 1);
            }
            if (jj >= 1 && jj + 2 < n_it) {
                // buffer (jj - 1) % 3 is free once its last reader retired: the S/dP products (bar_s, observed one
                // iteration ago) in the pre-sweep, the dQ product in the main sweep
                if (jj - 1 >= n_pre) mbar_wait(bar_dq + ((jj - 1 - n_pre) & 1), ((jj - 1 - n_pre) >> 1) & 1);
                load_kv(jj + 2);
            }
        }
        uint8_t* tile = MERGE ? sdS : sdS + (j2 & 1) * 32768;
        if (live) {
            const int key0 = j * 128 + c;
            const int n_vis = k_lim - key0;            // visible columns of this group (may be <= 0 or >= 32)
            if (!sweep2) {
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    float x = fmaf(__uint_as_float(rs[i]), c1, -my_lse);
                    if (kmask) x += s_kmask[key0 + i];
                    const float pr = i < n_vis ? ex2_approx(x) : 0.f;
                    float dp = __uint_as_float(rp[i]);
                    if (DROP == 1) dp = drop_keep(dkey, 2ull * row_pair0 + (uint32_t)(key0 + i), dth) ? dp * inv_keep : 0.f;
                    if (DROP == 2) dp = (keep_word >> i) & 1u ? dp * inv_keep : 0.f;
                    acc_p += pr;
                    // columns past the chunk's keys were never written by the S / dP products (stale TMEM, possibly
                    // NaN): 0 * NaN must not reach the row sum
                    acc_pd = fmaf(pr, i < n_vis ? dp : 0.f, acc_pd);
                }
            } else {
                uint32_t ds[16];
                const uint32_t pair0 = row_pair0 + (uint32_t)(key0 >> 1);
                const float* km = s_kmask + key0;
                if (MERGE) {
                    uint32_t pd[16];
                    if (__all_sync(0xffffffffu, n_vis >= 32)) {
                        if (kmask) dqkv_group_math<true, DROP, true>(rs, rp, ds, pd, c1, -my_lse, km, n_vis, scale_keep, neg_delta_scaled, inv_keep, dkey, pair0, dth, keep_word);
                        else dqkv_group_math<false, DROP, true>(rs, rp, ds, pd, c1, -my_lse, km, n_vis, scale_keep, neg_delta_scaled, inv_keep, dkey, pair0, dth, keep_word);
                    } else {
                        if (kmask) dqkv_group_math<true, DROP, false>(rs, rp, ds, pd, c1, -my_lse, km, n_vis, scale_keep, neg_delta_scaled, inv_keep, dkey, pair0, dth, keep_word);
                        else dqkv_group_math<false, DROP, false>(rs, rp, ds, pd, c1, -my_lse, km, n_vis, scale_keep, neg_delta_scaled, inv_keep, dkey, pair0, dth, keep_word);
                    }
                    store_tile_chunk32_packed(sPd, row, c, pd);
                } else if (__all_sync(0xffffffffu, n_vis >= 32)) {
                    if (kmask) dq_group_math<true, DROP, true>(rs, rp, ds, c1, -my_lse, km, n_vis, scale_keep, neg_delta_scaled, dkey, pair0, dth, keep_word);
                    else dq_group_math<false, DROP, true>(rs, rp, ds, c1, -my_lse, km, n_vis, scale_keep, neg_delta_scaled, dkey, pair0, dth, keep_word);
                } else {
                    if (kmask) dq_group_math<true, DROP, false>(rs, rp, ds, c1, -my_lse, km, n_vis, scale_keep, neg_delta_scaled, dkey, pair0, dth, keep_word);
                    else dq_group_math<false, DROP, false>(rs, rp, ds, c1, -my_lse, km, n_vis, scale_keep, neg_delta_scaled, dkey, pair0, dth, keep_word);
                }
                if (!MERGE && j2 >= 2) mbar_wait(bar_dq + (j2 & 1), ((j2 - 2) >> 1) & 1);   // dS tile free again
                store_tile_chunk32_packed(tile, row, c, ds);
                fence_proxy_async_smem();
            }
        }
        __syncthreads();
        if (jj == n_pre - 1) {
            // end of the pre-sweep: combine the four quarter-row partials
            s_part[(quad * 128 + row) * 2] = acc_p;
            s_part[(quad * 128 + row) * 2 + 1] = acc_pd;
            __syncthreads();
            float z = 0.f, pd = 0.f;
#pragma unroll
            for (int g = 0; g < 4; ++g) z += s_part[(g * 128 + row) * 2], pd += s_part[(g * 128 + row) * 2 + 1];
            const float my_delta = z > 0.f ? pd / z : 0.f;
            neg_delta_scaled = -my_delta * p.scale;
            if (quad == 0 && q < p.lq) p.delta[((size_t)b * p.heads + h) * p.lq + q] = my_delta;
        }
        if (tid == 0 && sweep2) {
            tc_fence_after();
            if (!MERGE && j2 >= 2) mbar_wait(bar_dq + (j2 & 1), ((j2 - 2) >> 1) & 1);
            const uint32_t a = smem_u32(tile), bk = smem_u32(sK + (jj % 3) * 16384);
            for (int ks = 0; ks < (n16 >> 4); ++ks)
                umma_bf16(t_dq, tile_a_desc(a, ks), rows_b_desc(bk, ks), idesc_dq, (j2 > 0 || ks > 0) ? 1u : 0u);
            if (MERGE) {
                // dV_j = dropout(P)^T dO, dK_j = dS^T Q: A = the same tiles read MN-major (64-key atoms 16 KB apart,
                // 8 query rows per 1 KB group), k runs over the 128 query rows
                const uint32_t ap = smem_u32(sPd), bdo = smem_u32(sdO), bq = smem_u32(sQ);
#pragma unroll
                for (int ks = 0; ks < 8; ++ks)
                    umma_bf16(t_dv, make_smem_desc(ap + ks * 2048, 16384, 1024), rows_b_desc(bdo, ks), idesc_kv, ks > 0);
#pragma unroll
                for (int ks = 0; ks < 8; ++ks)
                    umma_bf16(t_dk, make_smem_desc(a + ks * 2048, 16384, 1024), rows_b_desc(bq, ks), idesc_kv, ks > 0);
            }
            tc_commit(bar_dq + (j2 & 1));
        }
        __syncwarp();
    }
    mbar_wait(bar_dq + ((n_kc - 1) & 1), ((n_kc - 1) >> 1) & 1);
    tc_fence_after();
    if (MERGE) store_dkv(n_kc - 1);
    {
        uint32_t raw[16];
        tmem_ld16(t_dq + t_row + quad * 16, raw);
        tmem_ld_wait();
        if (q < p.lq) {
            __nv_bfloat16* dst = p.dq + ((size_t)b * p.lq + q) * p.lddq + h * kHeadDim + quad * 16;
#pragma unroll
            for (int i = 0; i < 16; i += 8) {
                uint4 u;
                u.x = pack_bf16x2(__uint_as_float(raw[i]), __uint_as_float(raw[i + 1]));
                u.y = pack_bf16x2(__uint_as_float(raw[i + 2]), __uint_as_float(raw[i + 3]));
                u.z = pack_bf16x2(__uint_as_float(raw[i + 4]), __uint_as_float(raw[i + 5]));
                u.w = pack_bf16x2(__uint_as_float(raw[i + 6]), __uint_as_float(raw[i + 7]));
                *reinterpret_cast<uint4*>(dst + i) = u;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem, 512);
    }
}

template <int DROP>
__global__ void __launch_bounds__(kBwdThreads)
attn_bwd_dkv_kernel(const __grid_constant__ CUtensorMap tma_q, const __grid_constant__ CUtensorMap tma_do,
                    const __grid_constant__ CUtensorMap tma_k, const __grid_constant__ CUtensorMap tma_v,
                    const AttnBwdParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sK = smem;                       // 16 KB
    uint8_t* sV = smem + 16384;               // 16 KB
    uint8_t* sQ = smem + 32768;               // 3 x 16 KB
    uint8_t* sdO = smem + 81920;              // 3 x 16 KB
    uint8_t* sPt = smem + 131072;             // 32 KB   P^T  [128 keys x 128 queries]
    uint8_t* sdSt = smem + 163840;            // 32 KB   dS^T
    float* s_lse = reinterpret_cast<float*>(smem + kBwdTileBytes);   // [4][128] ring over query chunks
    float* s_delta = s_lse + 512;                                      // [4][128]
    uint32_t* s_bits = reinterpret_cast<uint32_t*>(s_delta + 512);     // [4][4 key words][128 queries] keep-bit words
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_bits + 2048);
    uint64_t* bar_kv = bars;
    uint64_t* bar_q = bars + 1;      // [3]
    uint64_t* bar_s = bars + 4;
    uint64_t* bar_acc = bars + 5;    // dK/dV products of chunk i retired
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 6);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int k0 = blockIdx.x * 128, h = blockIdx.y, b = blockIdx.z;
    const int n_qc = (p.lq + 127) / 128;

    if (tid == 0) {
        mbar_init(bar_kv, 1);
        mbar_init(bar_q, 1), mbar_init(bar_q + 1, 1), mbar_init(bar_q + 2, 1);
        mbar_init(bar_s, 1);
        mbar_init(bar_acc, 1);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t t_s = tmem, t_dp = tmem + 128, t_dk = tmem + 256, t_dv = tmem + 320;
    const uint32_t idesc_acc = make_idesc_bf16(128, kHeadDim, 0, 1);
    auto chunk_n16 = [&](int i) { return (min(128, p.lq - i * 128) + 15) & ~15; };   // queries of chunk i, MMA granule

    auto issue_s = [&](int i) {   // S^T = K Q_i^T ; dP^T = V dO_i^T
        const uint32_t idesc_s = make_idesc_bf16(128, chunk_n16(i), 0, 0);
        const uint32_t ak = smem_u32(sK), av = smem_u32(sV);
        const uint32_t bq = smem_u32(sQ + (i % 3) * 16384), bd = smem_u32(sdO + (i % 3) * 16384);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
            umma_bf16(t_s, make_smem_desc(ak + ks * 32, 16, 1024), make_smem_desc(bq + ks * 32, 16, 1024), idesc_s, ks > 0);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
            umma_bf16(t_dp, make_smem_desc(av + ks * 32, 16, 1024), make_smem_desc(bd + ks * 32, 16, 1024), idesc_s, ks > 0);
        tc_commit(bar_s);
    };
    auto load_q = [&](int i) {
        const int st = i % 3;
        mbar_expect_tx(bar_q + st, 32768);
        tma_load_3d(sQ + st * 16384, &tma_q, bar_q + st, h * kHeadDim, i * 128, b);
        tma_load_3d(sdO + st * 16384, &tma_do, bar_q + st, h * kHeadDim, i * 128, b);
    };
    // Per-query vectors of a chunk (lse, delta, and with stored dropout masks the four keep-bit words of this key
    // tile) go global -> register at the TOP of an iteration and register -> shared memory at its END, two chunks
    // ahead of their use: a global load feeding a shared store directly would stall the thread for an HBM/L2 round
    // trip on every chunk.
    const int nw = (p.lk + 31) >> 5;
    auto fetch_vec = [&](int i, float& v, uint32_t& w) {
        const int qq = i * 128 + (tid & 127);
        const size_t base = ((size_t)b * p.heads + h) * p.lq;
        v = 0.f, w = 0u;
        if (tid < 128) v = qq < p.lq ? __ldg(p.lse + base + qq) * kLog2e : INFINITY;
        else if (tid < 256) v = qq < p.lq ? __ldg(p.delta + base + qq) * p.scale : 0.f;       // pre-scaled
        if (DROP == 2) {
            const int ql = i * 128 + (tid >> 2), kw = (k0 >> 5) + (tid & 3);
            if (ql < p.lq && kw < nw) w = __ldg(p.drop_bits + (base + ql) * nw + kw);
        }
    };
    auto stash_vec = [&](int i, float v, uint32_t w) {
        if (tid < 128) s_lse[(i & 3) * 128 + tid] = v;
        else if (tid < 256) s_delta[(i & 3) * 128 + (tid - 128)] = v;
        if (DROP == 2) s_bits[(i & 3) * 512 + (tid & 3) * 128 + (tid >> 2)] = w;
    };
    if (tid == 0) {
        mbar_expect_tx(bar_kv, 32768);
        tma_load_3d(sK, &tma_k, bar_kv, h * kHeadDim, k0, b);
        tma_load_3d(sV, &tma_v, bar_kv, h * kHeadDim, k0, b);
        load_q(0);
        if (n_qc > 1) load_q(1);
        if (n_qc > 2) load_q(2);
        mbar_wait(bar_kv, 0);
        mbar_wait(bar_q, 0);
        tc_fence_after();
        issue_s(0);
    }
    {
        float v;
        uint32_t w;
        fetch_vec(0, v, w), stash_vec(0, v, w);
        if (n_qc > 1) fetch_vec(1, v, w), stash_vec(1, v, w);
    }
    __syncthreads();

    const int row = (warp & 3) * 32 + lane, quad = warp >> 2;
    const int key = k0 + row;
    const uint32_t t_row = uint32_t((warp & 3) * 32) << 16;
    const float c1 = p.scale * kLog2e;
    const float kbias = (p.mask_mode == 1 && key < p.lk) ? p.key_mask[(size_t)b * p.lk + key] * kLog2e : 0.f;
    // block-causal: this key is visible to queries of frame >= its own frame, i.e. q >= q_first
    const int q_first = p.mask_mode == 2 ? (key / p.k_per_frame) * p.q_per_frame : 0;
    const bool key_ok = key < p.lk;
    uint32_t dkey = 0, dth = 0, lane_mask = 0, pair_stride = 0, col_pair0 = 0;
    float inv_keep = 1.f;
    if (DROP) {
        if (DROP == 1) dkey = drop_key(__ldg(p.drop.seed), p.drop.salt), dth = drop_thresh(p.drop.p);
        inv_keep = 1.f / (1.f - p.drop.p);
        lane_mask = (key & 1) ? 0xffff0000u : 0x0000ffffu;
        pair_stride = attn_row_pair0(0, p.heads, 0, p.lq, 1, p.lk);                 // pairs per query row
        col_pair0 = attn_row_pair0(b, p.heads, h, p.lq, 0, p.lk) + (uint32_t)(key >> 1);
    }

    const bool rows_live = k0 + (warp & 3) * 32 < p.lk;     // warps of key rows beyond lk: nothing to do (dK/dV rows unused)
    const float scale_keep = p.scale * inv_keep;
    for (int i = 0; i < n_qc; ++i) {
        float nxt_v = 0.f;
        uint32_t nxt_w = 0u;
        if (i + 2 < n_qc) fetch_vec(i + 2, nxt_v, nxt_w);
        const int n16 = chunk_n16(i);
        const int c = quad * 32;
        const bool live = rows_live && c < n16;     // warp-uniform
        mbar_wait(bar_s, i & 1);
        tc_fence_after();
        uint32_t rs[32], rp[32];
        if (live) {
            tmem_ld32(t_s + t_row + c, rs);
            tmem_ld32(t_dp + t_row + c, rp);
            tmem_ld_wait();
        }
        tc_fence_before();
        __syncthreads();                            // S^T/dP^T are in registers: their TMEM tiles are free
        if (tid == 0) {
            tc_fence_after();
            if (i + 1 < n_qc) {                     // next chunk's score products run under this chunk's exponentials
                mbar_wait(bar_q + (i + 1) % 3, ((i + 1) / 3) & 1);
                tc_fence_after();
                issue_s(i + 1);
            }
            if (i >= 1 && i + 2 < n_qc) {
                mbar_wait(bar_acc, (i - 1) & 1);    // dK/dV of chunk i-1 retired: its Q/dO buffer is free
                load_q(i + 2);
            }
        }
        if (live) {
            const float* lse_i = s_lse + (i & 3) * 128 + c;
            const float* del_i = s_delta + (i & 3) * 128 + c;
            const uint32_t* bits_i = s_bits + (i & 3) * 512 + (row >> 5) * 128 + c;
            uint32_t pr[16], ds[16];
            // first query of this group that sees the key (block-causal); queries beyond lq carry lse = +inf -> p = 0
            const int e_first = key_ok ? q_first - (i * 128 + c) : 32;
            const uint32_t pair0 = col_pair0 + (uint32_t)(i * 128 + c) * pair_stride;
            if (__all_sync(0xffffffffu, e_first <= 0))
                dkv_group_math<DROP, true>(rs, rp, pr, ds, c1, kbias, lse_i, del_i, e_first, scale_keep, inv_keep, dkey, pair0, pair_stride, dth, lane_mask, bits_i, row & 31);
            else
                dkv_group_math<DROP, false>(rs, rp, pr, ds, c1, kbias, lse_i, del_i, e_first, scale_keep, inv_keep, dkey, pair0, pair_stride, dth, lane_mask, bits_i, row & 31);
            if (i >= 1) mbar_wait(bar_acc, (i - 1) & 1);   // sPt / sdSt free again
            store_tile_chunk32_packed(sPt, row, c, pr);
            store_tile_chunk32_packed(sdSt, row, c, ds);
            fence_proxy_async_smem();
        }
        if (i + 2 < n_qc) stash_vec(i + 2, nxt_v, nxt_w);     // slot (i + 2) & 3 was last read two chunks ago
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
            if (i >= 1) mbar_wait(bar_acc, (i - 1) & 1);
            const uint32_t ap = smem_u32(sPt), ads = smem_u32(sdSt);
            const uint32_t bq = smem_u32(sQ + (i % 3) * 16384), bd = smem_u32(sdO + (i % 3) * 16384);
            for (int ks = 0; ks < (n16 >> 4); ++ks)
                umma_bf16(t_dv, tile_a_desc(ap, ks), rows_b_desc(bd, ks), idesc_acc, (i > 0 || ks > 0) ? 1u : 0u);
            for (int ks = 0; ks < (n16 >> 4); ++ks)
                umma_bf16(t_dk, tile_a_desc(ads, ks), rows_b_desc(bq, ks), idesc_acc, (i > 0 || ks > 0) ? 1u : 0u);
            tc_commit(bar_acc);
        }
        __syncwarp();
    }
    mbar_wait(bar_acc, (n_qc - 1) & 1);
    tc_fence_after();
    {
        // warp quads 0,1 write the two halves of dK, quads 2,3 those of dV (32 columns each)
        const uint32_t t_src = (quad < 2 ? t_dk : t_dv) + t_row;
        __nv_bfloat16* base = quad < 2 ? p.dk : p.dv;
        const int64_t ld = quad < 2 ? p.lddk : p.lddv;
        {
            const int c = (quad & 1) * 32;
            uint32_t raw[32];
            tmem_ld32(t_src + c, raw);
            tmem_ld_wait();
            if (key_ok) {
                __nv_bfloat16* dst = base + ((size_t)b * p.lk + key) * ld + h * kHeadDim + c;
#pragma unroll
                for (int e = 0; e < 32; e += 8) {
                    uint4 u;
                    u.x = pack_bf16x2(__uint_as_float(raw[e]), __uint_as_float(raw[e + 1]));
                    u.y = pack_bf16x2(__uint_as_float(raw[e + 2]), __uint_as_float(raw[e + 3]));
                    u.z = pack_bf16x2(__uint_as_float(raw[e + 4]), __uint_as_float(raw[e + 5]));
                    u.w = pack_bf16x2(__uint_as_float(raw[e + 6]), __uint_as_float(raw[e + 7]));
                    *reinterpret_cast<uint4*>(dst + e) = u;
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem, 512);
    }
}

// =====================================================================================================
// Backward, second generation (problems with more than one query tile, fp32 O saved by the forward): the same two
// passes, re-cut so that TWO CTAs share an SM, each CTA runs TWO independent half-chunk pipelines, and the bf16
// P / dS operands never leave tensor memory.
//   * The first generation (512 threads, 128-wide chunks, 197 KB, 512 TMEM columns, one CTA per SM) spent 35 % of its
//     issue slots at barriers: every chunk is a chain MMA -> tcgen05.ld -> exponentials -> operand tile -> MMA with two
//     block-wide barriers, the MMA / TMA issue sits on one math thread's path, and one CTA has nothing else to run.
//   * Here a chunk is 64 keys (dq pass) / 64 queries (dkv pass), cut into two 32-wide halves.  Warps 0-3 own half 0,
//     warps 4-7 half 1 (thread = tile row x 32 columns), warp 8 issues every TMA load and tcgen05.mma.  A half's score
//     products, its exponentials and its accumulating products form their own chain, linked to the issuer by two
//     mbarriers (values loaded / operand stored) — no block-wide barrier inside the loop; while one half does its
//     exponentials the tensor core runs the other half's (and the other CTA's) products.  (In the dq pass the two halves
//     share ONE set of score products per chunk — an instruction costs the issuing thread the same whatever its width —
//     and keep their own accumulating products.)
//   * the dS (dq pass) and P^T / dS^T (dkv pass) operands go registers -> tcgen05.st -> TMEM and feed the accumulating
//     products as TMEM A operands (tcgen05.mma [d], [a], b-desc): no swizzled shared-memory tile, no proxy fence.  In the
//     dkv pass they overwrite the first half of the very score columns their warp has just read (256 TMEM columns per
//     CTA: S^T 64 | dP^T 64 | dK 64 | dV 64); the next score products into those columns are issued behind the products
//     that read them, on the same in-order tensor pipe.
//   * 393 keys are 7 chunks of 64 (400 of 448 columns useful, the last chunk 16 wide) instead of 4 of 128 (400 of 512).
//   84 KB (dq) / 92 KB (dkv) of shared memory and 256 TMEM columns per CTA, 288 threads, <= 96 registers (a 288-thread CTA is allocated as 320 threads).
// =====================================================================================================
constexpr int kB2Threads = 288;
constexpr int kB2Chunk = 64;
constexpr int kB2TileBytes = 81920;     // 2 x 16 KB resident tiles + 3 stages x 2 x 8 KB streamed chunks
constexpr int kB2MaxLen = 512;          // longest lq / lk these kernels take (per-row vectors staged in shared memory)

struct B2Bars {
    uint64_t res;          // the CTA's resident tiles landed
    uint64_t chunk[3];     // streamed chunk stage landed
    uint64_t s[2];         // half h: score products of the current chunk complete
    uint64_t ld[2];        // half h: its four warps hold the score values in registers
    uint64_t st[2];        // half h: its four warps have stored the bf16 operand
    uint64_t acc[2];       // dq pass, half h: the dQ product of a chunk retired
    uint64_t done;         // every product retired
    uint32_t tmem_slot;
};

template <int DROP>
__global__ void __launch_bounds__(kB2Threads, 2)
attn_bwd2_dq_kernel(const __grid_constant__ CUtensorMap tma_q, const __grid_constant__ CUtensorMap tma_do,
                    const __grid_constant__ CUtensorMap tma_k, const __grid_constant__ CUtensorMap tma_v,
                    const AttnBwdParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sQ = smem;                       // 16 KB   [128 queries x 64]
    uint8_t* sdO = smem + 16384;              // 16 KB
    uint8_t* sK = smem + 32768;               // 3 x 8 KB  [64 keys x 64]
    uint8_t* sV = smem + 57344;               // 3 x 8 KB
    float* s_part = reinterpret_cast<float*>(smem + kB2TileBytes);   // [2 halves][128 rows]
    float* s_kmask = s_part + 256;                                     // [512] key mask x log2(e)
    B2Bars* bars = reinterpret_cast<B2Bars*>(s_kmask + kB2MaxLen);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int q0 = blockIdx.x * 128, h = blockIdx.y, b = blockIdx.z;
    const int n_kc = (p.lk + kB2Chunk - 1) / kB2Chunk;

    auto load_kv = [&](int j) {
        const int st = j % 3;
        mbar_expect_tx(&bars->chunk[st], 16384);
        tma_load_3d(sK + st * 8192, &tma_k, &bars->chunk[st], h * kHeadDim, j * kB2Chunk, b);
        tma_load_3d(sV + st * 8192, &tma_v, &bars->chunk[st], h * kHeadDim, j * kB2Chunk, b);
    };
    if (tid == 256) {
        // the issuing thread initialises the barriers and starts the first loads at once: they overlap the TMEM
        // allocation and the other warps' first global reads
        mbar_init(&bars->res, 1);
        for (int i = 0; i < 3; ++i) mbar_init(&bars->chunk[i], 1);
        for (int i = 0; i < 2; ++i)
            mbar_init(&bars->s[i], 1), mbar_init(&bars->ld[i], 4), mbar_init(&bars->st[i], 4), mbar_init(&bars->acc[i], 1);
        mbar_init(&bars->done, 1);
        fence_barrier_init();
        mbar_expect_tx(&bars->res, 32768);
        tma_load_3d(sQ, &tma_q, &bars->res, h * kHeadDim, q0, b);
        tma_load_3d(sdO, &tma_do, &bars->res, h * kHeadDim, q0, b);
        load_kv(0);
        if (n_kc > 1) load_kv(1);
        if (n_kc > 2) load_kv(2);
    }
    // delta = rowsum(dO o O_f32) (see attn_bwd_dq_kernel): each of the row's two threads takes 32 head columns.  The
    // global reads are issued before anything else so that their latency overlaps the set-up and the first products.
    float delta_part = 0.f, my_lse = INFINITY;
    if (warp < 8) {
        const int row_ = (warp & 3) * 32 + lane, half_ = warp >> 2, q_ = q0 + row_;
        if (q_ < p.lq) {
            my_lse = __ldg(p.lse + ((size_t)b * p.heads + h) * p.lq + q_) * kLog2e;
            const float4* po = reinterpret_cast<const float4*>(p.o_f32 + ((size_t)b * p.lq + q_) * p.ldof + h * kHeadDim + half_ * 32);
            const uint4* pd = reinterpret_cast<const uint4*>(p.d_o + ((size_t)b * p.lq + q_) * p.lddo + h * kHeadDim + half_ * 32);
            uint4 g[4];
            float4 o[8];
#pragma unroll
            for (int i = 0; i < 4; ++i) g[i] = __ldg(pd + i);
#pragma unroll
            for (int i = 0; i < 8; ++i) o[i] = __ldg(po + i);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float4 o0 = o[2 * i], o1 = o[2 * i + 1];
                float2 y;
                y = unpack_bf16x2(g[i].x), delta_part = fmaf(o0.x, y.x, fmaf(o0.y, y.y, delta_part));
                y = unpack_bf16x2(g[i].y), delta_part = fmaf(o0.z, y.x, fmaf(o0.w, y.y, delta_part));
                y = unpack_bf16x2(g[i].z), delta_part = fmaf(o1.x, y.x, fmaf(o1.y, y.y, delta_part));
                y = unpack_bf16x2(g[i].w), delta_part = fmaf(o1.z, y.x, fmaf(o1.w, y.y, delta_part));
            }
        }
    }
    if (warp == 0) tmem_alloc(&bars->tmem_slot, 256);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = bars->tmem_slot;
    const uint32_t t_s = tmem, t_dp = tmem + 64, t_dq = tmem + 128, t_ds = tmem + 192;   // dS: bf16 pairs, 2 x 16 columns
    // keys of chunk j that half hh covers, rounded up to the MMA granule (0: the half has nothing to do)
    auto half_n16 = [&](int j, int hh) { return max(0, min(32, ((min(kB2Chunk, p.lk - j * kB2Chunk) + 15) & ~15) - 32 * hh)); };

    if (warp == 8) {
        // ------------------------------------------------------------------ issuer: TMA loads and every product
        if (lane == 0) {
            const uint32_t idesc_dq = make_idesc_bf16(128, kHeadDim, 0, 1);
            // S = Q K_j^T ; dP = dO V_j^T for the whole 64-key chunk in ONE set of eight products (a tcgen05.mma costs the
            // issuing thread ~47 clk whatever its width — tools/micro/mma_issue.cu — so the halves share the score
            // products and keep only their own accumulating products); one commit per half that has keys
            auto issue_s = [&](int j) {
                const int n = half_n16(j, 0) + half_n16(j, 1);
                const uint32_t aq = smem_u32(sQ), ado = smem_u32(sdO);
                const uint32_t bk = smem_u32(sK + (j % 3) * 8192), bv = smem_u32(sV + (j % 3) * 8192);
                const uint32_t idesc_s = make_idesc_bf16(128, n, 0, 0);
#pragma unroll
                for (int ks = 0; ks < 4; ++ks)
                    umma_bf16(t_s, make_smem_desc(aq + ks * 32, 16, 1024), make_smem_desc(bk + ks * 32, 16, 1024), idesc_s, ks > 0);
#pragma unroll
                for (int ks = 0; ks < 4; ++ks)
                    umma_bf16(t_dp, make_smem_desc(ado + ks * 32, 16, 1024), make_smem_desc(bv + ks * 32, 16, 1024), idesc_s, ks > 0);
                tc_commit(&bars->s[0]);
                if (half_n16(j, 1) > 0) tc_commit(&bars->s[1]);
            };
            mbar_wait(&bars->res, 0);
            mbar_wait(&bars->chunk[0], 0);
            tc_fence_after();
            issue_s(0);
            bool first = true;
            for (int j = 0; j < n_kc; ++j) {
                if (j + 1 < n_kc) {
                    // chunk j is full here (only the last chunk can be short): both halves read their S/dP columns
                    mbar_wait(&bars->chunk[(j + 1) % 3], ((j + 1) / 3) & 1);
                    mbar_wait(&bars->ld[0], j & 1);
                    mbar_wait(&bars->ld[1], j & 1);
                    tc_fence_after();
                    issue_s(j + 1);
                }
                if (j >= 1 && j + 2 < n_kc) {
                    // K/V buffer (j - 1) % 3 is free once both halves' dQ products of chunk j - 1 retired
                    mbar_wait(&bars->acc[0], (j - 1) & 1);
                    mbar_wait(&bars->acc[1], (j - 1) & 1);
                    load_kv(j + 2);
                }
                const uint32_t bk = smem_u32(sK + (j % 3) * 8192);
                for (int hh = 0; hh < 2; ++hh) {
                    const int n = half_n16(j, hh);
                    if (n == 0) continue;
                    mbar_wait(&bars->st[hh], j & 1);            // the half's dS operand is in tensor memory
                    tc_fence_after();
                    for (int ks = 0; ks < (n >> 4); ++ks) {
                        umma_bf16_ts(t_dq, t_ds + hh * 16 + ks * 8, rows_b_desc(bk, 2 * hh + ks), idesc_dq, first ? 0u : 1u);
                        first = false;
                    }
                    tc_commit(&bars->acc[hh]);
                }
            }
            tc_commit(&bars->done);
        }
        __syncwarp();
    } else {
        // ------------------------------------------------------------------ math warps
        const int row = (warp & 3) * 32 + lane, half = warp >> 2;     // half: which 32 keys of the 64-key chunk
        const int q = q0 + row;
        const uint32_t t_row = uint32_t((warp & 3) * 32) << 16;
        const float* kmask = p.mask_mode == 1 ? p.key_mask + (size_t)b * p.lk : nullptr;
        const int k_lim = p.mask_mode == 2 ? min(p.lk, (q / p.q_per_frame + 1) * p.k_per_frame) : p.lk;
        const float c1 = p.scale * kLog2e;
        uint32_t dkey = 0, dth = 0, row_pair0 = 0;
        float inv_keep = 1.f;
        if (DROP) {
            dkey = drop_key(__ldg(p.drop.seed), p.drop.salt), dth = drop_thresh(p.drop.p);
            row_pair0 = attn_row_pair0(b, p.heads, h, p.lq, q, p.lk), inv_keep = 1.f / (1.f - p.drop.p);
        }
        if (kmask)
            for (int k = tid; k < kB2MaxLen; k += 256) s_kmask[k] = k < p.lk ? kmask[k] * kLog2e : 0.f;
        s_part[half * 128 + row] = delta_part;
        asm volatile("bar.sync 1, 256;" ::: "memory");      // the eight math warps (the issuer does not take part)
        const float my_delta = s_part[row] + s_part[128 + row];
        const float neg_delta_scaled = -my_delta * p.scale;
        if (half == 0 && q < p.lq) p.delta[((size_t)b * p.heads + h) * p.lq + q] = my_delta;
        const bool rows_live = q0 + (warp & 3) * 32 < p.lq;
        const float scale_keep = p.scale * inv_keep;
        const int c = half * 32;
        for (int j = 0; j < n_kc; ++j) {
            const int n16 = half_n16(j, half);
            if (n16 == 0) break;                        // only the last chunk can be short
            uint32_t keep_word = 0;
            if (DROP == 2 && rows_live && q < p.lq)     // requested ahead of the wait for the score products
                keep_word = __ldg(p.drop_bits + (((size_t)b * p.heads + h) * p.lq + q) * (size_t)((p.lk + 31) >> 5) + j * 2 + half);
            mbar_wait(&bars->s[half], j & 1);
            tc_fence_after();
            uint32_t ds0[8], ds1[8];
            const int key0 = j * kB2Chunk + c;
            // two 16-column batches keep the live registers under the 96-register cap
            auto batch = [&](const int sb, uint32_t(&dsb)[8]) {
                uint32_t rs[16], rp[16];
                if (rows_live) {
                    tmem_ld16(t_s + t_row + c + 16 * sb, rs);
                    tmem_ld16(t_dp + t_row + c + 16 * sb, rp);
                    tmem_ld_wait();
                }
                if (sb == 1) {
                    // the half's score columns are in registers: the issuer may overwrite them with the next chunk's
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_relaxed(&bars->ld[half]);
                }
                if (rows_live) {
                    const int n_vis = k_lim - key0 - 16 * sb;      // visible columns of this batch (may be <= 0 or >= 16)
                    const uint32_t pair0 = row_pair0 + (uint32_t)((key0 + 16 * sb) >> 1);
                    const float* km = s_kmask + key0 + 16 * sb;
                    const uint32_t kw = keep_word >> (16 * sb);
                    if (__all_sync(0xffffffffu, n_vis >= 16)) {
                        if (kmask) dq_group_math<true, DROP, true, 16>(rs, rp, dsb, c1, -my_lse, km, n_vis, scale_keep, neg_delta_scaled, dkey, pair0, dth, kw);
                        else dq_group_math<false, DROP, true, 16>(rs, rp, dsb, c1, -my_lse, km, n_vis, scale_keep, neg_delta_scaled, dkey, pair0, dth, kw);
                    } else {
                        if (kmask) dq_group_math<true, DROP, false, 16>(rs, rp, dsb, c1, -my_lse, km, n_vis, scale_keep, neg_delta_scaled, dkey, pair0, dth, kw);
                        else dq_group_math<false, DROP, false, 16>(rs, rp, dsb, c1, -my_lse, km, n_vis, scale_keep, neg_delta_scaled, dkey, pair0, dth, kw);
                    }
                }
            };
            batch(0, ds0);
            batch(1, ds1);
            if (j >= 1) {
                mbar_wait(&bars->acc[half], (j - 1) & 1);   // the half's previous dQ product has read its dS columns
                tc_fence_after();
            }
            if (rows_live) {
                tmem_st8(t_ds + t_row + half * 16, ds0);
                tmem_st8(t_ds + t_row + half * 16 + 8, ds1);
                tmem_st_wait();
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_relaxed(&bars->st[half]);
        }
        mbar_wait(&bars->done, 0);
        tc_fence_after();
        {
            uint32_t raw[32];
            tmem_ld32(t_dq + t_row + half * 32, raw);
            tmem_ld_wait();
            if (q < p.lq) {
                __nv_bfloat16* dst = p.dq + ((size_t)b * p.lq + q) * p.lddq + h * kHeadDim + half * 32;
#pragma unroll
                for (int i = 0; i < 32; i += 8) {
                    uint4 u;
                    u.x = pack_bf16x2(__uint_as_float(raw[i]), __uint_as_float(raw[i + 1]));
                    u.y = pack_bf16x2(__uint_as_float(raw[i + 2]), __uint_as_float(raw[i + 3]));
                    u.z = pack_bf16x2(__uint_as_float(raw[i + 4]), __uint_as_float(raw[i + 5]));
                    u.w = pack_bf16x2(__uint_as_float(raw[i + 6]), __uint_as_float(raw[i + 7]));
                    *reinterpret_cast<uint4*>(dst + i) = u;
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem, 256);
    }
}

// dkv pass: CTA = (128-key tile, head, batch), 64-query chunks, half = 32 queries.
template <int DROP>
__global__ void __launch_bounds__(kB2Threads, 2)
attn_bwd2_dkv_kernel(const __grid_constant__ CUtensorMap tma_q, const __grid_constant__ CUtensorMap tma_do,
                     const __grid_constant__ CUtensorMap tma_k, const __grid_constant__ CUtensorMap tma_v,
                     const AttnBwdParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sK = smem;                       // 16 KB   [128 keys x 64]
    uint8_t* sV = smem + 16384;               // 16 KB
    uint8_t* sQ = smem + 32768;               // 3 x 8 KB  [64 queries x 64]
    uint8_t* sdO = smem + 57344;              // 3 x 8 KB
    float* s_lse = reinterpret_cast<float*>(smem + kB2TileBytes);      // [512] lse x log2(e) of every query (+inf past lq)
    float* s_delta = s_lse + kB2MaxLen;                                  // [512] delta x scale
    uint32_t* s_bits = reinterpret_cast<uint32_t*>(s_delta + kB2MaxLen); // [4 key words][512 queries] keep-bit words
    B2Bars* bars = reinterpret_cast<B2Bars*>(s_bits + 4 * kB2MaxLen);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int k0 = blockIdx.x * 128, h = blockIdx.y, b = blockIdx.z;
    const int n_qc = (p.lq + kB2Chunk - 1) / kB2Chunk;

    auto load_q = [&](int i) {
        const int st = i % 3;
        mbar_expect_tx(&bars->chunk[st], 16384);
        tma_load_3d(sQ + st * 8192, &tma_q, &bars->chunk[st], h * kHeadDim, i * kB2Chunk, b);
        tma_load_3d(sdO + st * 8192, &tma_do, &bars->chunk[st], h * kHeadDim, i * kB2Chunk, b);
    };
    if (tid == 256) {
        // barriers and the first loads by the issuing thread, under the staging of the per-query vectors below
        mbar_init(&bars->res, 1);
        for (int i = 0; i < 3; ++i) mbar_init(&bars->chunk[i], 1);
        for (int i = 0; i < 2; ++i)
            mbar_init(&bars->s[i], 1), mbar_init(&bars->ld[i], 4), mbar_init(&bars->st[i], 4), mbar_init(&bars->acc[i], 1);
        mbar_init(&bars->done, 1);
        fence_barrier_init();
        mbar_expect_tx(&bars->res, 32768);
        tma_load_3d(sK, &tma_k, &bars->res, h * kHeadDim, k0, b);
        tma_load_3d(sV, &tma_v, &bars->res, h * kHeadDim, k0, b);
        load_q(0);
        if (n_qc > 1) load_q(1);
        if (n_qc > 2) load_q(2);
    }
    if (warp == 0) tmem_alloc(&bars->tmem_slot, 256);
    {
        // per-query vectors of this (batch, head), staged once
        const size_t base = ((size_t)b * p.heads + h) * p.lq;
        const int nw = (p.lk + 31) >> 5;
        for (int i = tid; i < kB2MaxLen; i += kB2Threads) {
            s_lse[i] = i < p.lq ? __ldg(p.lse + base + i) * kLog2e : INFINITY;
            s_delta[i] = i < p.lq ? __ldg(p.delta + base + i) * p.scale : 0.f;
        }
        if (DROP == 2)
            for (int i = tid; i < 4 * kB2MaxLen; i += kB2Threads) {
                const int ql = i >> 2, kw = (k0 >> 5) + (i & 3);
                s_bits[(i & 3) * kB2MaxLen + ql] = (ql < p.lq && kw < nw) ? __ldg(p.drop_bits + (base + ql) * nw + kw) : 0u;
            }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = bars->tmem_slot;
    const uint32_t t_s = tmem, t_dp = tmem + 64, t_dk = tmem + 128, t_dv = tmem + 192;
    auto half_n16 = [&](int i, int hh) { return max(0, min(32, ((min(kB2Chunk, p.lq - i * kB2Chunk) + 15) & ~15) - 32 * hh)); };

    if (warp == 8) {
        if (lane == 0) {
            const uint32_t idesc_acc = make_idesc_bf16(128, kHeadDim, 0, 1);
            auto issue_s = [&](int i, int hh) {   // S^T = K Q_i^T ; dP^T = V dO_i^T  for the 32 queries of half hh
                const int n = half_n16(i, hh);
                if (n == 0) return;
                const uint32_t idesc_s = make_idesc_bf16(128, n, 0, 0);
                const uint32_t ak = smem_u32(sK), av = smem_u32(sV);
                const uint32_t bq = smem_u32(sQ + (i % 3) * 8192) + hh * 4096, bd = smem_u32(sdO + (i % 3) * 8192) + hh * 4096;
#pragma unroll
                for (int ks = 0; ks < 4; ++ks)
                    umma_bf16(t_s + hh * 32, make_smem_desc(ak + ks * 32, 16, 1024), make_smem_desc(bq + ks * 32, 16, 1024), idesc_s, ks > 0);
#pragma unroll
                for (int ks = 0; ks < 4; ++ks)
                    umma_bf16(t_dp + hh * 32, make_smem_desc(av + ks * 32, 16, 1024), make_smem_desc(bd + ks * 32, 16, 1024), idesc_s, ks > 0);
                tc_commit(&bars->s[hh]);           // also covers every product issued before (dK / dV of the chunk before)
            };
            mbar_wait(&bars->res, 0);
            mbar_wait(&bars->chunk[0], 0);
            tc_fence_after();
            issue_s(0, 0), issue_s(0, 1);
            bool first = true;
            for (int i = 0; i < n_qc; ++i) {
                if (i >= 1 && i + 2 < n_qc) {
                    // Q/dO buffer (i - 1) % 3 was last read by the dK / dV products of chunk i - 1, which the commits
                    // of chunk i's score products cover (full chunks only: i + 2 < n_qc)
                    mbar_wait(&bars->s[0], i & 1);
                    mbar_wait(&bars->s[1], i & 1);
                    load_q(i + 2);
                }
                if (i + 1 < n_qc) mbar_wait(&bars->chunk[(i + 1) % 3], ((i + 1) / 3) & 1);
                const uint32_t bq = smem_u32(sQ + (i % 3) * 8192), bd = smem_u32(sdO + (i % 3) * 8192);
                for (int hh = 0; hh < 2; ++hh) {
                    const int n = half_n16(i, hh);
                    if (n == 0) continue;
                    mbar_wait(&bars->st[hh], i & 1);            // P^T / dS^T of the half are in tensor memory
                    tc_fence_after();
                    for (int ks = 0; ks < (n >> 4); ++ks)
                        umma_bf16_ts(t_dv, t_s + hh * 32 + ks * 8, rows_b_desc(bd, 2 * hh + ks), idesc_acc, (first && ks == 0) ? 0u : 1u);
                    for (int ks = 0; ks < (n >> 4); ++ks)
                        umma_bf16_ts(t_dk, t_dp + hh * 32 + ks * 8, rows_b_desc(bq, 2 * hh + ks), idesc_acc, (first && ks == 0) ? 0u : 1u);
                    first = false;
                    if (i + 1 < n_qc) issue_s(i + 1, hh);       // behind the products that read the same columns
                }
            }
            tc_commit(&bars->done);
        }
        __syncwarp();
    } else {
        const int row = (warp & 3) * 32 + lane, half = warp >> 2;
        const int key = k0 + row;
        const uint32_t t_row = uint32_t((warp & 3) * 32) << 16;
        const float c1 = p.scale * kLog2e;
        const float kbias = (p.mask_mode == 1 && key < p.lk) ? p.key_mask[(size_t)b * p.lk + key] * kLog2e : 0.f;
        const int q_first = p.mask_mode == 2 ? (key / p.k_per_frame) * p.q_per_frame : 0;   // block-causal: visible to q >= q_first
        const bool key_ok = key < p.lk;
        uint32_t dkey = 0, dth = 0, lane_mask = 0, pair_stride = 0, col_pair0 = 0;
        float inv_keep = 1.f;
        if (DROP) {
            if (DROP == 1) dkey = drop_key(__ldg(p.drop.seed), p.drop.salt), dth = drop_thresh(p.drop.p);
            inv_keep = 1.f / (1.f - p.drop.p);
            lane_mask = (key & 1) ? 0xffff0000u : 0x0000ffffu;
            pair_stride = attn_row_pair0(0, p.heads, 0, p.lq, 1, p.lk);
            col_pair0 = attn_row_pair0(b, p.heads, h, p.lq, 0, p.lk) + (uint32_t)(key >> 1);
        }
        const bool rows_live = k0 + (warp & 3) * 32 < p.lk;     // warps of key rows beyond lk: nothing to do
        const float scale_keep = p.scale * inv_keep;
        const int c = half * 32;
        for (int i = 0; i < n_qc; ++i) {
            const int n16 = half_n16(i, half);
            if (n16 == 0) break;
            mbar_wait(&bars->s[half], i & 1);
            tc_fence_after();
            uint32_t pr0[8], pr1[8], ds0[8], ds1[8];
            const int qc0 = i * kB2Chunk + c;           // first query of this half
            auto batch = [&](const int sb, uint32_t(&prb)[8], uint32_t(&dsb)[8]) {
                uint32_t rs[16], rp[16];
                tmem_ld16(t_s + t_row + c + 16 * sb, rs);
                tmem_ld16(t_dp + t_row + c + 16 * sb, rp);
                tmem_ld_wait();
                const int qb = qc0 + 16 * sb;
                // first query of this batch that sees the key; queries beyond lq carry lse = +inf -> p = 0
                const int e_first = key_ok ? q_first - qb : 16;
                const uint32_t pair0 = col_pair0 + (uint32_t)qb * pair_stride;
                const uint32_t* bits_i = s_bits + (row >> 5) * kB2MaxLen + qb;
                if (__all_sync(0xffffffffu, e_first <= 0))
                    dkv_group_math<DROP, true, 16>(rs, rp, prb, dsb, c1, kbias, s_lse + qb, s_delta + qb, e_first, scale_keep, inv_keep, dkey, pair0, pair_stride, dth, lane_mask, bits_i, row & 31);
                else
                    dkv_group_math<DROP, false, 16>(rs, rp, prb, dsb, c1, kbias, s_lse + qb, s_delta + qb, e_first, scale_keep, inv_keep, dkey, pair0, pair_stride, dth, lane_mask, bits_i, row & 31);
            };
            if (rows_live) {
                batch(0, pr0, ds0);
                batch(1, pr1, ds1);
            }
            if (rows_live) {
                // over the first half of the score columns this warp has just read (its lanes, its 32 columns)
                tmem_st8(t_s + t_row + c, pr0);
                tmem_st8(t_s + t_row + c + 8, pr1);
                tmem_st8(t_dp + t_row + c, ds0);
                tmem_st8(t_dp + t_row + c + 8, ds1);
                tmem_st_wait();
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_relaxed(&bars->st[half]);
        }
        mbar_wait(&bars->done, 0);
        tc_fence_after();
        {
            // half 0 writes dK, half 1 dV (64 columns each, two loads of 32)
            const uint32_t t_src = (half == 0 ? t_dk : t_dv) + t_row;
            __nv_bfloat16* base = half == 0 ? p.dk : p.dv;
            const int64_t ld = half == 0 ? p.lddk : p.lddv;
#pragma unroll
            for (int cc = 0; cc < 64; cc += 32) {
                uint32_t raw[32];
                tmem_ld32(t_src + cc, raw);
                tmem_ld_wait();
                if (key_ok) {
                    __nv_bfloat16* dst = base + ((size_t)b * p.lk + key) * ld + h * kHeadDim + cc;
#pragma unroll
                    for (int e = 0; e < 32; e += 8) {
                        uint4 u;
                        u.x = pack_bf16x2(__uint_as_float(raw[e]), __uint_as_float(raw[e + 1]));
                        u.y = pack_bf16x2(__uint_as_float(raw[e + 2]), __uint_as_float(raw[e + 3]));
                        u.z = pack_bf16x2(__uint_as_float(raw[e + 4]), __uint_as_float(raw[e + 5]));
                        u.w = pack_bf16x2(__uint_as_float(raw[e + 6]), __uint_as_float(raw[e + 7]));
                        *reinterpret_cast<uint4*>(dst + e) = u;
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem, 256);
    }
}

static const CUtensorMap* head_tmap(const void* ptr, int64_t ld, int heads, int len, int batch, int box_rows) {
    uint64_t dims[3] = {(uint64_t)heads * kHeadDim, (uint64_t)len, (uint64_t)batch};
    uint64_t strides[2] = {(uint64_t)ld * 2, (uint64_t)ld * 2 * (uint64_t)len};
    uint32_t box[3] = {kHeadDim, (uint32_t)box_rows, 1};
    return get_tmap_bf16(ptr, 3, dims, strides, box);
}

}  // namespace shg

static int g_fwd2_min_lk = -1;    // key length from which the chunked two-CTA-per-SM forward is used (default 257)
extern "C" int shg_set_attn_fwd_chunked_min_lk(int min_lk) {
    g_fwd2_min_lk = min_lk < 1 ? 1 : min_lk;
    return SHG_OK;
}
static int g_attn_merge = -1;     // single-query-tile backward in one merged pass (default on; SHG_ATTN_MERGE=0: two passes)
extern "C" int shg_set_attn_bwd_merge(int on) {
    g_attn_merge = on ? 1 : 0;
    return SHG_OK;
}
static int g_attn_bwd2 = -1;      // multi-query-tile backward on the two-CTA-per-SM kernels (default on; SHG_ATTN_BWD2=0: first generation)
extern "C" int shg_set_attn_bwd_chunked(int on) {
    g_attn_bwd2 = on ? 1 : 0;
    return SHG_OK;
}
extern "C" int shg_attn_fwd(const shg_attn_args* a, void* stream) {
    using namespace shg;
    SHG_REQUIRE(a && a->q && a->k && a->v && a->o, SHG_EINVAL, "shg_attn_fwd: null operand");
    SHG_REQUIRE(a->lk >= 1 && a->lk <= 416 && a->lq >= 1, SHG_ESHAPE, "shg_attn_fwd: lk=%d outside [1,416]", a->lk);
    SHG_REQUIRE(a->mask_mode >= 0 && a->mask_mode <= 2, SHG_EINVAL, "shg_attn_fwd: mask_mode=%d", a->mask_mode);
    SHG_REQUIRE(a->mask_mode != 1 || a->key_mask, SHG_EINVAL, "shg_attn_fwd: key mask missing");
    SHG_REQUIRE(a->mask_mode != 2 || (a->q_per_frame > 0 && a->k_per_frame > 0), SHG_EINVAL,
                "shg_attn_fwd: block-causal needs q_per_frame/k_per_frame");
    SHG_REQUIRE(a->ldo % 8 == 0 && (reinterpret_cast<uintptr_t>(a->o) & 15) == 0, SHG_ESHAPE,
                "shg_attn_fwd: output must be 16-byte aligned with ld %% 8 == 0");
    AttnParams p;
    p.o = (__nv_bfloat16*)a->o, p.ldo = a->ldo, p.lse = a->lse, p.key_mask = a->key_mask;
    p.o_f32 = a->o_f32, p.ldof = a->ldof;
    SHG_REQUIRE(!a->o_f32 || (a->ldof % 4 == 0 && (reinterpret_cast<uintptr_t>(a->o_f32) & 15) == 0), SHG_ESHAPE,
                "shg_attn_fwd: o_f32 must be 16-byte aligned with ld %% 4 == 0");
    p.batch = a->batch, p.heads = a->heads, p.lq = a->lq, p.lk = a->lk;
    p.lk16 = (a->lk + 15) / 16 * 16, p.lk32 = (a->lk + 31) / 32 * 32;
    p.n_kbox = (p.lk16 + 63) / 64;
    const int n_pblk = (p.lk32 + 63) / 64;
    const int qk_bytes = 16384 + p.n_kbox * 8192, p_bytes = n_pblk * 16384;
    p.v_off = qk_bytes > p_bytes ? qk_bytes : p_bytes;
    p.o_col = p.lk32;
    const int need = p.o_col + kHeadDim;
    p.tmem_cols = need <= 128 ? 128 : (need <= 256 ? 256 : 512);
    p.scale = a->scale, p.mask_mode = a->mask_mode, p.q_per_frame = a->q_per_frame, p.k_per_frame = a->k_per_frame;
    p.drop.p = a->dropout_seed ? a->dropout_p : 0.f, p.drop.salt = a->dropout_salt, p.drop.seed = a->dropout_seed;
    p.drop_bits = a->dropout_bits;
    SHG_REQUIRE(p.drop.p >= 0.f && p.drop.p < 1.f, SHG_EINVAL, "shg_attn_fwd: dropout_p=%f", (double)p.drop.p);
    const size_t smem = (size_t)p.v_off + p.n_kbox * 8192 + (448 + 512) * 4 + 64 + 1024;
    SHG_REQUIRE(smem <= 227 * 1024, SHG_ESHAPE, "shg_attn_fwd: %zu B of shared memory needed", smem);

    const CUtensorMap* tq = head_tmap(a->q, a->ldq, a->heads, a->lq, a->batch, 128);
    const CUtensorMap* tk = head_tmap(a->k, a->ldk, a->heads, a->lk, a->batch, 64);
    const CUtensorMap* tv = head_tmap(a->v, a->ldv, a->heads, a->lk, a->batch, 64);
    if (!tq || !tk || !tv) return SHG_ECUDA;
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(attn_fwd_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        cudaFuncSetAttribute(attn_fwd_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        cudaFuncSetAttribute(attn_fwd_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        cudaFuncSetAttribute(attn_fwd_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        attr_set = true;
    }
    cudaStream_t s = (cudaStream_t)stream;
    const bool drop = p.drop.p > 0.f;
    dim3 grid((a->lq + 127) / 128, a->heads, a->batch);
    if (g_fwd2_min_lk < 0) {
        const char* e = getenv("SHG_ATTN_FWD2_MIN_LK");
        g_fwd2_min_lk = e ? atoi(e) : 257;
    }
    if (a->lk >= g_fwd2_min_lk) {
        // long key sequences (the 393 visual tokens): chunked online-softmax kernel, two CTAs per SM
        const CUtensorMap* tk2 = head_tmap(a->k, a->ldk, a->heads, a->lk, a->batch, kFwd2Chunk);
        const CUtensorMap* tv2 = head_tmap(a->v, a->ldv, a->heads, a->lk, a->batch, kFwd2Chunk);
        if (!tk2 || !tv2) return SHG_ECUDA;
        const size_t smem2 = 49152 + (448 + 256) * 4 + sizeof(Fwd2Bars) + 16 + 1024;
        static bool attr2 = false;
        if (!attr2) {
            cudaFuncSetAttribute(attn_fwd2_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 113 * 1024);
            cudaFuncSetAttribute(attn_fwd2_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 113 * 1024);
            cudaFuncSetAttribute(attn_fwd2_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 113 * 1024);
            cudaFuncSetAttribute(attn_fwd2_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 113 * 1024);
            attr2 = true;
        }
        if (a->mask_mode == 1) {
            if (drop) shg::launch(attn_fwd2_kernel<true, true>, grid, kFwd2Threads, smem2, s, *tq, *tk2, *tv2, p);
            else shg::launch(attn_fwd2_kernel<true, false>, grid, kFwd2Threads, smem2, s, *tq, *tk2, *tv2, p);
        } else {
            if (drop) shg::launch(attn_fwd2_kernel<false, true>, grid, kFwd2Threads, smem2, s, *tq, *tk2, *tv2, p);
            else shg::launch(attn_fwd2_kernel<false, false>, grid, kFwd2Threads, smem2, s, *tq, *tk2, *tv2, p);
        }
        return check_launch("shg_attn_fwd(chunked)");
    }
    if (a->mask_mode == 1) {
        if (drop) shg::launch(attn_fwd_kernel<true, true>, grid, kAttnThreads, smem, s, *tq, *tk, *tv, p);
        else shg::launch(attn_fwd_kernel<true, false>, grid, kAttnThreads, smem, s, *tq, *tk, *tv, p);
    } else {
        if (drop) shg::launch(attn_fwd_kernel<false, true>, grid, kAttnThreads, smem, s, *tq, *tk, *tv, p);
        else shg::launch(attn_fwd_kernel<false, false>, grid, kAttnThreads, smem, s, *tq, *tk, *tv, p);
    }
    return check_launch("shg_attn_fwd");
}

extern "C" int shg_attn_bwd(const shg_attn_bwd_args* a, void* stream) {
    using namespace shg;
    SHG_REQUIRE(a && a->q && a->k && a->v && a->o && a->d_o && a->lse && a->delta && a->dq && a->dk && a->dv,
                SHG_EINVAL, "shg_attn_bwd: null operand");
    SHG_REQUIRE(a->lk >= 1 && a->lq >= 1, SHG_ESHAPE, "shg_attn_bwd: empty problem");
    SHG_REQUIRE(a->mask_mode >= 0 && a->mask_mode <= 2 && (a->mask_mode != 1 || a->key_mask), SHG_EINVAL,
                "shg_attn_bwd: bad mask");
    SHG_REQUIRE(a->ldo % 8 == 0 && a->lddo % 8 == 0 && a->lddq % 8 == 0 && a->lddk % 8 == 0 && a->lddv % 8 == 0,
                SHG_ESHAPE, "shg_attn_bwd: leading dimensions must be multiples of 8");
    AttnBwdParams p;
    p.o = (const __nv_bfloat16*)a->o, p.d_o = (const __nv_bfloat16*)a->d_o, p.ldo = a->ldo, p.lddo = a->lddo;
    p.o_f32 = a->o_f32, p.ldof = a->ldof;
    SHG_REQUIRE(!a->o_f32 || (a->ldof % 4 == 0 && (reinterpret_cast<uintptr_t>(a->o_f32) & 15) == 0), SHG_ESHAPE,
                "shg_attn_bwd: o_f32 must be 16-byte aligned with ld %% 4 == 0");
    p.lse = a->lse, p.delta = a->delta;
    p.dq = (__nv_bfloat16*)a->dq, p.dk = (__nv_bfloat16*)a->dk, p.dv = (__nv_bfloat16*)a->dv;
    p.lddq = a->lddq, p.lddk = a->lddk, p.lddv = a->lddv;
    p.key_mask = a->key_mask, p.batch = a->batch, p.heads = a->heads, p.lq = a->lq, p.lk = a->lk;
    p.scale = a->scale, p.mask_mode = a->mask_mode, p.q_per_frame = a->q_per_frame, p.k_per_frame = a->k_per_frame;
    p.drop.p = a->dropout_seed ? a->dropout_p : 0.f, p.drop.salt = a->dropout_salt, p.drop.seed = a->dropout_seed;
    SHG_REQUIRE(p.drop.p >= 0.f && p.drop.p < 1.f, SHG_EINVAL, "shg_attn_bwd: dropout_p=%f", (double)p.drop.p);
    p.drop_bits = a->dropout_bits;
    const int drop = p.drop.p > 0.f ? (p.drop_bits ? 2 : 1) : 0;
    const CUtensorMap* tq = head_tmap(a->q, a->ldq, a->heads, a->lq, a->batch, 128);
    const CUtensorMap* tdo = head_tmap(a->d_o, a->lddo, a->heads, a->lq, a->batch, 128);
    const CUtensorMap* tk = head_tmap(a->k, a->ldk, a->heads, a->lk, a->batch, 128);
    const CUtensorMap* tv = head_tmap(a->v, a->ldv, a->heads, a->lk, a->batch, 128);
    if (!tq || !tdo || !tk || !tv) return SHG_ECUDA;
    const size_t smem = kBwdTileBytes + 12288 + 128 + 1024;   // tiles + per-row vectors / key mask / keep words + barriers + slack
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(attn_bwd_dq_kernel<0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        cudaFuncSetAttribute(attn_bwd_dkv_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        cudaFuncSetAttribute(attn_bwd_dq_kernel<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        cudaFuncSetAttribute(attn_bwd_dkv_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        cudaFuncSetAttribute(attn_bwd_dq_kernel<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        cudaFuncSetAttribute(attn_bwd_dkv_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        cudaFuncSetAttribute(attn_bwd_dq_kernel<0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        cudaFuncSetAttribute(attn_bwd_dq_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        cudaFuncSetAttribute(attn_bwd_dq_kernel<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        attr_set = true;
    }
    cudaStream_t s = (cudaStream_t)stream;
    const dim3 gq((a->lq + 127) / 128, a->heads, a->batch), gk((a->lk + 127) / 128, a->heads, a->batch);
    if (g_attn_merge < 0) {
        const char* e = getenv("SHG_ATTN_MERGE");
        g_attn_merge = (e && e[0] == '0') ? 0 : 1;
    }
    if (a->lq <= 128 && a->o_f32 && g_attn_merge) {
        // one query tile: dQ, dK and dV from a single pass (see attn_bwd_dq_kernel<., true>)
        if (drop == 2) shg::launch(attn_bwd_dq_kernel<2, true>, gq, kBwdThreads, smem, s, *tq, *tdo, *tk, *tv, p);
        else if (drop == 1) shg::launch(attn_bwd_dq_kernel<1, true>, gq, kBwdThreads, smem, s, *tq, *tdo, *tk, *tv, p);
        else shg::launch(attn_bwd_dq_kernel<0, true>, gq, kBwdThreads, smem, s, *tq, *tdo, *tk, *tv, p);
        return check_launch("shg_attn_bwd(merged)");
    }
    if (g_attn_bwd2 < 0) {
        const char* e = getenv("SHG_ATTN_BWD2");
        g_attn_bwd2 = (e && e[0] == '0') ? 0 : 1;
    }
    if (a->o_f32 && a->lk <= kB2MaxLen && a->lq <= kB2MaxLen && g_attn_bwd2) {
        // several query tiles: the two-CTA-per-SM passes with TMEM-resident P / dS operands (attn_bwd2_*_kernel)
        const CUtensorMap* tk64 = head_tmap(a->k, a->ldk, a->heads, a->lk, a->batch, kB2Chunk);
        const CUtensorMap* tv64 = head_tmap(a->v, a->ldv, a->heads, a->lk, a->batch, kB2Chunk);
        const CUtensorMap* tq64 = head_tmap(a->q, a->ldq, a->heads, a->lq, a->batch, kB2Chunk);
        const CUtensorMap* tdo64 = head_tmap(a->d_o, a->lddo, a->heads, a->lq, a->batch, kB2Chunk);
        if (!tk64 || !tv64 || !tq64 || !tdo64) return SHG_ECUDA;
        const size_t smem2 = kB2TileBytes + 12288 + sizeof(B2Bars) + 16 + 1024;
        static bool attr2 = false;
        if (!attr2) {
            cudaFuncSetAttribute(attn_bwd2_dq_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 113 * 1024);
            cudaFuncSetAttribute(attn_bwd2_dq_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 113 * 1024);
            cudaFuncSetAttribute(attn_bwd2_dq_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 113 * 1024);
            cudaFuncSetAttribute(attn_bwd2_dkv_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 113 * 1024);
            cudaFuncSetAttribute(attn_bwd2_dkv_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 113 * 1024);
            cudaFuncSetAttribute(attn_bwd2_dkv_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 113 * 1024);
            attr2 = true;
        }
        if (drop == 2) shg::launch(attn_bwd2_dq_kernel<2>, gq, kB2Threads, smem2, s, *tq, *tdo, *tk64, *tv64, p);
        else if (drop == 1) shg::launch(attn_bwd2_dq_kernel<1>, gq, kB2Threads, smem2, s, *tq, *tdo, *tk64, *tv64, p);
        else shg::launch(attn_bwd2_dq_kernel<0>, gq, kB2Threads, smem2, s, *tq, *tdo, *tk64, *tv64, p);
        int rc2 = check_launch("shg_attn_bwd(dq, chunked)");
        if (rc2) return rc2;
        if (drop == 2) shg::launch(attn_bwd2_dkv_kernel<2>, gk, kB2Threads, smem2, s, *tq64, *tdo64, *tk, *tv, p);
        else if (drop == 1) shg::launch(attn_bwd2_dkv_kernel<1>, gk, kB2Threads, smem2, s, *tq64, *tdo64, *tk, *tv, p);
        else shg::launch(attn_bwd2_dkv_kernel<0>, gk, kB2Threads, smem2, s, *tq64, *tdo64, *tk, *tv, p);
        return check_launch("shg_attn_bwd(dkv, chunked)");
    }
    if (drop == 2) shg::launch(attn_bwd_dq_kernel<2, false>, gq, kBwdThreads, smem, s, *tq, *tdo, *tk, *tv, p);
    else if (drop == 1) shg::launch(attn_bwd_dq_kernel<1, false>, gq, kBwdThreads, smem, s, *tq, *tdo, *tk, *tv, p);
    else shg::launch(attn_bwd_dq_kernel<0, false>, gq, kBwdThreads, smem, s, *tq, *tdo, *tk, *tv, p);
    int rc = check_launch("shg_attn_bwd(dq)");
    if (rc) return rc;
    if (drop == 2) shg::launch(attn_bwd_dkv_kernel<2>, gk, kBwdThreads, smem, s, *tq, *tdo, *tk, *tv, p);
    else if (drop == 1) shg::launch(attn_bwd_dkv_kernel<1>, gk, kBwdThreads, smem, s, *tq, *tdo, *tk, *tv, p);
    else shg::launch(attn_bwd_dkv_kernel<0>, gk, kBwdThreads, smem, s, *tq, *tdo, *tk, *tv, p);
    return check_launch("shg_attn_bwd(dkv)");
}
