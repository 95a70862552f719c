// Persistent, warp-specialised tcgen05 GEMM for every dense projection on the path.
//
//   warp 0      : TMA producer (one elected lane) — bf16 operand tiles, 128B swizzle, mbarrier ring
//   warp 1      : TMEM allocation (whole warp) + MMA issuer (one lane): tcgen05.mma 128 x block_n x 16,
//                 fp32 accumulators in TMEM, two accumulator stages so the epilogue of tile t overlaps
//                 the main loop of tile t+1
//   warps 2..17 : epilogue — four warps per TMEM lane quarter (16 columns each of a 64-column block), tcgen05.ld with
//                 thread = one output row; bias staged in smem, residual / activation-derivative operands
//                 prefetched from HBM ahead of the accumulator wait; bias / pre-activation store /
//                 GELU|ReLU|tanh / x act'(aux) / residual / bf16|fp32 store or fp32 atomic accumulate (split-K)
//
// Operand majors are runtime-selectable so the forward (x.W^T), dgrad (dy.W) and wgrad (dy^T.x) products
// all read the tensors where they already lie in HBM — no transposed copies.
#include <stdlib.h>

#include "shg_common.cuh"

namespace shg {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 bf16 = one 128-byte swizzle row
constexpr int kEpiSub = 4;          // epilogue warps per TMEM lane quarter
constexpr int kGemmThreads = 64 + kEpiSub * 128;   // TMA warp + MMA warp + 4 * kEpiSub epilogue warps
constexpr int kMaxStages = 8;

struct GemmParams {
    void* d;
    int64_t ldd;
    int m, n, k;
    const float* bias;
    const __nv_bfloat16* residual;
    int64_t ldr;
    __nv_bfloat16* preact;
    const __nv_bfloat16* mul_aux;
    int mul_mode;
    float mul_scale;
    DropCfg drop;
    int act, out_f32, accumulate;
    int tma_store;          // kStoreDirect | kStoreBf16 | kStoreReduceF32 (see epilogue_tile)
    int fast_epi;           // the launch qualifies for epilogue_tile_fast
    int splits, block_n, n_stages;
    int m_tiles, n_tiles, kb_total, kb_per_split;
    unsigned long long* dbg;   // optional phase timestamps of CTA 0 (tools/gemm_phases.py), normally null
    int split_m;               // grouped product: rows >= split_m use the second B operand / bias (INT_MAX: one group)
    const float* bias2;
    int kb_group;              // grouped split-K accumulation: splits starting at k-block >= kb_group add into d2
    void* d2;
    // tail split (CTA-pair kernel): the tiles of the last, sparsely filled wave are cut along K into tail_s slices
    // each, spread over the otherwise idle pairs; partial sums meet in `ws`, the last slice to arrive does the epilogue
    int tail_first;            // first split tile (= number of whole tiles before it); INT_MAX: off
    int tail_s, tail_kb;       // slices per split tile, k-blocks per slice
    int total_items;           // whole tiles + split tiles * tail_s
    float* ws;                 // [split tile][CTA rank][block_n / 4][128 rows] float4, all zero between launches
    unsigned int* ws_cnt;      // [split tile][CTA rank] arrival counters, zero between launches
};

// work item `it` of the pair kernel -> output tile, k-block range, whether it is a K-slice of a split tile
struct PairItem {
    int tile, kb0, kb1;
    bool partial;
};
template <bool TAIL>
__device__ __forceinline__ PairItem pair_item(const GemmParams& p, int it) {
    PairItem r;
    if (!TAIL || it < p.tail_first) {
        r.tile = it;
        const int split = it / (p.m_tiles * p.n_tiles);
        r.kb0 = split * p.kb_per_split;
        r.kb1 = min(r.kb0 + p.kb_per_split, p.kb_total);
        r.partial = false;
    } else {
        const int j = it - p.tail_first;
        r.tile = p.tail_first + j / p.tail_s;
        r.kb0 = (j % p.tail_s) * p.tail_kb;
        r.kb1 = min(r.kb0 + p.tail_kb, p.kb_total);
        r.partial = true;
    }
    return r;
}

#define SHG_STAMP(i) do { if (p.dbg && blockIdx.x == 0) p.dbg[i] = globaltimer_ns(); } while (0)

__device__ __forceinline__ float apply_act(float x, int act) {
    switch (act) {
        case SHG_ACT_GELU_ERF: return gelu_erf(x);
        case SHG_ACT_RELU: return fmaxf(x, 0.0f);
        case SHG_ACT_TANH: return tanhf(x);
        default: return x;
    }
}

constexpr int kStageGroupBytes = 16384;          // one [128 x 64] bf16 or [128 x 32] fp32 staging block per warpgroup
constexpr int kStageOutBytes = 4 * kStageGroupBytes;

// p.tma_store: how D leaves the SM
constexpr int kStoreDirect = 0;   // st.global / atomics from registers (unaligned or plain-fp32 outputs)
constexpr int kStoreBf16 = 1;     // bf16 D (and pre-activation copy): smem staging + TMA bulk store
constexpr int kStoreReduceF32 = 2;   // fp32 accumulation (split-K weight gradients): smem staging + TMA bulk reduce-add

// One output tile's epilogue for one warpgroup (4 warps = the 4 TMEM lane quarters = 128 output rows; thread = row).
// The tile's column blocks (64 wide for bf16 outputs, 32 wide for fp32) are dealt round-robin to the FOUR epilogue
// warpgroups, which run independently of each other: each has its own staging block, its own named barrier and its
// own bulk-store group.  (Round 1 had all 16 warps cooperate on every 64-column block behind two 512-thread barriers
// and one store-issuing thread: 0.8 us per block, 3.5 us for a 256-wide tile — longer than the 3.1 us main loop of a
// K = 768 projection, so the epilogue, not the tensor core, paced those kernels.)
//
// bf16 outputs: the 128 threads write a [128 x 64] block 128B-swizzled (thread = row: 16 B units, conflict-free), one
// thread hands it to a TMA store that moves whole lines and clips the M / N tails.  fp32 accumulation (split-K weight
// gradients): the same with [128 x 32] fp32 blocks and a TMA REDUCE-ADD (cp.reduce.async.bulk.tensor ... .add):
// the L2 does the fp32 adds on whole lines instead of 16-byte atomics scattered over 32 rows per warp instruction.
// `wait_accumulator` blocks until the tile's accumulator is complete in TMEM; HBM operands of the first chunk are
// requested before it.
template <class WaitFn>
__device__ __forceinline__ void epilogue_tile(const GemmParams& p, const float* __restrict__ bias, uint32_t t_row,
                                              int row, bool row_ok, int m0, int n0, int wg, int wtid, uint8_t* stg,
                                              const CUtensorMap* tm_d, const CUtensorMap* tm_pre, void* dbase,
                                              float* sbias, WaitFn wait_accumulator, float4* ws_src = nullptr) {
    constexpr int CH = 16, NV = 2;               // register chunk: 16 columns = 2 x 16-byte bf16 vectors
    const int mode = p.tma_store;
    const int uw = mode == kStoreReduceF32 ? 32 : 64;          // unit block width
    const bool staged = mode != kStoreDirect;
    const bool vec_ok_bf16 = (p.ldd % 8 == 0) && ((reinterpret_cast<uintptr_t>(dbase) & 15) == 0);
    const bool vec_ok_f32 = (p.ldd % 4 == 0) && ((reinterpret_cast<uintptr_t>(dbase) & 15) == 0);
    // the one bf16 side operand of the tile: the saved activation input/output (backward through an activation) or
    // the residual — never both on this path (checked on the host)
    const __nv_bfloat16* side_ptr = p.mul_aux ? p.mul_aux : p.residual;
    const int64_t ld_side = p.mul_aux ? p.ldd : p.ldr;
    const bool vec_ok_side = side_ptr && (ld_side % 8 == 0) && ((reinterpret_cast<uintptr_t>(side_ptr) & 15) == 0);
    const int r_local = row - m0;
    uint8_t* buf = stg + wg * kStageGroupBytes;
    const int bar_id = 1 + wg;
    auto wg_sync = [&] { asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory"); };
    // side operand of a chunk: requested from HBM ahead of its use (first chunk: before the accumulator wait)
    uint4 side[NV];
    auto prefetch = [&](int c) {
        const int col0 = n0 + c;
        if (row_ok && col0 + CH <= p.n && vec_ok_side) {
            const uint4* src = reinterpret_cast<const uint4*>(side_ptr + (size_t)row * ld_side + col0);
#pragma unroll
            for (int i = 0; i < NV; ++i) side[i] = __ldg(src + i);
        }
    };
    if (wg * uw < p.block_n) prefetch(wg * uw);
    // bias of this warpgroup's column blocks -> its 64-float strip of shared memory (zero past N), while the main
    // loop is still running: a per-chunk __ldg would put an L2 round trip on the epilogue's critical path
    if (bias) {
        if (wtid < 64) {
            const int c = wg * uw + (wtid / uw) * 4 * uw + (wtid % uw);       // tile column of strip entry wtid
            sbias[wtid] = (c < p.block_n && n0 + c < p.n) ? __ldg(bias + n0 + c) : 0.f;
        }
        wg_sync();
    }
    wait_accumulator();
    tc_fence_after();
#define EPI_STAMP(i) do { if (p.dbg && wg == 0 && wtid == 0 && blockIdx.x == 0) p.dbg[i] = globaltimer_ns(); } while (0)
    EPI_STAMP(8);
    const int n_pass = p.preact ? 2 : 1;          // the pre-activation copy is a pass of its own over the block
    for (int blk0 = wg * uw; blk0 < p.block_n; blk0 += 4 * uw) {
        if (n0 + blk0 >= p.n) break;               // uniform over the warpgroup
        for (int pass = 0; pass < n_pass; ++pass) {
            const bool is_pre = p.preact && pass == 0;
            if (staged) {
                if (wtid == 0) bulk_wait_read<0>();            // the store that last used this block has read it
                wg_sync();
            }
            EPI_STAMP(9);
            for (int c0 = 0; c0 < uw; c0 += CH) {
                const int c = blk0 + c0;
                const int col0 = n0 + c;
                if (col0 >= p.n) break;                        // uniform; the bulk store clips those columns
                const bool full_chunk = col0 + CH <= p.n;
                float acc[CH];
                {
                    uint32_t raw[CH];
                    __syncwarp();
                    if (ws_src) {
                        // split tile, last slice to arrive: the accumulator is the sum the slices left in the
                        // workspace ([column / 4][row] float4: a warp reads 512 contiguous bytes); the last pass over
                        // a chunk leaves zeros behind for the next launch
                        const bool last_pass = pass == n_pass - 1;
#pragma unroll
                        for (int j = 0; j < CH / 4; ++j) {
                            float4* w4 = ws_src + (size_t)((c >> 2) + j) * 128 + r_local;
                            const float4 v = __ldcg(w4);
                            raw[4 * j] = __float_as_uint(v.x), raw[4 * j + 1] = __float_as_uint(v.y);
                            raw[4 * j + 2] = __float_as_uint(v.z), raw[4 * j + 3] = __float_as_uint(v.w);
                            if (last_pass) __stcg(w4, make_float4(0.f, 0.f, 0.f, 0.f));
                        }
                    } else {
                        tmem_ld16(t_row + c, raw);
                        tmem_ld_wait();
                    }
                    if (c0 == 0) EPI_STAMP(10);
                    if (bias) {
                        const float* sb = sbias + ((blk0 - wg * uw) / (4 * uw)) * uw + c0;
#pragma unroll
                        for (int i = 0; i < CH; i += 4) {
                            const float4 b4 = *reinterpret_cast<const float4*>(sb + i);
                            acc[i] = __uint_as_float(raw[i]) + b4.x, acc[i + 1] = __uint_as_float(raw[i + 1]) + b4.y;
                            acc[i + 2] = __uint_as_float(raw[i + 2]) + b4.z, acc[i + 3] = __uint_as_float(raw[i + 3]) + b4.w;
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < CH; ++i) acc[i] = __uint_as_float(raw[i]);
                    }
                }
                if (!is_pre) {
                    // activation first: only the accumulators (and the packed side operand) are live here, which
                    // leaves the register allocator room to interleave the 32 independent GELU chains
                    if (p.act == SHG_ACT_GELU_ERF) {
                        gelu_erf_batch(acc);
                    } else if (p.act != SHG_ACT_NONE) {
#pragma unroll
                        for (int i = 0; i < CH; ++i) acc[i] = apply_act(acc[i], p.act);
                    }
                    float sv[CH];
                    if (side_ptr) {
                        if (row_ok && full_chunk && vec_ok_side) {
#pragma unroll
                            for (int i = 0; i < NV; ++i) {
                                float2 f;
                                f = unpack_bf16x2(side[i].x), sv[8 * i] = f.x, sv[8 * i + 1] = f.y;
                                f = unpack_bf16x2(side[i].y), sv[8 * i + 2] = f.x, sv[8 * i + 3] = f.y;
                                f = unpack_bf16x2(side[i].z), sv[8 * i + 4] = f.x, sv[8 * i + 5] = f.y;
                                f = unpack_bf16x2(side[i].w), sv[8 * i + 6] = f.x, sv[8 * i + 7] = f.y;
                            }
                        } else {
                            const __nv_bfloat16* sp = side_ptr + (size_t)row * ld_side + col0;
#pragma unroll
                            for (int i = 0; i < CH; ++i)
                                sv[i] = (row_ok && col0 + i < p.n) ? __bfloat162float(sp[i]) : 0.f;
                        }
                        // next chunk of this warpgroup: the rest of the block, else its next block
                        const int cn = c0 + CH < uw ? c + CH : blk0 + 4 * uw;
                        if (cn < p.block_n) prefetch(cn);
                    }
                    if (p.mul_mode) {
                        // backward through an activation: acc *= act'(aux) (aux = saved pre-activation for GELU,
                        // saved output for ReLU)
                        if (p.mul_mode == SHG_MUL_GELU_GRAD) {
                            gelu_erf_grad_batch(acc, sv);
                        } else {
#pragma unroll
                            for (int i = 0; i < CH; ++i) acc[i] = sv[i] > 0.f ? acc[i] * p.mul_scale : 0.f;
                        }
                    }
                    if (p.drop.p > 0.f) {
                        // nn.Dropout on the projection output, before the residual add (BertAttOutput / BertOutput /
                        // decoder dropout1-3 and the decoder FFN's inner dropout)
                        const uint32_t key = drop_key(__ldg(p.drop.seed), p.drop.salt), th = drop_thresh(p.drop.p);
                        const float inv_keep = 1.0f / (1.0f - p.drop.p);
                        const uint64_t base = (uint64_t)row * (uint64_t)p.n + col0;     // col0 is even
#pragma unroll
                        for (int i = 0; i < CH; i += 2) {
                            const uint32_t kb = drop_pair(key, (uint32_t)((base + i) >> 1), th);
                            acc[i] = (kb & 1u) ? acc[i] * inv_keep : 0.f;
                            acc[i + 1] = (kb & 2u) ? acc[i + 1] * inv_keep : 0.f;
                        }
                    }
                    if (p.residual) {
#pragma unroll
                        for (int i = 0; i < CH; ++i) acc[i] += sv[i];
                    }
                }
                if (mode == kStoreBf16) {
                    // [128 x 64] bf16 block: row r at r * 128 B, 16-byte unit u at (u ^ (r & 7))
                    uint8_t* dst = buf + r_local * 128;
                    const int u0 = c0 >> 3;
#pragma unroll
                    for (int j = 0; j < NV; ++j) {
                        uint4 u;
                        u.x = pack_bf16x2(acc[8 * j], acc[8 * j + 1]), u.y = pack_bf16x2(acc[8 * j + 2], acc[8 * j + 3]);
                        u.z = pack_bf16x2(acc[8 * j + 4], acc[8 * j + 5]), u.w = pack_bf16x2(acc[8 * j + 6], acc[8 * j + 7]);
                        *reinterpret_cast<uint4*>(dst + (((u0 + j) ^ (r_local & 7)) << 4)) = u;
                    }
                } else if (mode == kStoreReduceF32) {
                    // [128 x 32] fp32 block: row r at r * 128 B, 16-byte unit (4 floats) u at (u ^ (r & 7))
                    uint8_t* dst = buf + r_local * 128;
                    const int u0 = c0 >> 2;
#pragma unroll
                    for (int j = 0; j < CH / 4; ++j)
                        *reinterpret_cast<float4*>(dst + (((u0 + j) ^ (r_local & 7)) << 4)) =
                            make_float4(acc[4 * j], acc[4 * j + 1], acc[4 * j + 2], acc[4 * j + 3]);
                } else if (row_ok) {
                    if (is_pre) {
                        __nv_bfloat16* pr = p.preact + (size_t)row * p.ldd + col0;
#pragma unroll
                        for (int i = 0; i < CH; ++i)
                            if (col0 + i < p.n) pr[i] = __float2bfloat16(acc[i]);
                    } else if (p.accumulate) {
                        float* o = reinterpret_cast<float*>(dbase) + (size_t)row * p.ldd + col0;
                        if (full_chunk && vec_ok_f32) {
#pragma unroll
                            for (int i = 0; i < CH; i += 4)
                                atomicAdd(reinterpret_cast<float4*>(o + i),
                                          make_float4(acc[i], acc[i + 1], acc[i + 2], acc[i + 3]));
                        } else {
#pragma unroll
                            for (int i = 0; i < CH; ++i)
                                if (col0 + i < p.n) atomicAdd(o + i, acc[i]);
                        }
                    } else if (p.out_f32) {
                        float* o = reinterpret_cast<float*>(dbase) + (size_t)row * p.ldd + col0;
                        if (full_chunk && vec_ok_f32) {
#pragma unroll
                            for (int i = 0; i < CH; i += 4)
                                *reinterpret_cast<float4*>(o + i) = make_float4(acc[i], acc[i + 1], acc[i + 2], acc[i + 3]);
                        } else {
#pragma unroll
                            for (int i = 0; i < CH; ++i)
                                if (col0 + i < p.n) o[i] = acc[i];
                        }
                    } else {
                        __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(dbase) + (size_t)row * p.ldd + col0;
                        if (full_chunk && vec_ok_bf16) {
#pragma unroll
                            for (int i = 0; i < CH; i += 8) {
                                uint4 u;
                                u.x = pack_bf16x2(acc[i], acc[i + 1]), u.y = pack_bf16x2(acc[i + 2], acc[i + 3]);
                                u.z = pack_bf16x2(acc[i + 4], acc[i + 5]), u.w = pack_bf16x2(acc[i + 6], acc[i + 7]);
                                *reinterpret_cast<uint4*>(o + i) = u;
                            }
                        } else {
#pragma unroll
                            for (int i = 0; i < CH; ++i)
                                if (col0 + i < p.n) o[i] = __float2bfloat16(acc[i]);
                        }
                    }
                }
            }
            if (staged) {
                EPI_STAMP(11);
                fence_proxy_async_smem();          // generic-proxy writes of the block -> visible to the TMA engine
                EPI_STAMP(12);
                wg_sync();
                EPI_STAMP(13);
                if (wtid == 0) {
                    if (mode == kStoreReduceF32) tma_reduce_add_2d(tm_d, buf, n0 + blk0, m0);
                    else tma_store_2d(is_pre ? tm_pre : tm_d, buf, n0 + blk0, m0);
                    bulk_commit();
                }
            }
        }
    }
}

// The common epilogues as a tight loop: bias (+ bf16 residual) -> bf16 through a staged TMA store, or fp32 partial
// sums -> staged TMA reduce-add; no activation / dropout / pre-activation copy / activation-derivative operand (those
// launches take epilogue_tile above).  Phase stamps of epilogue_tile on a single-tile CTA (tools/gemm_phases2.py,
// 5632x768x768, 128x256 tile): 0.4 us between "accumulator visible" and the first chunk (staging-buffer wait +
// warpgroup barrier AFTER the accumulator wait) and 0.49 us per 16-column chunk — one warp per scheduler walking a
// dependent chain of ~60 instructions behind a dozen uniform branches on launch parameters, with the TMEM load latency
// exposed every chunk.  Here the buffer wait and barrier are taken BEFORE the accumulator wait, the TMEM load of chunk
// c+1 (and its residual vectors) is in flight while chunk c is converted, and the loop body has no mode branches.
template <class WaitFn>
__device__ __forceinline__ void epilogue_tile_fast(const GemmParams& p, const float* __restrict__ bias, uint32_t t_row,
                                                   int row, bool row_ok, int m0, int n0, int wg, int wtid, uint8_t* stg,
                                                   const CUtensorMap* tm_d, float* sbias, WaitFn wait_accumulator) {
    const bool f32 = p.tma_store == kStoreReduceF32;
    const int uw = f32 ? 32 : 64;
    const int nch = uw >> 4;                     // 16-column chunks per block: 4 (bf16) or 2 (fp32)
    const int r_local = row - m0;
    uint8_t* buf = stg + wg * kStageGroupBytes;
    uint8_t* dst = buf + r_local * 128;
    const int swz = r_local & 7;
    const int bar_id = 1 + wg;
    auto wg_sync = [&] { asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory"); };
    const __nv_bfloat16* res = p.residual ? p.residual + (size_t)row * p.ldr + n0 : nullptr;
    if (bias && wtid < 64) {
        const int c = wg * uw + (wtid / uw) * 4 * uw + (wtid % uw);
        sbias[wtid] = (c < p.block_n && n0 + c < p.n) ? __ldg(bias + n0 + c) : 0.f;
    }
    if (wtid == 0) bulk_wait_read<0>();          // this warpgroup's previous store has read the staging block
    wg_sync();
    wait_accumulator();
    tc_fence_after();
    int k_blk = 0;
    for (int blk0 = wg * uw; blk0 < p.block_n; blk0 += 4 * uw, ++k_blk) {
        if (n0 + blk0 >= p.n) break;             // uniform over the warpgroup
        if (k_blk > 0) {
            if (wtid == 0) bulk_wait_read<0>();
            wg_sync();
        }
        uint32_t raw_a[16], raw_b[16];
        tmem_ld16(t_row + blk0, raw_a);
        // one 16-column chunk: wait for its TMEM load, put the next chunk's load and residual vectors in flight,
        // convert, write the swizzled staging row
        auto chunk = [&](int ci, const uint32_t(&rw)[16], uint32_t(&nxt)[16]) {
            const int c = blk0 + ci * 16;
            const bool live = n0 + c < p.n;      // else past N: the bulk store clips these columns
            uint4 r0 = make_uint4(0, 0, 0, 0), r1 = r0;
            if (res && row_ok && live) {         // residual of this chunk: in flight during the TMEM wait
                r0 = __ldg(reinterpret_cast<const uint4*>(res + c));
                r1 = __ldg(reinterpret_cast<const uint4*>(res + c) + 1);
            }
            tmem_ld_wait();
            if (ci + 1 < nch) tmem_ld16(t_row + c + 16, nxt);        // next chunk's accumulators in flight
            if (!live) return;
            if (f32) {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    *reinterpret_cast<uint4*>(dst + (((ci * 4 + j) ^ swz) << 4)) =
                        make_uint4(rw[4 * j], rw[4 * j + 1], rw[4 * j + 2], rw[4 * j + 3]);
                return;
            }
            const float* sb = sbias + k_blk * uw + ci * 16;
#pragma unroll
            for (int j = 0; j < 2; ++j) {        // 8 columns at a time: one 16-byte unit of the staging row
                float a8[8];
                float4 b0 = make_float4(0.f, 0.f, 0.f, 0.f), b1 = b0;
                if (bias) b0 = *reinterpret_cast<const float4*>(sb + 8 * j), b1 = *reinterpret_cast<const float4*>(sb + 8 * j + 4);
                a8[0] = __uint_as_float(rw[8 * j]) + b0.x, a8[1] = __uint_as_float(rw[8 * j + 1]) + b0.y;
                a8[2] = __uint_as_float(rw[8 * j + 2]) + b0.z, a8[3] = __uint_as_float(rw[8 * j + 3]) + b0.w;
                a8[4] = __uint_as_float(rw[8 * j + 4]) + b1.x, a8[5] = __uint_as_float(rw[8 * j + 5]) + b1.y;
                a8[6] = __uint_as_float(rw[8 * j + 6]) + b1.z, a8[7] = __uint_as_float(rw[8 * j + 7]) + b1.w;
                if (res) {
                    const uint4 rr = j ? r1 : r0;
                    float2 f2;
                    f2 = unpack_bf16x2(rr.x), a8[0] += f2.x, a8[1] += f2.y;
                    f2 = unpack_bf16x2(rr.y), a8[2] += f2.x, a8[3] += f2.y;
                    f2 = unpack_bf16x2(rr.z), a8[4] += f2.x, a8[5] += f2.y;
                    f2 = unpack_bf16x2(rr.w), a8[6] += f2.x, a8[7] += f2.y;
                }
                uint4 u;
                u.x = pack_bf16x2(a8[0], a8[1]), u.y = pack_bf16x2(a8[2], a8[3]);
                u.z = pack_bf16x2(a8[4], a8[5]), u.w = pack_bf16x2(a8[6], a8[7]);
                *reinterpret_cast<uint4*>(dst + (((ci * 2 + j) ^ swz) << 4)) = u;
            }
        };
        chunk(0, raw_a, raw_b);
        chunk(1, raw_b, raw_a);
        if (nch > 2) {
            chunk(2, raw_a, raw_b);
            chunk(3, raw_b, raw_a);
        }
        fence_proxy_async_smem();                // generic-proxy writes of the block -> visible to the TMA engine
        wg_sync();
        if (wtid == 0) {
            if (f32) tma_reduce_add_2d(tm_d, buf, n0 + blk0, m0);
            else tma_store_2d(tm_d, buf, n0 + blk0, m0);
            bulk_commit();
        }
    }
}

// (576 threads = 18 warps: the register file is allocated as if for 20, i.e. 96 registers per thread — a kernel built
// with __maxnreg__(112) is refused at launch, "too many resources requested")
template <bool A_MN, bool B_MN, bool FAST>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
            const __grid_constant__ CUtensorMap tma_d, const __grid_constant__ CUtensorMap tma_pre,
            const __grid_constant__ CUtensorMap tma_b2, const __grid_constant__ CUtensorMap tma_d2,
            const GemmParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* stg = smem;                               // TMA-store staging (present only when p.tma_store)
    if (p.tma_store) smem += kStageOutBytes;
    const uint32_t a_bytes = BM * BK * 2;
    const uint32_t b_bytes = p.block_n * BK * 2;
    const uint32_t stage_bytes = a_bytes + b_bytes;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + (size_t)p.n_stages * stage_bytes);
    uint64_t* empty_bar = full_bar + kMaxStages;
    uint64_t* tmem_full = empty_bar + kMaxStages;
    uint64_t* tmem_empty = tmem_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
    float* s_bias = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(full_bar) + 1024);   // [4 warpgroups][64]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        SHG_STAMP(0);
        for (int s = 0; s < p.n_stages; ++s) {
            mbar_init(full_bar + s, 1);
            mbar_init(empty_bar + s, 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(tmem_full + s, 1);
            mbar_init(tmem_empty + s, 4 * kEpiSub);
        }
        fence_barrier_init();
        tma_prefetch_desc(&tma_a);
        tma_prefetch_desc(&tma_b);
    }
    if (warp == 1) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int total_tiles = p.m_tiles * p.n_tiles * p.splits;
    if (threadIdx.x == 0) SHG_STAMP(1);

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                const int split = tile / (p.m_tiles * p.n_tiles);
                const int rem = tile - split * (p.m_tiles * p.n_tiles);
                const int m0 = (rem / p.n_tiles) * BM, n0 = (rem % p.n_tiles) * p.block_n;
                const int kb0 = split * p.kb_per_split;
                const int kb1 = min(kb0 + p.kb_per_split, p.kb_total);
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(empty_bar + stage, phase ^ 1);
                    uint8_t* sa = smem + (size_t)stage * stage_bytes;
                    uint8_t* sb = sa + a_bytes;
                    mbar_expect_tx(full_bar + stage, stage_bytes);
                    const int k0 = kb * BK;
                    if (!A_MN) {
                        tma_load_2d(sa, &tma_a, full_bar + stage, k0, m0);
                    } else {
                        tma_load_2d(sa, &tma_a, full_bar + stage, m0, k0);
                        tma_load_2d(sa + 8192, &tma_a, full_bar + stage, m0 + 64, k0);
                    }
                    const CUtensorMap* tb = m0 >= p.split_m ? &tma_b2 : &tma_b;
                    if (!B_MN) {
                        tma_load_2d(sb, tb, full_bar + stage, k0, n0);
                    } else {
                        for (int h = 0; h * 64 < p.block_n; ++h)
                            tma_load_2d(sb + h * 8192, tb, full_bar + stage, n0 + h * 64, k0);
                    }
                    if (++stage == p.n_stages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = make_idesc_bf16(BM, p.block_n, A_MN, B_MN);
            int stage = 0, as = 0;
            uint32_t phase = 0, aphase = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                const int split = tile / (p.m_tiles * p.n_tiles);
                const int kb0 = split * p.kb_per_split;
                const int kb1 = min(kb0 + p.kb_per_split, p.kb_total);
                mbar_wait(tmem_empty + as, aphase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + as * 256;
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(full_bar + stage, phase);
                    if (kb == kb0 && tile == (int)blockIdx.x) SHG_STAMP(2);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + (size_t)stage * stage_bytes);
                    const uint32_t sb = sa + a_bytes;
#pragma unroll
                    for (int ks = 0; ks < BK / 16; ++ks) {
                        const uint64_t adesc = A_MN ? make_smem_desc(sa + ks * 2048, 8192, 1024)
                                                    : make_smem_desc(sa + ks * 32, 16, 1024);
                        const uint64_t bdesc = B_MN ? make_smem_desc(sb + ks * 2048, 8192, 1024)
                                                    : make_smem_desc(sb + ks * 32, 16, 1024);
                        umma_bf16(d_tmem, adesc, bdesc, idesc, (kb > kb0 || ks > 0) ? 1u : 0u);
                    }
                    tc_commit(empty_bar + stage);  // frees the smem slot once these MMAs retire
                    if (++stage == p.n_stages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                tc_commit(tmem_full + as);
                if (++as == 2) {
                    as = 0;
                    aphase ^= 1;
                }
            }
        }
    } else {
        // ---------------------------------------------------------------- epilogue (warps 2..17)
        // Four warpgroups; warp w may touch TMEM lane quarter w % 4, so every group of four consecutive warps covers
        // the 128 rows of the tile.  Each group takes every fourth column block and runs on its own (epilogue_tile).
        const int quarter = warp & 3;
        const int wg = (warp - 2) >> 2;
        const int wtid = ((warp - 2) & 3) * 32 + lane;
        int as = 0;
        uint32_t aphase = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            const int rem = tile % (p.m_tiles * p.n_tiles);
            const int m0 = (rem / p.n_tiles) * BM, n0 = (rem % p.n_tiles) * p.block_n;
            const int row = m0 + quarter * 32 + lane;
            const bool row_ok = row < p.m;
            const float* bias = m0 >= p.split_m ? p.bias2 : p.bias;
            const uint32_t t_row = tmem_base + (uint32_t(quarter * 32) << 16) + as * 256;
            const bool second = (tile / (p.m_tiles * p.n_tiles)) * p.kb_per_split >= p.kb_group;
            auto wait_acc = [&] { mbar_wait(tmem_full + as, aphase); if (wg == 0 && wtid == 0 && tile == (int)blockIdx.x) SHG_STAMP(3); };
            if constexpr (FAST)
                epilogue_tile_fast(p, bias, t_row, row, row_ok, m0, n0, wg, wtid, stg, second ? &tma_d2 : &tma_d,
                                   s_bias + wg * 64, wait_acc);
            else
                epilogue_tile(p, bias, t_row, row, row_ok, m0, n0, wg, wtid, stg, second ? &tma_d2 : &tma_d, &tma_pre,
                              second ? p.d2 : p.d, s_bias + wg * 64, wait_acc);
            if (wg == 0 && wtid == 0 && tile == (int)blockIdx.x) SHG_STAMP(4);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_relaxed(tmem_empty + as);
            if (++as == 2) {
                as = 0;
                aphase ^= 1;
            }
        }
        if (wtid == 0) bulk_wait_read<0>();            // staging must outlive the last bulk store's read
        if (wg == 0 && wtid == 0) SHG_STAMP(5);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
    if (threadIdx.x == 0) SHG_STAMP(6);
}

// ---------------------------------------------------------------------------------------------------------
// CTA-pair variant (cta_group::2): two CTAs of a cluster share one 256 x block_n tile.  Each CTA stages its own
// 128 rows of A and HALF of the B tile; the leader's tcgen05.mma reads both shared memories, so per output
// element the pair pulls 1.5x fewer operand bytes through L2 than two independent 128 x block_n tiles — the
// main-loop limiter of the K = 768 projections.  Accumulators: 128 lanes x block_n columns in each CTA's TMEM.
// ---------------------------------------------------------------------------------------------------------
template <bool A_MN, bool B_MN, bool TAIL, bool FAST>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kGemmThreads, 1)
gemm2_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
             const __grid_constant__ CUtensorMap tma_d, const __grid_constant__ CUtensorMap tma_pre,
             const __grid_constant__ CUtensorMap tma_b2, const __grid_constant__ CUtensorMap tma_d2,
             const GemmParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* stg = smem;
    if (p.tma_store) smem += kStageOutBytes;
    const uint32_t a_bytes = BM * BK * 2;
    const int half_n = p.block_n / 2;
    const uint32_t b_bytes = half_n * BK * 2;
    const uint32_t stage_bytes = a_bytes + b_bytes;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + (size_t)p.n_stages * stage_bytes);
    uint64_t* empty_bar = full_bar + kMaxStages;
    uint64_t* tmem_full = empty_bar + kMaxStages;
    uint64_t* tmem_empty = tmem_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
    float* s_bias = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(full_bar) + 1024);   // [4 warpgroups][64]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    if (threadIdx.x == 0) {
        for (int s = 0; s < p.n_stages; ++s) {
            mbar_init(full_bar + s, 1);      // leader's: one arrive.expect_tx covering both CTAs' bytes
            mbar_init(empty_bar + s, 1);     // each CTA's: multicast commit from the leader's MMA
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(tmem_full + s, 1);     // each CTA's: multicast commit
            mbar_init(tmem_empty + s, 8 * kEpiSub);   // leader's: the epilogue warps of both CTAs
        }
        fence_barrier_init();
        tma_prefetch_desc(&tma_a);
        tma_prefetch_desc(&tma_b);
    }
    cluster_sync_all();                      // barriers of both CTAs exist before anything signals across
    if (warp == 1) tmem_alloc2(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int n_pairs = gridDim.x >> 1, pair = blockIdx.x >> 1;
    const int total_tiles = p.total_items;                        // whole tiles (x splits) + K-slices of the split tiles

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int it = pair; it < total_tiles; it += n_pairs) {
                const PairItem w = pair_item<TAIL>(p, it);
                const int rem = w.tile % (p.m_tiles * p.n_tiles);
                const int m0 = (rem / p.n_tiles) * 256 + rank * BM;
                const int n0 = (rem % p.n_tiles) * p.block_n + rank * half_n;
                const int kb0 = w.kb0, kb1 = w.kb1;
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(empty_bar + stage, phase ^ 1);
                    uint8_t* sa = smem + (size_t)stage * stage_bytes;
                    uint8_t* sb = sa + a_bytes;
                    if (leader) mbar_expect_tx(full_bar + stage, 2 * stage_bytes);
                    const int k0 = kb * BK;
                    if (!A_MN) {
                        tma_load_2d_pair(sa, &tma_a, full_bar + stage, k0, m0);
                    } else {
                        tma_load_2d_pair(sa, &tma_a, full_bar + stage, m0, k0);
                        tma_load_2d_pair(sa + 8192, &tma_a, full_bar + stage, m0 + 64, k0);
                    }
                    const CUtensorMap* tb = m0 >= p.split_m ? &tma_b2 : &tma_b;
                    if (!B_MN) {
                        tma_load_2d_pair(sb, tb, full_bar + stage, k0, n0);
                    } else {
                        for (int h = 0; h * 64 < half_n; ++h)
                            tma_load_2d_pair(sb + h * 8192, tb, full_bar + stage, n0 + h * 64, k0);
                    }
                    if (++stage == p.n_stages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (leader && lane == 0) {
            const uint32_t idesc = make_idesc_bf16(256, p.block_n, A_MN, B_MN);
            int stage = 0, as = 0;
            uint32_t phase = 0, aphase = 0;
            for (int it = pair; it < total_tiles; it += n_pairs) {
                const PairItem w = pair_item<TAIL>(p, it);
                const int kb0 = w.kb0, kb1 = w.kb1;
                mbar_wait(tmem_empty + as, aphase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + as * 256;
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(full_bar + stage, phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + (size_t)stage * stage_bytes);
                    const uint32_t sb = sa + a_bytes;
#pragma unroll
                    for (int ks = 0; ks < BK / 16; ++ks) {
                        const uint64_t adesc = A_MN ? make_smem_desc(sa + ks * 2048, 8192, 1024)
                                                    : make_smem_desc(sa + ks * 32, 16, 1024);
                        const uint64_t bdesc = B_MN ? make_smem_desc(sb + ks * 2048, 8192, 1024)
                                                    : make_smem_desc(sb + ks * 32, 16, 1024);
                        umma_bf16_pair(d_tmem, adesc, bdesc, idesc, (kb > kb0 || ks > 0) ? 1u : 0u);
                    }
                    tc_commit_pair(empty_bar + stage);
                    if (++stage == p.n_stages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                tc_commit_pair(tmem_full + as);
                if (++as == 2) {
                    as = 0;
                    aphase ^= 1;
                }
            }
        }
    } else {
        const int quarter = warp & 3;
        const int wg = (warp - 2) >> 2;
        const int wtid = ((warp - 2) & 3) * 32 + lane;
        int as = 0;
        uint32_t aphase = 0;
        const uint32_t leader_empty0 = mapa_u32(smem_u32(tmem_empty), 0);
        volatile int& s_last_slice = *reinterpret_cast<volatile int*>(tmem_slot + 2);
        for (int it = pair; it < total_tiles; it += n_pairs) {
            const PairItem w = pair_item<TAIL>(p, it);
            const int tile = w.tile;
            const int rem = tile % (p.m_tiles * p.n_tiles);
            const int m0 = (rem / p.n_tiles) * 256 + rank * BM, n0 = (rem % p.n_tiles) * p.block_n;
            const int row = m0 + quarter * 32 + lane;
            const bool row_ok = row < p.m;
            const float* bias = m0 >= p.split_m ? p.bias2 : p.bias;
            const uint32_t t_row = tmem_base + (uint32_t(quarter * 32) << 16) + as * 256;
            const bool second = (tile / (p.m_tiles * p.n_tiles)) * p.kb_per_split >= p.kb_group;
            const bool partial = TAIL && w.partial;
            float4* ws4 = nullptr;
            bool run_epilogue = true;
            if (partial) {
                // K-slice of a split tile: this CTA's 128 x block_n partial sums join the workspace ([column / 4][row]
                // float4, so a warp's vector atomics cover 512 contiguous bytes); the slice that finds all the others
                // arrived runs the tile's epilogue from the workspace and leaves it zeroed.
                const int slot = (tile - p.tail_first) * 2 + (int)rank;
                ws4 = reinterpret_cast<float4*>(p.ws) + (size_t)slot * (128 * (p.block_n >> 2));
                const int r_local = quarter * 32 + lane;
                mbar_wait(tmem_full + as, aphase);
                tc_fence_after();
                for (int blk0 = wg * 64; blk0 < p.block_n; blk0 += 256) {
                    if (n0 + blk0 >= p.n) break;
#pragma unroll 1
                    for (int c0 = 0; c0 < 64; c0 += 16) {
                        uint32_t raw[16];
                        tmem_ld16(t_row + blk0 + c0, raw);
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            atomicAdd(ws4 + (size_t)(((blk0 + c0) >> 2) + j) * 128 + r_local,
                                      make_float4(__uint_as_float(raw[4 * j]), __uint_as_float(raw[4 * j + 1]),
                                                  __uint_as_float(raw[4 * j + 2]), __uint_as_float(raw[4 * j + 3])));
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(leader_empty0 + as * 8);      // the accumulator stage is free again
                __threadfence();
                asm volatile("bar.sync 5, 512;" ::: "memory");
                if (wg == 0 && wtid == 0) {
                    const unsigned int n_before = atomicAdd(p.ws_cnt + slot, 1u);
                    s_last_slice = n_before == (unsigned int)(p.tail_s - 1);
                    if (s_last_slice) p.ws_cnt[slot] = 0u;
                }
                asm volatile("bar.sync 5, 512;" ::: "memory");
                run_epilogue = s_last_slice != 0;
                if (run_epilogue) __threadfence();
            }
            auto wait_acc = [&] { if (!partial) mbar_wait(tmem_full + as, aphase); };
            if (run_epilogue) {
                if constexpr (FAST && !TAIL)
                    epilogue_tile_fast(p, bias, t_row, row, row_ok, m0, n0, wg, wtid, stg, second ? &tma_d2 : &tma_d,
                                       s_bias + wg * 64, wait_acc);
                else
                    epilogue_tile(p, bias, t_row, row, row_ok, m0, n0, wg, wtid, stg, second ? &tma_d2 : &tma_d, &tma_pre,
                                  second ? p.d2 : p.d, s_bias + wg * 64, wait_acc, ws4);
            }
            if (!partial) {
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(leader_empty0 + as * 8);
            } else {
                asm volatile("bar.sync 5, 512;" ::: "memory");                   // s_last_slice may be rewritten
            }
            if (++as == 2) {
                as = 0;
                aphase ^= 1;
            }
        }
        if (wtid == 0) bulk_wait_read<0>();
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                      // the peer may still be reading this CTA's smem / signalling its barriers
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc2(tmem_base, 512);
    }
}

// Tile selection.  Every candidate (single CTA 128 x {64,128,192,256}; CTA pair 256 x {128,192,256}) is costed with
// a wave model whose per-k-block time is the largest of: the tensor-core time (2*bn clk), the SM's own operand
// ingress (64 B/clk) and the chip-wide L2 feed shared by the active CTAs; plus a fixed launch/prologue/epilogue
// term (larger for pairs: cluster launch, 2-SM commits).  The four constants were fitted to forced-configuration
// timings of the 45 shapes of the bs32 step (tools/gemm_probe.py tune, profiles/gemm_tune_r01.txt): the model's pick is
// within 2 % of the best measured configuration in aggregate.  K = 768 projections are L2-feed-bound on B200
// (256 x 256 pair tiles need 64 B/clk/SM at full tensor rate; the L2 delivers ~43 B/clk/SM chip-wide), which is why
// wide pair tiles win once a grid is more than a wave deep.
static double config_cost(int m, int n, int k, int splits, int bn, bool pair, bool accumulate) {
    const double kFixSingle = 6000.0, kFixPair = 12000.0, kL2BytesPerClk = 5000.0, kDedup = 0.7;
    const int rows = pair ? 256 : 128;
    const long tiles = (long)((m + rows - 1) / rows) * ((n + bn - 1) / bn) * splits;
    const int slots = pair ? kNumSMs / 2 : kNumSMs;
    const double kb = (double)((k + BK - 1) / BK) / splits;
    const double bytes_cta = (128.0 + (pair ? bn / 2 : bn)) * 128.0;
    const double mma = 2.0 * bn;
    auto wave = [&](double active_ctas) {
        double t = mma;
        if (bytes_cta / 64.0 > t) t = bytes_cta / 64.0;
        const double l2 = bytes_cta * active_ctas / kL2BytesPerClk * kDedup;
        if (l2 > t) t = l2;
        return kb * t;
    };
    const long full = tiles / slots, rem = tiles % slots;
    double t = full * wave(kNumSMs) + (rem ? wave((double)rem * (pair ? 2 : 1)) : 0.0);
    t += (pair ? kFixPair : kFixSingle) + 6.0 * bn + (accumulate ? 4.0 * bn : 0.0);
    return t;
}
// best (pair, block_n) for the problem; pair_ok / single_ok restrict the search (forced by the caller's flags)
static void pick_config(int m, int n, int k, int splits, bool b_mn, bool accumulate, bool pair_ok, bool single_ok,
                        bool* pair, int* bn) {
    double best = 1e30;
    *pair = !single_ok, *bn = 128;
    const int cands[4] = {256, 192, 128, 64};
    for (int pass = 0; pass < 2; ++pass) {
        const bool pr = pass == 1;
        if (pr ? !pair_ok : !single_ok) continue;
        for (int c : cands) {
            if (pr && (c == 64 || (c == 192 && b_mn))) continue;    // pair tiles: 128 | 256 (192 with a K-major B)
            const double t = config_cost(m, n, k, splits, c, pr, accumulate);
            if (t < best) best = t, *pair = pr, *bn = c;
        }
    }
}

}  // namespace shg

static unsigned long long* g_gemm_dbg = nullptr;
// Debug aid (tools/gemm_phases.py): CTA 0 of subsequent shg_gemm launches writes 7 globaltimer stamps to `buf`
// (device memory, 8 x u64); NULL switches it off.  Not part of the product path.
extern "C" int shg_debug_gemm_stamps(void* buf) {
    g_gemm_dbg = (unsigned long long*)buf;
    return SHG_OK;
}

extern "C" int shg_gemm(const shg_gemm_args* a, void* stream) {
    using namespace shg;
    SHG_REQUIRE(a && a->a && a->b && a->d, SHG_EINVAL, "shg_gemm: null operand");
    SHG_REQUIRE(a->m > 0 && a->n > 0 && a->k > 0, SHG_EINVAL, "shg_gemm: empty problem %dx%dx%d", a->m, a->n, a->k);
    const bool accumulate = (a->flags & SHG_GEMM_ACCUMULATE) != 0;
    int splits = a->splits;
    if (splits < 1) {
        // auto split-K (accumulating products only): enough CTAs to cover the chip, >= 8 k-blocks per split
        splits = 1;
        if (accumulate) {
            const int tiles = ((a->m + BM - 1) / BM) * ((a->n + 255) / 256);
            const int kb = (a->k + BK - 1) / BK;
            splits = (kNumSMs + tiles / 2) / tiles;
            if (splits > kb / 8) splits = kb / 8;
            if (splits < 1) splits = 1;
        }
    }
    SHG_REQUIRE(splits == 1 || accumulate, SHG_EINVAL, "shg_gemm: splits>1 needs SHG_GEMM_ACCUMULATE");
    SHG_REQUIRE(!accumulate || !(a->preact || a->residual || a->act || a->mul_aux), SHG_EINVAL,
                "shg_gemm: accumulate mode takes no activation/residual/preact");
    SHG_REQUIRE(!(a->residual && a->mul_aux), SHG_EINVAL, "shg_gemm: residual and mul_aux are mutually exclusive");
    GemmParams p;
    p.dbg = g_gemm_dbg;
    p.d = a->d, p.ldd = a->ldd, p.m = a->m, p.n = a->n, p.k = a->k;
    p.bias = a->bias, p.residual = (const __nv_bfloat16*)a->residual, p.ldr = a->ldr;
    p.preact = (__nv_bfloat16*)a->preact, p.act = a->act;
    p.mul_aux = (const __nv_bfloat16*)a->mul_aux, p.mul_mode = a->mul_aux ? a->mul_mode : 0;
    p.mul_scale = a->mul_scale != 0.f ? a->mul_scale : 1.0f;
    p.drop.p = a->dropout_seed ? a->dropout_p : 0.f, p.drop.salt = a->dropout_salt, p.drop.seed = a->dropout_seed;
    SHG_REQUIRE(p.drop.p >= 0.f && p.drop.p < 1.f, SHG_EINVAL, "shg_gemm: dropout_p=%f", (double)p.drop.p);
    SHG_REQUIRE(p.drop.p == 0.f || (a->n % 2 == 0 && !accumulate), SHG_ESHAPE, "shg_gemm: dropout needs even N and a plain store");
    p.out_f32 = (a->flags & SHG_GEMM_OUT_F32) || accumulate, p.accumulate = accumulate;
    const bool kgrouped = a->d2 != nullptr;     // two K-ranges accumulating into two outputs (stacked-row weight gradients)
    SHG_REQUIRE(!kgrouped || (accumulate && a->k_split > 0 && a->k_split < a->k && a->k_split % BK == 0 && a->splits <= 0),
                SHG_ESHAPE, "shg_gemm: d2 needs SHG_GEMM_ACCUMULATE, 0 < k_split < k, k_split %% 64 == 0 and automatic splits");
    p.kb_group = kgrouped ? a->k_split / BK : 0x7fffffff, p.d2 = a->d2;
    const bool grouped = a->b2 != nullptr;
    p.split_m = grouped ? a->split_m : 0x7fffffff, p.bias2 = a->bias2;
    SHG_REQUIRE(!grouped || (a->split_m > 0 && a->split_m < a->m && a->split_m % BM == 0 && !a->a_mn_major && !accumulate),
                SHG_ESHAPE, "shg_gemm: grouped product needs 0 < split_m < m, split_m %% 128 == 0, row-major A, no split-K");
    p.m_tiles = (a->m + BM - 1) / BM;
    // block_n: 0 = choose everything; low bits = tile width; | 1024 forces CTA-pair tiles (256 x bn), | 2048 forbids them
    const int bn_req = a->block_n & 1023;            // (| 4096: direct-store epilogue, for A/B comparison)
    bool pair = (a->block_n & 1024) != 0;
    int bn = bn_req;
    if (!bn_req) {
        const bool forced = (a->block_n & (1024 | 2048)) != 0;
        pick_config(a->m, a->n, a->k, splits, a->b_mn_major != 0, accumulate, !forced || pair, !forced || !pair, &pair, &bn);
        if (pair && grouped && a->split_m % 256 != 0)       // a 256-row pair tile must not straddle the groups
            pick_config(a->m, a->n, a->k, splits, a->b_mn_major != 0, accumulate, false, true, &pair, &bn);
    }
    SHG_REQUIRE(!(pair && grouped && a->split_m % 256 != 0), SHG_ESHAPE, "shg_gemm: CTA-pair tiles need split_m %% 256 == 0");
    SHG_REQUIRE(bn == 64 || bn == 128 || bn == 192 || bn == 256, SHG_ESHAPE, "shg_gemm: block_n=%d", bn);
    SHG_REQUIRE(!pair || bn == 128 || bn == 256 || (bn == 192 && !a->b_mn_major), SHG_ESHAPE,
                "shg_gemm: CTA-pair tiles need block_n 128 or 256 (192 with a K-major B operand)");
    if (pair) p.m_tiles = (a->m + 255) / 256;
    p.block_n = bn;
    p.n_tiles = (a->n + bn - 1) / bn;
    p.kb_total = (a->k + BK - 1) / BK;
    p.splits = splits > p.kb_total ? p.kb_total : splits;
    p.kb_per_split = (p.kb_total + p.splits - 1) / p.splits;
    if (kgrouped) {
        // a split never straddles the two K-ranges: the largest divisor of the boundary not above the chosen depth
        int per = p.kb_per_split < p.kb_group ? p.kb_per_split : p.kb_group;
        while (p.kb_group % per) --per;
        p.kb_per_split = per;
    }
    p.splits = (p.kb_total + p.kb_per_split - 1) / p.kb_per_split;  // no empty split
    const int stage_bytes = BM * BK * 2 + (pair ? bn / 2 : bn) * BK * 2;
    const int bar_bytes = 2048;          // barriers + TMEM slot + flag | per-warpgroup bias strips
    // how D leaves the SM (see epilogue_tile): bf16 outputs with TMA-legal geometry through staged bulk stores, fp32
    // accumulation (split-K weight gradients) through staged bulk reduce-adds; | 4096 in block_n forces the direct path
    const bool direct = (a->block_n & 4096) != 0;
    p.tma_store = kStoreDirect;
    if (!direct && !p.out_f32 && a->ldd % 8 == 0 && (reinterpret_cast<uintptr_t>(a->d) & 15) == 0 &&
        (!a->preact || (reinterpret_cast<uintptr_t>(a->preact) & 15) == 0))
        p.tma_store = kStoreBf16;
    else if (!direct && accumulate && a->ldd % 4 == 0 && (reinterpret_cast<uintptr_t>(a->d) & 15) == 0 &&
             (!a->d2 || (reinterpret_cast<uintptr_t>(a->d2) & 15) == 0))
        p.tma_store = kStoreReduceF32;
    // the tight epilogue (epilogue_tile_fast): staged output, bias / residual only; a residual must be vector-loadable
    // in whole 16-column chunks.  | 32768 in block_n forces the general epilogue (A/B comparison)
    p.fast_epi = (p.tma_store != kStoreDirect && !a->preact && a->act == SHG_ACT_NONE && !a->mul_aux && p.drop.p == 0.f &&
                  !(a->block_n & 32768) && (p.tma_store == kStoreBf16 || !a->bias) &&
                  (!a->residual || (a->ldr % 8 == 0 && (reinterpret_cast<uintptr_t>(a->residual) & 15) == 0 && a->n % 16 == 0)))
                     ? 1 : 0;
    const int out_bytes = p.tma_store ? kStageOutBytes : 0;
    int n_stages = (227 * 1024 - 1024 - bar_bytes - out_bytes) / stage_bytes;
    if (n_stages > kMaxStages) n_stages = kMaxStages;
    p.n_stages = n_stages;
    const size_t smem = (size_t)n_stages * stage_bytes + bar_bytes + out_bytes + 1024;

    // TMA descriptors over the operands as they lie in HBM
    uint64_t dims[2], strides[1];
    uint32_t box[2];
    const CUtensorMap *ta, *tb;
    if (!a->a_mn_major) { dims[0] = a->k, dims[1] = a->m, box[0] = BK, box[1] = BM; }
    else { dims[0] = a->m, dims[1] = a->k, box[0] = 64, box[1] = BK; }
    strides[0] = (uint64_t)a->lda * 2;
    ta = get_tmap_bf16(a->a, 2, dims, strides, box);
    if (!ta) return SHG_ECUDA;
    if (!a->b_mn_major) { dims[0] = a->k, dims[1] = a->n, box[0] = BK, box[1] = (uint32_t)(pair ? bn / 2 : bn); }
    else { dims[0] = a->n, dims[1] = a->k, box[0] = 64, box[1] = BK; }
    strides[0] = (uint64_t)a->ldb * 2;
    tb = get_tmap_bf16(a->b, 2, dims, strides, box);
    if (!tb) return SHG_ECUDA;
    const CUtensorMap* tb2 = tb;
    if (grouped) {
        tb2 = get_tmap_bf16(a->b2, 2, dims, strides, box);     // same geometry as B
        if (!tb2) return SHG_ECUDA;
    }
    const CUtensorMap *td = ta, *tpre = ta, *td2 = ta;          // placeholders when the staged store is off
    if (p.tma_store == kStoreBf16) {
        dims[0] = a->n, dims[1] = a->m, box[0] = 64, box[1] = BM;
        strides[0] = (uint64_t)a->ldd * 2;
        td = get_tmap_bf16(a->d, 2, dims, strides, box);
        if (!td) return SHG_ECUDA;
        if (a->preact) {
            tpre = get_tmap_bf16(a->preact, 2, dims, strides, box);
            if (!tpre) return SHG_ECUDA;
        }
    } else if (p.tma_store == kStoreReduceF32) {
        dims[0] = a->n, dims[1] = a->m, box[0] = 32, box[1] = BM;
        strides[0] = (uint64_t)a->ldd * 4;
        td = get_tmap(a->d, 2, dims, strides, box, kTmapF32);
        if (!td) return SHG_ECUDA;
        td2 = td;
        if (a->d2) {
            td2 = get_tmap(a->d2, 2, dims, strides, box, kTmapF32);
            if (!td2) return SHG_ECUDA;
        }
    }

    int total = p.m_tiles * p.n_tiles * p.splits;
    p.tail_first = 0x7fffffff, p.tail_s = 0, p.tail_kb = 0, p.total_items = total, p.ws = nullptr, p.ws_cnt = nullptr;
    SHG_REQUIRE(a->max_ctas >= 0, SHG_EINVAL, "shg_gemm: max_ctas=%d", a->max_ctas);
    const int cta_cap = a->max_ctas > 0 && a->max_ctas < kNumSMs ? (a->max_ctas < 2 ? 2 : a->max_ctas) : kNumSMs;
    if (pair && p.splits == 1 && !kgrouped && !accumulate && a->workspace && !(a->block_n & 16384) && cta_cap == kNumSMs) {
        // Tail split: with T tiles on P = 74 pair slots, the last wave holds only T % P tiles while the other pairs
        // idle for a whole tile time — at M = 12576 (bs32 x 393 tokens) that is 2 of 150 (N = 768), 6 of 450 (N = 2304)
        // and 8 of 600 (N = 3072) tiles: 3 waves for 2.03 waves of work.  When the last wave is at most a quarter
        // full, its tiles are cut along K into S slices spread over the idle pairs (fp32 partial sums through L2).
        const int slots = kNumSMs / 2;
        const int rem = total % slots;
        // Measured (profiles/gemm_tune_r02.txt, p256 vs p256notail): the split costs ~8 us per launch (fp32 atomics of
        // the partial tiles through L2, a device-wide fence, the counter round trip and an epilogue fed from L2), so it
        // pays only when a whole extra wave is long: K >= 2560 (12576x768x3072: 57.6 -> 54.0 us; 12576x768x15360:
        // 255 -> 210 us) and loses at K = 768 (22.6 -> 29.2 us), where the plain schedule stays.
        if (total > slots && rem > 0 && rem * 4 <= slots && p.kb_total >= 40) {
            int S = slots / rem;
            if (S > p.kb_total) S = p.kb_total;
            const int per = (p.kb_total + S - 1) / S;
            S = (p.kb_total + per - 1) / per;
            const size_t need = 1024 + (size_t)rem * 2 * 128 * bn * sizeof(float);
            if (S >= 2 && (size_t)a->workspace_bytes >= need && (reinterpret_cast<uintptr_t>(a->workspace) & 15) == 0) {
                p.tail_first = total - rem, p.tail_s = S, p.tail_kb = per;
                p.total_items = p.tail_first + rem * S;
                p.ws_cnt = reinterpret_cast<unsigned int*>(a->workspace);
                p.ws = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(a->workspace) + 1024);
                total = p.total_items;
            }
        }
    }
    int grid = total < cta_cap ? total : cta_cap;
    if (pair) grid = 2 * (total < cta_cap / 2 ? total : cta_cap / 2);
    cudaStream_t s = (cudaStream_t)stream;
#define SHG_GEMM2_LAUNCH_T(AM, BMJ, TL, FS)                                                                    \
    do {                                                                                                       \
        static bool attr_set2 = false;                                                                         \
        if (!attr_set2) {                                                                                      \
            cudaFuncSetAttribute(gemm2_kernel<AM, BMJ, TL, FS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024); \
            attr_set2 = true;                                                                                  \
        }                                                                                                      \
        shg::launch(gemm2_kernel<AM, BMJ, TL, FS>, grid, kGemmThreads, smem, s, *ta, *tb, *td, *tpre, *tb2, *td2, p); \
    } while (0)
#define SHG_GEMM2_LAUNCH(AM, BMJ)                                                                              \
    do {                                                                                                       \
        if (p.tail_s > 0) SHG_GEMM2_LAUNCH_T(AM, BMJ, true, false);                                            \
        else if (p.fast_epi) SHG_GEMM2_LAUNCH_T(AM, BMJ, false, true);                                         \
        else SHG_GEMM2_LAUNCH_T(AM, BMJ, false, false);                                                        \
    } while (0)
    if (pair) {
        if (!a->a_mn_major && !a->b_mn_major) SHG_GEMM2_LAUNCH(false, false);
        else if (!a->a_mn_major && a->b_mn_major) SHG_GEMM2_LAUNCH(false, true);
        else if (a->a_mn_major && a->b_mn_major) SHG_GEMM2_LAUNCH(true, true);
        else SHG_GEMM2_LAUNCH(true, false);
        return check_launch("shg_gemm(pair)");
    }
#undef SHG_GEMM2_LAUNCH
#define SHG_GEMM_LAUNCH_F(AM, BMJ, FS)                                                                         \
    do {                                                                                                      \
        static bool attr_set = false;                                                                         \
        if (!attr_set) {                                                                                      \
            cudaFuncSetAttribute(gemm_kernel<AM, BMJ, FS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024); \
            attr_set = true;                                                                                  \
        }                                                                                                     \
        shg::launch(gemm_kernel<AM, BMJ, FS>, grid, kGemmThreads, smem, s, *ta, *tb, *td, *tpre, *tb2, *td2, p); \
    } while (0)
#define SHG_GEMM_LAUNCH(AM, BMJ)                                                                              \
    do {                                                                                                      \
        if (p.fast_epi) SHG_GEMM_LAUNCH_F(AM, BMJ, true);                                                     \
        else SHG_GEMM_LAUNCH_F(AM, BMJ, false);                                                               \
    } while (0)
    if (!a->a_mn_major && !a->b_mn_major) SHG_GEMM_LAUNCH(false, false);
    else if (!a->a_mn_major && a->b_mn_major) SHG_GEMM_LAUNCH(false, true);
    else if (a->a_mn_major && a->b_mn_major) SHG_GEMM_LAUNCH(true, true);
    else SHG_GEMM_LAUNCH(true, false);
#undef SHG_GEMM_LAUNCH
    return check_launch("shg_gemm");
}
