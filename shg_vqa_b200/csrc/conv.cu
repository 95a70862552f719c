// Conv3d visual front-end (lxrt/modeling_capsbert.py:989-996, 1037-1073) as tcgen05 implicit GEMM — SURVEY §8f-1.
//
//   ZeroPad2d(1) -> Conv3d(C_in, C_out, (5,3,3)) -> GELU, twice: [B,2048,16,7,7] -> [B,768,12,7,7] -> [B,768,8,7,7],
//   then tokens = dropout(cat(cls, permute) + pe).
//
// No im2col buffer exists anywhere.  Activations are channels-last bf16 [B, T, 7, 7, C]; the weight is packed
// [C_out, 45 taps, C_in] (tap = (dt*3 + dh)*3 + dw).  One 5-D TMA box {64 channels, 7 w, 7 h, 2 t, 1 b} starting at
// (c0, dw-1, dh-1, t0+dt, b) IS the im2col tile of tap (dt,dh,dw) for two output frames: the spatial zero padding is
// the TMA's out-of-bounds fill (negative / past-the-end coordinates), rows arrive in (t,h,w) order = output order.
//
// The GEMM is laid out transposed, D^T[channel, position]: the channel dimension (768 / 2048) tiles exactly into
// the 256-row CTA-pair MMA, and the awkward 49-position frames go to the N side, where any multiple of 16 is a legal
// MMA width: each CTA of the pair stages 2 frames (98 rows, padded to 104), N = 208 -> 94 % of the issued MMA
// columns are real outputs (a positions-as-M layout would waste 23 %).  K runs over 45 taps x C/64 chunks
// (1440 / 540 k-blocks per tile), so the epilogue is off the critical path and simply scatters bf16 values
// (thread = channel, warp = 32 consecutive channels of one position = 64 contiguous bytes).
//
//   forward : A = W  [C_out, 45*C_in] K-major 2-D,           B = X  5-D box, K-major   -> + bias, GELU (+pe, dropout)
//   dgrad   : A = W  viewed [C_in, (tap, C_out)] MN-major,   B = dY 5-D box at (1-dw, 1-dh, t-dt): the temporal
//             and spatial borders are again OOB fill                                 -> x gelu'(saved pre-activation)
//   wgrad   : per tap a [C_out x C_in] product over all positions: A = dY [rows, C_out] MN-major, B = shifted X 5-D
//             box MN-major; a k-block is 2 frames = 98 rows padded to 112 with rows that are zeroed once and never
//             written by the TMA.
#include <string.h>

#include "shg_common.cuh"

namespace shg {

constexpr int kConvThreads = 64 + 256;      // TMA warp + MMA warp + 8 epilogue warps
constexpr int kConvMaxStages = 8;
constexpr int kFrame = 49;                  // 7 x 7 positions
constexpr int kRowsCta = 2 * kFrame;        // rows one 5-D box delivers
constexpr int kHalfN = 104;                 // fwd/dgrad: N columns per CTA (98 rounded up to 8)
constexpr int kTaps = 45;
constexpr int kWgK = 112;                   // wgrad: k rows per stage (98 rounded up to 16)

__device__ __forceinline__ void tma_load_5d_pair(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2,
                                                 int c3, int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], "
        "[%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(c0), "r"(c1), "r"(c2), "r"(c3),
        "r"(c4)
        : "memory");
}

struct ConvParams {
    int batch, of;                 // clips; frames per clip in the OUTPUT space of this launch
    int m_total, m_tiles, p_tiles_per_clip, total_tiles;
    int kchunks, kb_total, n_stages;
    const float* bias;             // fwd: [m_total]
    __nv_bfloat16* out;            // row (b, pos) at out + (b*out_rows_per_clip + out_row_off + pos) * ldo
    int64_t ldo;
    int out_rows_per_clip, out_row_off;
    __nv_bfloat16* preact_out;     // fwd: [batch*of*49, m_total] pre-activation (saved for the backward) or null
    const __nv_bfloat16* preact_in;   // dgrad: same layout, multiply by gelu'(.) or null
    const float* pos_emb;          // fwd: fp32 [>= out_row_off + of*49, m_total] added after the activation, or null
    DropCfg drop;                  // fwd: nn.Dropout on the token (index = global out row * m_total + channel)
};

// MODE 0: forward, MODE 1: dgrad
template <int MODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kConvThreads, 1)
conv_gemm_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                 const ConvParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    constexpr uint32_t a_bytes = 128 * 64 * 2;
    constexpr uint32_t b_bytes = kHalfN * 128;
    constexpr uint32_t stage_bytes = a_bytes + b_bytes;              // 29696 = 29 * 1024
    constexpr uint32_t tx_bytes = a_bytes + kRowsCta * 128;          // what the TMA actually delivers per CTA
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + (size_t)p.n_stages * stage_bytes);
    uint64_t* empty_bar = full_bar + kConvMaxStages;
    uint64_t* tmem_full = empty_bar + kConvMaxStages;
    uint64_t* tmem_empty = tmem_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    if (threadIdx.x == 0) {
        for (int s = 0; s < p.n_stages; ++s) {
            mbar_init(full_bar + s, 1);
            mbar_init(empty_bar + s, 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(tmem_full + s, 1);
            mbar_init(tmem_empty + s, 16);           // 8 epilogue warps of each CTA
        }
        fence_barrier_init();
        tma_prefetch_desc(&tma_a);
        tma_prefetch_desc(&tma_b);
    }
    // rows 98..103 of every B stage are never written by the TMA: they only feed accumulator columns that are never
    // stored, but keep them finite
    for (int i = threadIdx.x; i < p.n_stages * (int)((kHalfN - kRowsCta) * 128 / 16); i += kConvThreads) {
        const int s = i / ((kHalfN - kRowsCta) * 8), r = i % ((kHalfN - kRowsCta) * 8);
        *reinterpret_cast<uint4*>(smem + (size_t)s * stage_bytes + a_bytes + kRowsCta * 128 + r * 16) = make_uint4(0, 0, 0, 0);
    }
    fence_proxy_async_smem();
    cluster_sync_all();
    if (warp == 1) tmem_alloc2(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int n_pairs = gridDim.x >> 1, pair = blockIdx.x >> 1;

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = pair; tile < p.total_tiles; tile += n_pairs) {
                const int mblk = tile % p.m_tiles, pt = tile / p.m_tiles;
                const int b = pt / p.p_tiles_per_clip, t0 = (pt % p.p_tiles_per_clip) * 4 + 2 * (int)rank;
                const int m0 = mblk * 256 + (int)rank * 128;
                int tap = 0, ch = 0, dt = 0, dh = 0, dw = 0;
                for (int kb = 0; kb < p.kb_total; ++kb) {
                    mbar_wait(empty_bar + stage, phase ^ 1);
                    uint8_t* sa = smem + (size_t)stage * stage_bytes;
                    uint8_t* sb = sa + a_bytes;
                    if (leader) mbar_expect_tx(full_bar + stage, 2 * tx_bytes);
                    if (MODE == 0) {
                        tma_load_2d_pair(sa, &tma_a, full_bar + stage, kb * 64, m0);
                        tma_load_5d_pair(sb, &tma_b, full_bar + stage, ch * 64, dw - 1, dh - 1, t0 + dt, b);
                    } else {
                        const int c0 = tap * p.m_total + m0;
                        tma_load_2d_pair(sa, &tma_a, full_bar + stage, c0, ch * 64);
                        tma_load_2d_pair(sa + 8192, &tma_a, full_bar + stage, c0 + 64, ch * 64);
                        tma_load_5d_pair(sb, &tma_b, full_bar + stage, ch * 64, 1 - dw, 1 - dh, t0 - dt, b);
                    }
                    if (++ch == p.kchunks) {
                        ch = 0, ++tap;
                        if (++dw == 3) {
                            dw = 0;
                            if (++dh == 3) dh = 0, ++dt;
                        }
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
            const uint32_t idesc = make_idesc_bf16(256, 2 * kHalfN, MODE == 1, 0);
            int stage = 0, as = 0;
            uint32_t phase = 0, aphase = 0;
            for (int tile = pair; tile < p.total_tiles; tile += n_pairs) {
                mbar_wait(tmem_empty + as, aphase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + as * 256;
                for (int kb = 0; kb < p.kb_total; ++kb) {
                    mbar_wait(full_bar + stage, phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + (size_t)stage * stage_bytes);
                    const uint32_t sb = sa + a_bytes;
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {
                        const uint64_t adesc = MODE == 1 ? make_smem_desc(sa + ks * 2048, 8192, 1024)
                                                         : make_smem_desc(sa + ks * 32, 16, 1024);
                        const uint64_t bdesc = make_smem_desc(sb + ks * 32, 16, 1024);
                        umma_bf16_pair(d_tmem, adesc, bdesc, idesc, (kb > 0 || ks > 0) ? 1u : 0u);
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
        // epilogue: thread = channel row of the transposed product, 13 chunks of 16 positions, two warps per lane
        // quarter taking alternate chunks
        const int quarter = warp & 3, sub = (warp - 2) >> 2;
        int as = 0;
        uint32_t aphase = 0;
        const uint32_t leader_empty0 = mapa_u32(smem_u32(tmem_empty), 0);
        const uint32_t key = p.drop.p > 0.f ? drop_key(__ldg(p.drop.seed), p.drop.salt) : 0u;
        const uint32_t th = drop_thresh(p.drop.p);
        const float inv_keep = 1.0f / (1.0f - p.drop.p);
        for (int tile = pair; tile < p.total_tiles; tile += n_pairs) {
            const int mblk = tile % p.m_tiles, pt = tile / p.m_tiles;
            const int b = pt / p.p_tiles_per_clip, t0 = (pt % p.p_tiles_per_clip) * 4;
            const int m = mblk * 256 + (int)rank * 128 + quarter * 32 + lane;
            const float bias = (MODE == 0 && p.bias) ? __ldg(p.bias + m) : 0.f;
            const uint32_t t_row = tmem_base + (uint32_t(quarter * 32) << 16) + as * 256;
            mbar_wait(tmem_full + as, aphase);
            tc_fence_after();
            for (int chunk = sub; chunk < 13; chunk += 2) {
                uint32_t raw[16];
                __syncwarp();
                tmem_ld16(t_row + chunk * 16, raw);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const int c = chunk * 16 + i;
                    const int half = c >= kHalfN ? 1 : 0;
                    const int r = c - half * kHalfN;
                    if (r >= kRowsCta) continue;
                    const int pos = (t0 + 2 * half) * kFrame + r;             // position inside the clip (t, h, w)
                    const size_t prow = (size_t)b * p.of * kFrame + pos;      // row of the compact [B*of*49, M] tensors
                    float v = __uint_as_float(raw[i]) + bias;
                    if (MODE == 0) {
                        if (p.preact_out) p.preact_out[prow * p.m_total + m] = __float2bfloat16(v);
                        v = gelu_erf(v);
                        const size_t orow = (size_t)b * p.out_rows_per_clip + p.out_row_off + pos;
                        if (p.pos_emb) v += __ldg(p.pos_emb + (size_t)(p.out_row_off + pos) * p.m_total + m);
                        if (p.drop.p > 0.f)
                            v = drop_keep(key, (uint64_t)orow * (uint64_t)p.m_total + m, th) ? v * inv_keep : 0.f;
                        p.out[orow * p.ldo + m] = __float2bfloat16(v);
                    } else {
                        if (p.preact_in) v *= gelu_erf_grad(__bfloat162float(p.preact_in[prow * p.m_total + m]));
                        const size_t orow = (size_t)b * p.out_rows_per_clip + p.out_row_off + pos;
                        p.out[orow * p.ldo + m] = __float2bfloat16(v);
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(leader_empty0 + as * 8);
            if (++as == 2) {
                as = 0;
                aphase ^= 1;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc2(tmem_base, 512);
    }
}

// ------------------------------------------------------------------------------------------------ wgrad
struct WgradParams {
    int k_stages;                 // batch * t_out / 2 frame pairs
    int t_out;
    int c_in, c_out, m_tiles, n_tiles, total_tiles, n_stages;
    float* dw;                    // fp32 packed [c_out, 45, c_in], accumulated into
};

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kConvThreads, 1)
conv_wgrad_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                  const WgradParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    constexpr uint32_t sub_bytes = kWgK * 128;                       // one 64-wide MN atom column: 112 k rows
    constexpr uint32_t stage_bytes = 4 * sub_bytes;                  // A: 2 atoms (128 c_out), B: 2 atoms (128 c_in)
    constexpr uint32_t tx_bytes = 4 * kRowsCta * 128;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + (size_t)p.n_stages * stage_bytes);
    uint64_t* empty_bar = full_bar + kConvMaxStages;
    uint64_t* tmem_full = empty_bar + kConvMaxStages;
    uint64_t* tmem_empty = tmem_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    if (threadIdx.x == 0) {
        for (int s = 0; s < p.n_stages; ++s) {
            mbar_init(full_bar + s, 1);
            mbar_init(empty_bar + s, 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(tmem_full + s, 1);
            mbar_init(tmem_empty + s, 16);
        }
        fence_barrier_init();
        tma_prefetch_desc(&tma_a);
        tma_prefetch_desc(&tma_b);
    }
    // k rows 98..111 of every operand atom column are the padding of the 2-frame k-block: zero, and the TMA boxes
    // (98 rows) never touch them
    {
        constexpr int vec_per_sub = (kWgK - kRowsCta) * 128 / 16;    // 112
        for (int i = threadIdx.x; i < p.n_stages * 4 * vec_per_sub; i += kConvThreads) {
            const int sidx = i / vec_per_sub, r = i % vec_per_sub;
            *reinterpret_cast<uint4*>(smem + (size_t)sidx * sub_bytes + kRowsCta * 128 + r * 16) = make_uint4(0, 0, 0, 0);
        }
    }
    fence_proxy_async_smem();
    cluster_sync_all();
    if (warp == 1) tmem_alloc2(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int n_pairs = gridDim.x >> 1, pair = blockIdx.x >> 1;

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = pair; tile < p.total_tiles; tile += n_pairs) {
                const int nblk = tile % p.n_tiles, rem = tile / p.n_tiles;
                const int mblk = rem % p.m_tiles, tap = rem / p.m_tiles;
                const int dt = tap / 9, dh = (tap / 3) % 3, dw = tap % 3;
                const int m0 = mblk * 256 + (int)rank * 128, n0 = nblk * 256 + (int)rank * 128;
                int b = 0, t = 0;
                for (int s = 0; s < p.k_stages; ++s) {
                    mbar_wait(empty_bar + stage, phase ^ 1);
                    uint8_t* sa = smem + (size_t)stage * stage_bytes;
                    uint8_t* sb = sa + 2 * sub_bytes;
                    if (leader) mbar_expect_tx(full_bar + stage, 2 * tx_bytes);
                    tma_load_2d_pair(sa, &tma_a, full_bar + stage, m0, s * kRowsCta);
                    tma_load_2d_pair(sa + sub_bytes, &tma_a, full_bar + stage, m0 + 64, s * kRowsCta);
                    tma_load_5d_pair(sb, &tma_b, full_bar + stage, n0, dw - 1, dh - 1, t + dt, b);
                    tma_load_5d_pair(sb + sub_bytes, &tma_b, full_bar + stage, n0 + 64, dw - 1, dh - 1, t + dt, b);
                    t += 2;
                    if (t == p.t_out) t = 0, ++b;
                    if (++stage == p.n_stages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (leader && lane == 0) {
            const uint32_t idesc = make_idesc_bf16(256, 256, 1, 1);
            int stage = 0, as = 0;
            uint32_t phase = 0, aphase = 0;
            for (int tile = pair; tile < p.total_tiles; tile += n_pairs) {
                mbar_wait(tmem_empty + as, aphase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + as * 256;
                for (int s = 0; s < p.k_stages; ++s) {
                    mbar_wait(full_bar + stage, phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + (size_t)stage * stage_bytes);
                    const uint32_t sb = sa + 2 * sub_bytes;
#pragma unroll
                    for (int ks = 0; ks < kWgK / 16; ++ks) {
                        const uint64_t adesc = make_smem_desc(sa + ks * 2048, sub_bytes, 1024);
                        const uint64_t bdesc = make_smem_desc(sb + ks * 2048, sub_bytes, 1024);
                        umma_bf16_pair(d_tmem, adesc, bdesc, idesc, (s > 0 || ks > 0) ? 1u : 0u);
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
        const int quarter = warp & 3, sub = (warp - 2) >> 2;
        int as = 0;
        uint32_t aphase = 0;
        const uint32_t leader_empty0 = mapa_u32(smem_u32(tmem_empty), 0);
        for (int tile = pair; tile < p.total_tiles; tile += n_pairs) {
            const int nblk = tile % p.n_tiles, rem = tile / p.n_tiles;
            const int mblk = rem % p.m_tiles, tap = rem / p.m_tiles;
            const int m = mblk * 256 + (int)rank * 128 + quarter * 32 + lane;
            float* dst = p.dw + ((size_t)m * kTaps + tap) * p.c_in + nblk * 256;
            const uint32_t t_row = tmem_base + (uint32_t(quarter * 32) << 16) + as * 256;
            mbar_wait(tmem_full + as, aphase);
            tc_fence_after();
            for (int chunk = sub; chunk < 8; chunk += 2) {
                uint32_t raw[32];
                __syncwarp();
                tmem_ld32(t_row + chunk * 32, raw);
                tmem_ld_wait();
                float4* o = reinterpret_cast<float4*>(dst + chunk * 32);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    float4 g = o[i];
                    g.x += __uint_as_float(raw[4 * i]), g.y += __uint_as_float(raw[4 * i + 1]);
                    g.z += __uint_as_float(raw[4 * i + 2]), g.w += __uint_as_float(raw[4 * i + 3]);
                    o[i] = g;
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(leader_empty0 + as * 8);
            if (++as == 2) {
                as = 0;
                aphase ^= 1;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc2(tmem_base, 512);
    }
}

// ------------------------------------------------------------------------------------------------ layout / token glue
// fp32 [B, C, S] (NCDHW, S = T*49) -> bf16 [B, S, C] (channels-last), and back (bf16 -> fp32).  64-channel x 32-position
// tiles through shared memory: reads are contiguous along S, writes along C.
__global__ void ncs_to_nsc_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, int C, int S) {
    __shared__ float tile[64][33];
    const int b = blockIdx.z, c0 = blockIdx.y * 64, s0 = blockIdx.x * 32;
    const int tx = threadIdx.x, ty = threadIdx.y;      // 32 x 8
    for (int j = ty; j < 64; j += 8) {
        const int c = c0 + j, s = s0 + tx;
        tile[j][tx] = (c < C && s < S) ? x[((size_t)b * C + c) * S + s] : 0.f;
    }
    __syncthreads();
    for (int j = ty; j < 32; j += 8) {
        const int s = s0 + j, c = c0 + 2 * tx;
        if (s < S && c < C) {
            const __nv_bfloat162 v = __floats2bfloat162_rn(tile[2 * tx][j], tile[2 * tx + 1][j]);
            *reinterpret_cast<__nv_bfloat162*>(y + ((size_t)b * S + s) * C + c) = v;
        }
    }
}
__global__ void nsc_to_ncs_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ y, int C, int S) {
    __shared__ float tile[32][65];
    const int b = blockIdx.z, c0 = blockIdx.y * 64, s0 = blockIdx.x * 32;
    const int tx = threadIdx.x, ty = threadIdx.y;
    for (int j = ty; j < 32; j += 8) {
        const int s = s0 + j, c = c0 + 2 * tx;
        float2 f = make_float2(0.f, 0.f);
        if (s < S && c < C) f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(x + ((size_t)b * S + s) * C + c));
        tile[j][2 * tx] = f.x, tile[j][2 * tx + 1] = f.y;
    }
    __syncthreads();
    for (int j = ty; j < 64; j += 8) {
        const int c = c0 + j, s = s0 + tx;
        if (c < C && s < S) y[((size_t)b * C + c) * S + s] = tile[tx][j];
    }
}

// tokens[b, 0, :] = dropout(cls + pe[0]) — the row torch.cat((cls_tokens, out), 1) + pe puts in front of the 392 conv
// tokens (lxrt/modeling_capsbert.py:1049-1069)
__global__ void frontend_cls_kernel(const float* __restrict__ cls, const float* __restrict__ pe,
                                    __nv_bfloat16* __restrict__ tokens, int64_t ld, int rows_per_clip, int n, DropCfg drop) {
    const int b = blockIdx.x;
    const uint32_t key = drop.p > 0.f ? drop_key(__ldg(drop.seed), drop.salt) : 0u;
    const uint32_t th = drop_thresh(drop.p);
    const float inv_keep = 1.0f / (1.0f - drop.p);
    const size_t row = (size_t)b * rows_per_clip;
    for (int c = threadIdx.x; c < n; c += blockDim.x) {
        float v = cls[c] + pe[c];
        if (drop.p > 0.f) v = drop_keep(key, (uint64_t)row * n + c, th) ? v * inv_keep : 0.f;
        tokens[row * ld + c] = __float2bfloat16(v);
    }
}

// Backward of the token assembly: for token row `tok` (block) and 8 channels (thread), over the clips
//   g = d_tok[b, tok] * dropmask / (1-p);  d_pe[tok] += sum_b g;  tok == 0: d_cls += sum_b g;
//   tok >= 1: d_pre2[b, tok-1] = g * gelu'(pre2[b, tok-1])   (gradient wrt the second conv's pre-activation)
__global__ void frontend_bwd_kernel(const __nv_bfloat16* __restrict__ d_tok, int64_t ld, const __nv_bfloat16* __restrict__ pre,
                                    __nv_bfloat16* __restrict__ d_pre, float* __restrict__ d_pe, float* __restrict__ d_cls,
                                    int batch, int rows_per_clip, int n, DropCfg drop) {
    const int tok = blockIdx.x, c0 = threadIdx.x * 8;
    if (c0 >= n) return;
    const uint32_t key = drop.p > 0.f ? drop_key(__ldg(drop.seed), drop.salt) : 0u;
    const uint32_t th = drop_thresh(drop.p);
    const float inv_keep = 1.0f / (1.0f - drop.p);
    float sum[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int b = 0; b < batch; ++b) {
        const size_t row = (size_t)b * rows_per_clip + tok;
        const uint4 u = *reinterpret_cast<const uint4*>(d_tok + row * ld + c0);
        float g[8];
        float2 f;
        f = unpack_bf16x2(u.x), g[0] = f.x, g[1] = f.y;
        f = unpack_bf16x2(u.y), g[2] = f.x, g[3] = f.y;
        f = unpack_bf16x2(u.z), g[4] = f.x, g[5] = f.y;
        f = unpack_bf16x2(u.w), g[6] = f.x, g[7] = f.y;
        if (drop.p > 0.f) {
            const uint64_t base = (uint64_t)row * n + c0;               // even
#pragma unroll
            for (int i = 0; i < 8; i += 2) {
                const uint32_t kb = drop_pair(key, (uint32_t)((base + i) >> 1), th);
                g[i] = (kb & 1u) ? g[i] * inv_keep : 0.f;
                g[i + 1] = (kb & 2u) ? g[i + 1] * inv_keep : 0.f;
            }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) sum[i] += g[i];
        if (tok >= 1) {
            const size_t prow = (size_t)b * (rows_per_clip - 1) + (tok - 1);
            const uint4 pu = *reinterpret_cast<const uint4*>(pre + prow * n + c0);
            float x[8];
            f = unpack_bf16x2(pu.x), x[0] = f.x, x[1] = f.y;
            f = unpack_bf16x2(pu.y), x[2] = f.x, x[3] = f.y;
            f = unpack_bf16x2(pu.z), x[4] = f.x, x[5] = f.y;
            f = unpack_bf16x2(pu.w), x[6] = f.x, x[7] = f.y;
            gelu_erf_grad_batch(g, x);
            uint4 o;
            o.x = pack_bf16x2(g[0], g[1]), o.y = pack_bf16x2(g[2], g[3]);
            o.z = pack_bf16x2(g[4], g[5]), o.w = pack_bf16x2(g[6], g[7]);
            *reinterpret_cast<uint4*>(d_pre + prow * n + c0) = o;
        }
    }
    float* dp = d_pe + (size_t)tok * n + c0;
#pragma unroll
    for (int i = 0; i < 8; ++i) dp[i] += sum[i];
    if (tok == 0 && d_cls) {
#pragma unroll
        for (int i = 0; i < 8; ++i) d_cls[c0 + i] += sum[i];
    }
}

static const CUtensorMap* act_map5(const void* x, int c, int t, int batch) {
    uint64_t dims[5] = {(uint64_t)c, 7, 7, (uint64_t)t, (uint64_t)batch};
    uint64_t strides[4] = {(uint64_t)c * 2, (uint64_t)c * 14, (uint64_t)c * 98, (uint64_t)t * c * 98};
    uint32_t box[5] = {64, 7, 7, 2, 1};
    return get_tmap_bf16(x, 5, dims, strides, box);
}

static int check_geometry(const char* who, int batch, int t_in, int c_in, int c_out) {
    SHG_REQUIRE(batch > 0 && t_in >= 8 && (t_in % 4) == 0, SHG_ESHAPE, "%s: batch=%d t_in=%d (t_in must be a multiple of 4, >= 8)",
                who, batch, t_in);
    SHG_REQUIRE(c_in % 256 == 0 && c_out % 256 == 0, SHG_ESHAPE, "%s: channels %d -> %d must be multiples of 256", who, c_in,
                c_out);
    return SHG_OK;
}

template <class K>
static void set_smem_once(K kernel, bool* done) {
    if (!*done) {
        cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        *done = true;
    }
}

}  // namespace shg

extern "C" int shg_conv3d_fwd(const shg_conv3d_args* a, void* stream) {
    using namespace shg;
    SHG_REQUIRE(a && a->x && a->w && a->y, SHG_EINVAL, "shg_conv3d_fwd: null operand");
    if (int rc = check_geometry("shg_conv3d_fwd", a->batch, a->t_in, a->c_in, a->c_out)) return rc;
    const int t_out = a->t_in - 4;
    ConvParams p;
    memset(&p, 0, sizeof(p));
    p.batch = a->batch, p.of = t_out, p.m_total = a->c_out, p.m_tiles = a->c_out / 256;
    p.p_tiles_per_clip = t_out / 4, p.total_tiles = p.m_tiles * a->batch * p.p_tiles_per_clip;
    p.kchunks = a->c_in / 64, p.kb_total = kTaps * p.kchunks, p.n_stages = 7;
    p.bias = a->bias, p.out = (__nv_bfloat16*)a->y, p.ldo = a->ldy;
    p.out_rows_per_clip = a->y_rows_per_clip > 0 ? a->y_rows_per_clip : t_out * kFrame, p.out_row_off = a->y_row_off;
    p.preact_out = (__nv_bfloat16*)a->preact, p.pos_emb = a->pos_emb;
    p.drop.p = a->dropout_seed ? a->dropout_p : 0.f, p.drop.salt = a->dropout_salt, p.drop.seed = a->dropout_seed;
    SHG_REQUIRE(p.drop.p >= 0.f && p.drop.p < 1.f, SHG_EINVAL, "shg_conv3d_fwd: dropout_p=%f", (double)p.drop.p);
    SHG_REQUIRE(p.ldo >= a->c_out, SHG_ESHAPE, "shg_conv3d_fwd: ldy=%lld < c_out", (long long)p.ldo);
    uint64_t dims[2] = {(uint64_t)kTaps * a->c_in, (uint64_t)a->c_out}, strides[1] = {(uint64_t)kTaps * a->c_in * 2};
    uint32_t box[2] = {64, 128};
    const CUtensorMap* ta = get_tmap_bf16(a->w, 2, dims, strides, box);
    const CUtensorMap* tb = act_map5(a->x, a->c_in, a->t_in, a->batch);
    if (!ta || !tb) return SHG_ECUDA;
    const size_t smem = (size_t)p.n_stages * (16384 + kHalfN * 128) + 1024 + 1024;
    static bool attr = false;
    set_smem_once(conv_gemm_kernel<0>, &attr);
    const int grid = 2 * (p.total_tiles < kNumSMs / 2 ? p.total_tiles : kNumSMs / 2);
    shg::launch(conv_gemm_kernel<0>, grid, kConvThreads, smem, (cudaStream_t)stream, *ta, *tb, p);
    return check_launch("shg_conv3d_fwd");
}

extern "C" int shg_conv3d_dgrad(const void* dy, const void* w, const void* preact, void* dx, int batch, int t_in, int c_in,
                                int c_out, void* stream) {
    using namespace shg;
    SHG_REQUIRE(dy && w && dx, SHG_EINVAL, "shg_conv3d_dgrad: null operand");
    if (int rc = check_geometry("shg_conv3d_dgrad", batch, t_in, c_in, c_out)) return rc;
    const int t_out = t_in - 4;
    ConvParams p;
    memset(&p, 0, sizeof(p));
    p.batch = batch, p.of = t_in, p.m_total = c_in, p.m_tiles = c_in / 256;
    p.p_tiles_per_clip = t_in / 4, p.total_tiles = p.m_tiles * batch * p.p_tiles_per_clip;
    p.kchunks = c_out / 64, p.kb_total = kTaps * p.kchunks, p.n_stages = 7;
    p.out = (__nv_bfloat16*)dx, p.ldo = c_in, p.out_rows_per_clip = t_in * kFrame, p.out_row_off = 0;
    p.preact_in = (const __nv_bfloat16*)preact;
    uint64_t dims[2] = {(uint64_t)kTaps * c_in, (uint64_t)c_out}, strides[1] = {(uint64_t)kTaps * c_in * 2};
    uint32_t box[2] = {64, 64};
    const CUtensorMap* ta = get_tmap_bf16(w, 2, dims, strides, box);
    const CUtensorMap* tb = act_map5(dy, c_out, t_out, batch);
    if (!ta || !tb) return SHG_ECUDA;
    const size_t smem = (size_t)p.n_stages * (16384 + kHalfN * 128) + 1024 + 1024;
    static bool attr = false;
    set_smem_once(conv_gemm_kernel<1>, &attr);
    const int grid = 2 * (p.total_tiles < kNumSMs / 2 ? p.total_tiles : kNumSMs / 2);
    shg::launch(conv_gemm_kernel<1>, grid, kConvThreads, smem, (cudaStream_t)stream, *ta, *tb, p);
    return check_launch("shg_conv3d_dgrad");
}

extern "C" int shg_conv3d_wgrad(const void* dy, const void* x, float* dw, int batch, int t_in, int c_in, int c_out,
                                void* stream) {
    using namespace shg;
    SHG_REQUIRE(dy && x && dw, SHG_EINVAL, "shg_conv3d_wgrad: null operand");
    if (int rc = check_geometry("shg_conv3d_wgrad", batch, t_in, c_in, c_out)) return rc;
    SHG_REQUIRE((reinterpret_cast<uintptr_t>(dw) & 15) == 0, SHG_ESHAPE, "shg_conv3d_wgrad: dw must be 16-byte aligned");
    const int t_out = t_in - 4;
    WgradParams p;
    p.k_stages = batch * t_out / 2, p.t_out = t_out, p.c_in = c_in, p.c_out = c_out;
    p.m_tiles = c_out / 256, p.n_tiles = c_in / 256, p.total_tiles = kTaps * p.m_tiles * p.n_tiles, p.n_stages = 3;
    p.dw = dw;
    uint64_t dims[2] = {(uint64_t)c_out, (uint64_t)batch * t_out * kFrame}, strides[1] = {(uint64_t)c_out * 2};
    uint32_t box[2] = {64, (uint32_t)kRowsCta};
    const CUtensorMap* ta = get_tmap_bf16(dy, 2, dims, strides, box);
    const CUtensorMap* tb = act_map5(x, c_in, t_in, batch);
    if (!ta || !tb) return SHG_ECUDA;
    const size_t smem = (size_t)p.n_stages * 4 * kWgK * 128 + 1024 + 1024;
    static bool attr = false;
    set_smem_once(conv_wgrad_kernel, &attr);
    const int grid = 2 * (p.total_tiles < kNumSMs / 2 ? p.total_tiles : kNumSMs / 2);
    shg::launch(conv_wgrad_kernel, grid, kConvThreads, smem, (cudaStream_t)stream, *ta, *tb, p);
    return check_launch("shg_conv3d_wgrad");
}

extern "C" int shg_ncdhw_to_ndhwc(const float* x, void* y, int batch, int channels, int positions, void* stream) {
    using namespace shg;
    SHG_REQUIRE(x && y && batch > 0 && channels > 0 && positions > 0 && channels % 2 == 0, SHG_EINVAL,
                "shg_ncdhw_to_ndhwc: bad arguments");
    dim3 grid((positions + 31) / 32, (channels + 63) / 64, batch), block(32, 8);
    shg::launch(ncs_to_nsc_kernel, grid, block, 0, (cudaStream_t)stream, x, (__nv_bfloat16*)y, channels, positions);
    return check_launch("shg_ncdhw_to_ndhwc");
}

extern "C" int shg_ndhwc_to_ncdhw(const void* x, float* y, int batch, int channels, int positions, void* stream) {
    using namespace shg;
    SHG_REQUIRE(x && y && batch > 0 && channels > 0 && positions > 0 && channels % 2 == 0, SHG_EINVAL,
                "shg_ndhwc_to_ncdhw: bad arguments");
    dim3 grid((positions + 31) / 32, (channels + 63) / 64, batch), block(32, 8);
    shg::launch(nsc_to_ncs_kernel, grid, block, 0, (cudaStream_t)stream, (const __nv_bfloat16*)x, y, channels, positions);
    return check_launch("shg_ndhwc_to_ncdhw");
}

extern "C" int shg_frontend_cls(const float* cls, const float* pe, void* tokens, int64_t ld, int batch, int rows_per_clip,
                                int n, float drop_p, uint32_t drop_salt, const uint32_t* seed, void* stream) {
    using namespace shg;
    SHG_REQUIRE(cls && pe && tokens && batch > 0, SHG_EINVAL, "shg_frontend_cls: bad arguments");
    DropCfg d;
    d.p = seed ? drop_p : 0.f, d.salt = drop_salt, d.seed = seed;
    shg::launch(frontend_cls_kernel, batch, 256, 0, (cudaStream_t)stream, cls, pe, (__nv_bfloat16*)tokens, ld, rows_per_clip, n, d);
    return check_launch("shg_frontend_cls");
}

extern "C" int shg_frontend_bwd(const void* d_tok, int64_t ld, const void* preact, void* d_pre, float* d_pe, float* d_cls,
                                int batch, int rows_per_clip, int n, float drop_p, uint32_t drop_salt, const uint32_t* seed,
                                void* stream) {
    using namespace shg;
    SHG_REQUIRE(d_tok && preact && d_pre && d_pe && batch > 0, SHG_EINVAL, "shg_frontend_bwd: bad arguments");
    SHG_REQUIRE(n % 8 == 0 && ld % 8 == 0 && n <= 8 * 1024, SHG_ESHAPE, "shg_frontend_bwd: n=%d ld=%lld", n, (long long)ld);
    DropCfg d;
    d.p = seed ? drop_p : 0.f, d.salt = drop_salt, d.seed = seed;
    const int threads = ((n / 8 + 31) / 32) * 32;
    shg::launch(frontend_bwd_kernel, rows_per_clip, threads, 0, (cudaStream_t)stream, 
        (const __nv_bfloat16*)d_tok, ld, (const __nv_bfloat16*)preact, (__nv_bfloat16*)d_pre, d_pe, d_cls, batch, rows_per_clip,
        n, d);
    return check_launch("shg_frontend_bwd");
}
