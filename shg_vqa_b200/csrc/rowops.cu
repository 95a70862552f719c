// HBM-bound row kernels: LayerNorm(+residual) forward/backward, embedding gathers fused with LayerNorm,
// column sums (bias gradients), row-broadcast adds and small dtype helpers.  One warp per row, 128-bit loads
// and stores, fp32 statistics via warp shuffles; nothing here is reshaped into a GEMM.
#include "shg_common.cuh"

namespace shg {

constexpr int kRowThreads = 256;
constexpr int kRowWarps = kRowThreads / 32;

__device__ __forceinline__ void load8(const __nv_bfloat16* p, float (&v)[8]) {
    const uint4 u = *reinterpret_cast<const uint4*>(p);
    float2 f;
    f = unpack_bf16x2(u.x), v[0] = f.x, v[1] = f.y;
    f = unpack_bf16x2(u.y), v[2] = f.x, v[3] = f.y;
    f = unpack_bf16x2(u.z), v[4] = f.x, v[5] = f.y;
    f = unpack_bf16x2(u.w), v[6] = f.x, v[7] = f.y;
}
__device__ __forceinline__ void load8(const float* p, float (&v)[8]) {
    const float4 a = *reinterpret_cast<const float4*>(p);
    const float4 b = *reinterpret_cast<const float4*>(p + 4);
    v[0] = a.x, v[1] = a.y, v[2] = a.z, v[3] = a.w, v[4] = b.x, v[5] = b.y, v[6] = b.z, v[7] = b.w;
}
__device__ __forceinline__ void store8(__nv_bfloat16* p, const float (&v)[8]) {
    uint4 u;
    u.x = pack_bf16x2(v[0], v[1]), u.y = pack_bf16x2(v[2], v[3]);
    u.z = pack_bf16x2(v[4], v[5]), u.w = pack_bf16x2(v[6], v[7]);
    *reinterpret_cast<uint4*>(p) = u;
}
__device__ __forceinline__ void store8(float* p, const float (&v)[8]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}

// v[0..8) *= keep/(1-p) for the 8 consecutive elements starting at (even) flat index idx
__device__ __forceinline__ void drop8(float (&v)[8], uint32_t key, uint32_t th, float inv_keep, uint64_t idx) {
#pragma unroll
    for (int t = 0; t < 8; t += 2) {
        const uint32_t kb = drop_pair(key, (uint32_t)((idx + t) >> 1), th);
        v[t] = (kb & 1u) ? v[t] * inv_keep : 0.f;
        v[t + 1] = (kb & 2u) ? v[t + 1] * inv_keep : 0.f;
    }
}

// CH = 16-byte chunks of 8 elements per lane: n = 256 * CH  (768 -> 3, 1536 -> 6)
template <int CH>
__device__ __forceinline__ void row_stats(const float (&v)[CH][8], float inv_n, float eps, float& mean, float& rstd) {
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < CH; ++j)
#pragma unroll
        for (int i = 0; i < 8; ++i) s += v[j][i];
    mean = warp_sum(s) * inv_n;
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < CH; ++j)
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float d = v[j][i] - mean;
            q += d * d;
        }
    rstd = rsqrtf(warp_sum(q) * inv_n + eps);
}

template <int CH>
__global__ void __launch_bounds__(kRowThreads)
layernorm_fwd_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ residual,
                     const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                     __nv_bfloat16* __restrict__ y, __nv_bfloat16* __restrict__ y2,
                     const __nv_bfloat16* __restrict__ addend, int addend_rows, float* __restrict__ mean_out,
                     float* __restrict__ rstd_out, int rows, int split_row, const float* __restrict__ gamma2,
                     const float* __restrict__ beta2) {
    constexpr int N = 256 * CH;
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * kRowWarps + (threadIdx.x >> 5);
    if (row >= rows) return;
    if (row >= split_row) gamma = gamma2, beta = beta2;      // second parameter group (warp-uniform)
    float v[CH][8];
#pragma unroll
    for (int j = 0; j < CH; ++j) {
        const int c = (j * 32 + lane) * 8;
        load8(x + (size_t)row * N + c, v[j]);
        if (residual) {
            float r[8];
            load8(residual + (size_t)row * N + c, r);
#pragma unroll
            for (int i = 0; i < 8; ++i) v[j][i] += r[i];
        }
    }
    float mean, rstd;
    row_stats<CH>(v, 1.0f / N, eps, mean, rstd);
    if (mean_out && lane == 0) mean_out[row] = mean, rstd_out[row] = rstd;
    const int arow = addend ? (addend_rows > 0 ? row % addend_rows : row) : 0;
#pragma unroll
    for (int j = 0; j < CH; ++j) {
        const int c = (j * 32 + lane) * 8;
        float g[8], b[8], o[8];
        load8(gamma + c, g);
        load8(beta + c, b);
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = (v[j][i] - mean) * rstd * g[i] + b[i];
        store8(y + (size_t)row * N + c, o);
        if (y2) {
            float a[8];
            load8(addend + (size_t)arow * N + c, a);
#pragma unroll
            for (int i = 0; i < 8; ++i) o[i] += a[i];
            store8(y2 + (size_t)row * N + c, o);
        }
    }
}

// dx = rstd * (g*gamma - mean(g*gamma) - xhat * mean(g*gamma*xhat)),  g = dy (+ dy2)
// dgamma += sum_rows g*xhat ; dbeta += sum_rows g ; dxsum += sum_rows dx (bias gradient of the Linear that fed x)
template <int CH>
__global__ void __launch_bounds__(kRowThreads)
layernorm_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ dy2,
                     const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ residual,
                     const float* __restrict__ gamma, const float* __restrict__ mean_in,
                     const float* __restrict__ rstd_in, __nv_bfloat16* __restrict__ dx, float* __restrict__ dgamma,
                     float* __restrict__ dbeta, float* __restrict__ dxsum, int rows,
                     __nv_bfloat16* __restrict__ dx_drop, DropCfg out_drop, DropCfg in_drop, int split_row,
                     const float* __restrict__ gamma2, float* __restrict__ dgamma2, float* __restrict__ dbeta2,
                     float* __restrict__ dxsum2) {
    constexpr int N = 256 * CH;
    // grouped launch: blockIdx.y = 1 handles rows [split_row, rows) with the second parameter / gradient set
    int row_begin = 0;
    if (blockIdx.y == 1) {
        row_begin = split_row;
        gamma = gamma2, dgamma = dgamma2, dbeta = dbeta2, dxsum = dxsum2;
    } else if (gridDim.y == 2) {
        rows = split_row;
    }
    __shared__ float red[kRowWarps][256 + 8];
    uint32_t okey = 0, oth = 0, ikey = 0, ith = 0;
    float oinv = 1.f, iinv = 1.f;
    if (out_drop.p > 0.f) okey = drop_key(__ldg(out_drop.seed), out_drop.salt), oth = drop_thresh(out_drop.p), oinv = 1.f / (1.f - out_drop.p);
    if (in_drop.p > 0.f) ikey = drop_key(__ldg(in_drop.seed), in_drop.salt), ith = drop_thresh(in_drop.p), iinv = 1.f / (1.f - in_drop.p);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float ag[CH][8], ab[CH][8], ax[CH][8];
#pragma unroll
    for (int j = 0; j < CH; ++j)
#pragma unroll
        for (int i = 0; i < 8; ++i) ag[j][i] = ab[j][i] = ax[j][i] = 0.f;
    float gm[CH][8];
#pragma unroll
    for (int j = 0; j < CH; ++j) load8(gamma + (j * 32 + lane) * 8, gm[j]);

    // The loop is a dependent chain per row (loads -> two warp reductions -> stores) and the kernel keeps 72 column
    // accumulators per thread (one block of 8 warps per SM): with only the current row's 6 x 512 B in flight per warp
    // it ran at 1.3 TB/s.  The NEXT row's dy / x (packed, 24 registers at width 768) are requested before the current
    // row's math; at width 1536 the register file has no room for that.
    constexpr bool kPrefetch = CH <= 3;
    uint4 ndy[kPrefetch ? CH : 1], nx[kPrefetch ? CH : 1];
    float nmean = 0.f, nrstd = 0.f;
    auto fetch = [&](int r) {
#pragma unroll
        for (int j = 0; j < (kPrefetch ? CH : 0); ++j) {
            const int c = (j * 32 + lane) * 8;
            ndy[j] = __ldg(reinterpret_cast<const uint4*>(dy + (size_t)r * N + c));
            nx[j] = __ldg(reinterpret_cast<const uint4*>(x + (size_t)r * N + c));
        }
        nmean = __ldg(mean_in + r), nrstd = __ldg(rstd_in + r);
    };
    auto unpack8 = [](const uint4& u, float (&v)[8]) {
        float2 f;
        f = unpack_bf16x2(u.x), v[0] = f.x, v[1] = f.y;
        f = unpack_bf16x2(u.y), v[2] = f.x, v[3] = f.y;
        f = unpack_bf16x2(u.z), v[4] = f.x, v[5] = f.y;
        f = unpack_bf16x2(u.w), v[6] = f.x, v[7] = f.y;
    };
    const int row_first = row_begin + blockIdx.x * kRowWarps + warp, row_step = gridDim.x * kRowWarps;
    if (kPrefetch && row_first < rows) fetch(row_first);
    for (int row = row_first; row < rows; row += row_step) {
        float mean, rstd;
        float g[CH][8], xh[CH][8];
        if (kPrefetch) {
            mean = nmean, rstd = nrstd;
#pragma unroll
            for (int j = 0; j < CH; ++j) unpack8(ndy[j], g[j]), unpack8(nx[j], xh[j]);
            if (row + row_step < rows) fetch(row + row_step);
        } else {
            mean = mean_in[row], rstd = rstd_in[row];
        }
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int j = 0; j < CH; ++j) {
            const int c = (j * 32 + lane) * 8;
            if (!kPrefetch) load8(dy + (size_t)row * N + c, g[j]);
            if (dy2) {
                float t[8];
                load8(dy2 + (size_t)row * N + c, t);
#pragma unroll
                for (int i = 0; i < 8; ++i) g[j][i] += t[i];
            }
            if (in_drop.p > 0.f) drop8(g[j], ikey, ith, iinv, (uint64_t)row * N + c);   // y = dropout(LN(x))
            if (!kPrefetch) load8(x + (size_t)row * N + c, xh[j]);
            if (residual) {
                float t[8];
                load8(residual + (size_t)row * N + c, t);
#pragma unroll
                for (int i = 0; i < 8; ++i) xh[j][i] += t[i];
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                xh[j][i] = (xh[j][i] - mean) * rstd;
                const float gg = g[j][i] * gm[j][i];
                s1 += gg;
                s2 += gg * xh[j][i];
                ag[j][i] += g[j][i] * xh[j][i];
                ab[j][i] += g[j][i];
            }
        }
        s1 = warp_sum(s1) * (1.0f / N);
        s2 = warp_sum(s2) * (1.0f / N);
#pragma unroll
        for (int j = 0; j < CH; ++j) {
            float o[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) o[i] = rstd * (g[j][i] * gm[j][i] - s1 - xh[j][i] * s2);
            store8(dx + (size_t)row * N + (j * 32 + lane) * 8, o);
            if (out_drop.p > 0.f) {
                // x = dropout(dense_out) + residual: the projection sees the masked, rescaled gradient
                drop8(o, okey, oth, oinv, (uint64_t)row * N + (j * 32 + lane) * 8);
                store8(dx_drop + (size_t)row * N + (j * 32 + lane) * 8, o);
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) ax[j][i] += o[i];
        }
    }
    // block reduction of the three column accumulators (one 256-column chunk at a time), then one atomic per
    // column per block
    for (int pass = 0; pass < 3; ++pass) {
        float* dst = pass == 0 ? dgamma : (pass == 1 ? dbeta : dxsum);
        if (!dst) continue;  // uniform
#pragma unroll
        for (int j = 0; j < CH; ++j) {
#pragma unroll
            for (int i = 0; i < 8; ++i)
                red[warp][lane * 8 + i] = pass == 0 ? ag[j][i] : (pass == 1 ? ab[j][i] : ax[j][i]);
            __syncthreads();
            float sum = 0.f;
#pragma unroll
            for (int w = 0; w < kRowWarps; ++w) sum += red[w][threadIdx.x];
            atomicAdd(dst + j * 256 + threadIdx.x, sum);
            __syncthreads();
        }
    }
}

// out[c] += sum_rows x[row, c]  (bias gradients).  Block = 256 threads over 8-column groups x row slices.
__global__ void __launch_bounds__(256) colsum_kernel(const __nv_bfloat16* __restrict__ x, int64_t ld, int rows, int n,
                                                    float* __restrict__ out, int rows_per_block) {
    // thread t owns column group (blockIdx.x*32 + t%32) of 8 columns, row phase t/32; a warp reads 512 contiguous
    // bytes of a row.  Four rows per iteration, all loads issued before the adds: the kernel is a pure streaming
    // read and lives on memory-level parallelism (one outstanding 16-byte load per thread measured 0.16-0.22 of HBM).
    const int cg = blockIdx.x * 32 + (threadIdx.x & 31);
    const int c = cg * 8;
    const int r0 = blockIdx.y * rows_per_block;
    const int r1 = min(r0 + rows_per_block, rows);
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    auto add8 = [&](const uint4& q) {
        float2 f;
        f = unpack_bf16x2(q.x), acc[0] += f.x, acc[1] += f.y;
        f = unpack_bf16x2(q.y), acc[2] += f.x, acc[3] += f.y;
        f = unpack_bf16x2(q.z), acc[4] += f.x, acc[5] += f.y;
        f = unpack_bf16x2(q.w), acc[6] += f.x, acc[7] += f.y;
    };
    if (c < n) {
        const __nv_bfloat16* col = x + c;
        int r = r0 + (threadIdx.x >> 5);
        for (; r + 24 < r1; r += 32) {
            const uint4 q0 = __ldg(reinterpret_cast<const uint4*>(col + (size_t)r * ld));
            const uint4 q1 = __ldg(reinterpret_cast<const uint4*>(col + (size_t)(r + 8) * ld));
            const uint4 q2 = __ldg(reinterpret_cast<const uint4*>(col + (size_t)(r + 16) * ld));
            const uint4 q3 = __ldg(reinterpret_cast<const uint4*>(col + (size_t)(r + 24) * ld));
            add8(q0), add8(q1), add8(q2), add8(q3);
        }
        for (; r < r1; r += 8) add8(__ldg(reinterpret_cast<const uint4*>(col + (size_t)r * ld)));
    }
    __shared__ float red[8][32][9];
#pragma unroll
    for (int i = 0; i < 8; ++i) red[threadIdx.x >> 5][threadIdx.x & 31][i] = acc[i];
    __syncthreads();
    if (threadIdx.x < 32 && c < n) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float s = 0.f;
#pragma unroll
            for (int w = 0; w < 8; ++w) s += red[w][threadIdx.x][i];
            if (c + i < n) atomicAdd(out + c + i, s);
        }
    }
}

// Embedding gathers fused with LayerNorm (BertEmbeddings, lxrt/modeling_capsbert.py:341-355, and
// HGEmbeddings, :313-325).  word_ids == NULL means "row r of the table" (HGEmbeddings uses every row).
template <int CH>
__global__ void __launch_bounds__(kRowThreads)
embed_ln_fwd_kernel(const float* __restrict__ word, const int64_t* __restrict__ word_ids,
                    const float* __restrict__ pos, const float* __restrict__ type,
                    const int64_t* __restrict__ type_ids, const float* __restrict__ gamma,
                    const float* __restrict__ beta, float eps, __nv_bfloat16* __restrict__ y,
                    __nv_bfloat16* __restrict__ pre_out, float* __restrict__ mean_out, float* __restrict__ rstd_out,
                    int rows, int seq_len, DropCfg drop) {
    constexpr int N = 256 * CH;
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * kRowWarps + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int t = row % seq_len;
    const int64_t wid = word_ids ? word_ids[row] : t;
    const int64_t tid = type_ids ? type_ids[row] : 0;
    float v[CH][8];
#pragma unroll
    for (int j = 0; j < CH; ++j) {
        const int c = (j * 32 + lane) * 8;
        load8(word + (size_t)wid * N + c, v[j]);
        float a[8];
        if (pos) {
            load8(pos + (size_t)t * N + c, a);
#pragma unroll
            for (int i = 0; i < 8; ++i) v[j][i] += a[i];
        }
        if (type) {
            load8(type + (size_t)tid * N + c, a);
#pragma unroll
            for (int i = 0; i < 8; ++i) v[j][i] += a[i];
        }
    }
    if (pre_out) {
        // the backward re-reads the pre-LN sum in bf16, so the statistics are taken on the rounded values
#pragma unroll
        for (int j = 0; j < CH; ++j) {
            store8(pre_out + (size_t)row * N + (j * 32 + lane) * 8, v[j]);
#pragma unroll
            for (int i = 0; i < 8; ++i) v[j][i] = __bfloat162float(__float2bfloat16(v[j][i]));
        }
    }
    float mean, rstd;
    row_stats<CH>(v, 1.0f / N, eps, mean, rstd);
    if (rstd_out && lane == 0) mean_out[row] = mean, rstd_out[row] = rstd;
#pragma unroll
    for (int j = 0; j < CH; ++j) {
        const int c = (j * 32 + lane) * 8;
        float g[8], b[8], o[8];
        load8(gamma + c, g);
        load8(beta + c, b);
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = (v[j][i] - mean) * rstd * g[i] + b[i];
        if (drop.p > 0.f)
            drop8(o, drop_key(__ldg(drop.seed), drop.salt), drop_thresh(drop.p), 1.f / (1.f - drop.p),
                  (uint64_t)row * N + c);
        store8(y + (size_t)row * N + c, o);
    }
}

// Backward of the gathers: scatter-add of the (already LayerNorm-back-propagated) row gradient into the
// tables with fp32 atomics.  padding_idx rows (< 0 = none) receive no gradient, as nn.Embedding's do.
template <int CH>
__global__ void __launch_bounds__(kRowThreads)
embed_scatter_bwd_kernel(const __nv_bfloat16* __restrict__ de, const int64_t* __restrict__ word_ids,
                         const int64_t* __restrict__ type_ids, float* __restrict__ dword, int word_pad,
                         float* __restrict__ dpos, int pos_pad, float* __restrict__ dtype, int type_pad, int rows,
                         int seq_len) {
    constexpr int N = 256 * CH;
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * kRowWarps + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int t = row % seq_len;
    const int64_t wid = word_ids ? word_ids[row] : t;
    const int64_t tid = type_ids ? type_ids[row] : 0;
#pragma unroll
    for (int j = 0; j < CH; ++j) {
        const int c = (j * 32 + lane) * 8;
        float d[8];
        load8(de + (size_t)row * N + c, d);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (dword && wid != word_pad) atomicAdd(dword + (size_t)wid * N + c + i, d[i]);
            if (dpos && t != pos_pad) atomicAdd(dpos + (size_t)t * N + c + i, d[i]);
            if (dtype && tid != type_pad) atomicAdd(dtype + (size_t)tid * N + c + i, d[i]);
        }
    }
}

// y[r, :] = x[r, :] + add[r % add_rows, :]   (type tokens / position rows; bf16 in, bf16 out)
__global__ void __launch_bounds__(256) add_rows_kernel(const __nv_bfloat16* __restrict__ x,
                                                      const __nv_bfloat16* __restrict__ add, int add_rows,
                                                      __nv_bfloat16* __restrict__ y, int64_t total8, int n8) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total8; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / n8;
        const int c8 = (int)(i - r * n8);
        float a[8], b[8];
        load8(x + i * 8, a);
        load8(add + ((r % add_rows) * n8 + c8) * 8, b);
#pragma unroll
        for (int k = 0; k < 8; ++k) a[k] += b[k];
        store8(y + i * 8, a);
    }
}

// dst[i0*d0 + i1*d1 + e] = src[i0*s0 + i1*s1 + e] (+ add[e]) for e < chunk: strided block gather/scatter used to
// assemble the hypergraph token sequence (tasks/agqa_model.py:255-258, lxrt/modeling_capsbert.py:1168-1184),
// to pick [CLS] rows for the pooler, and to route the matching gradients back.  chunk % 8 == 0.
__global__ void __launch_bounds__(256) copy3d_kernel(const __nv_bfloat16* __restrict__ src,
                                                    __nv_bfloat16* __restrict__ dst,
                                                    const __nv_bfloat16* __restrict__ add, int n0, int n1, int chunk8,
                                                    int64_t s0, int64_t s1, int64_t d0, int64_t d1, int accumulate) {
    const int64_t total = (int64_t)n0 * n1 * chunk8;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int e8 = (int)(i % chunk8);
        const int64_t blk = i / chunk8;
        const int i1 = (int)(blk % n1);
        const int64_t i0 = blk / n1;
        float v[8];
        load8(src + i0 * s0 + i1 * s1 + e8 * 8, v);
        if (add) {
            float a[8];
            load8(add + e8 * 8, a);
#pragma unroll
            for (int k = 0; k < 8; ++k) v[k] += a[k];
        }
        __nv_bfloat16* o = dst + i0 * d0 + i1 * d1 + e8 * 8;
        if (accumulate) {
            float a[8];
            load8(o, a);
#pragma unroll
            for (int k = 0; k < 8; ++k) v[k] += a[k];
        }
        store8(o, v);
    }
}

// dx = dy * act'(aux): GELU from the saved pre-activation, tanh / ReLU from the saved output.
__global__ void __launch_bounds__(256) act_bwd_kernel(const __nv_bfloat16* __restrict__ dy,
                                                     const __nv_bfloat16* __restrict__ aux,
                                                     __nv_bfloat16* __restrict__ dx, int64_t n8, int mode) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n8; i += (int64_t)gridDim.x * blockDim.x) {
        float g[8], a[8];
        load8(dy + i * 8, g);
        load8(aux + i * 8, a);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if (mode == SHG_ACT_GELU_ERF) g[k] *= gelu_erf_grad(a[k]);
            else if (mode == SHG_ACT_TANH) g[k] *= 1.0f - a[k] * a[k];
            else g[k] = a[k] > 0.f ? g[k] : 0.f;
        }
        store8(dx + i * 8, g);
    }
}

// BCEWithLogitsLoss()(x, t) * n_cols (tasks/agqaHGQA.py:344-345) = (1/rows) * sum over all elements; and its
// gradient (sigmoid(x) - t) * grad_scale / rows as bf16.  out[0] += loss.
__global__ void __launch_bounds__(256) bce_logits_kernel(const float* __restrict__ x, int64_t ldx,
                                                        const float* __restrict__ t, int rows, int cols,
                                                        float grad_scale, float* __restrict__ out,
                                                        __nv_bfloat16* __restrict__ dx, int64_t lddx) {
    float acc = 0.f;
    const int64_t total = (int64_t)rows * cols;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int r = (int)(i / cols), c = (int)(i - (int64_t)r * cols);
        const float v = x[r * ldx + c], tv = t[i];
        acc += fmaxf(v, 0.f) - v * tv + log1pf(expf(-fabsf(v)));
        if (dx) dx[r * lddx + c] = __float2bfloat16((1.0f / (1.0f + expf(-v)) - tv) * grad_scale / rows);
    }
    acc = warp_sum(acc);
    __shared__ float red[8];
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
        for (int w = 0; w < 8; ++w) s += red[w];
        atomicAdd(out, s / rows);
    }
}

__global__ void __launch_bounds__(256) f32_to_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y,
                                                         int64_t n) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t n8 = n / 8;
    for (; i < n8; i += stride) {
        float v[8];
        load8(x + i * 8, v);
        store8(y + i * 8, v);
    }
    for (int64_t t = n8 * 8 + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += stride)
        y[t] = __float2bfloat16(x[t]);
}

}  // namespace shg

extern "C" {

int shg_layernorm_fwd_grouped(const void* x, const void* residual, const float* gamma, const float* beta, float eps,
                              void* y, void* y2, const void* addend, int addend_rows, float* mean, float* rstd, int rows,
                              int n, int split_row, const float* gamma2, const float* beta2, void* stream) {
    using namespace shg;
    SHG_REQUIRE(n == 768 || n == 1536, SHG_ESHAPE, "shg_layernorm_fwd: width %d (768 or 1536 supported)", n);
    SHG_REQUIRE((y2 == nullptr) == (addend == nullptr), SHG_EINVAL, "shg_layernorm_fwd: y2 and addend go together");
    SHG_REQUIRE((mean == nullptr) == (rstd == nullptr), SHG_EINVAL, "shg_layernorm_fwd: mean and rstd go together");
    SHG_REQUIRE(split_row >= rows || (gamma2 && beta2 && split_row > 0), SHG_EINVAL,
                "shg_layernorm_fwd_grouped: second parameter set missing");
    if (rows <= 0) return SHG_OK;
    const int blocks = (rows + kRowWarps - 1) / kRowWarps;
    cudaStream_t s = (cudaStream_t)stream;
    if (n == 768)
        shg::launch(layernorm_fwd_kernel<3>, blocks, kRowThreads, 0, s,
            (const __nv_bfloat16*)x, (const __nv_bfloat16*)residual, gamma, beta, eps, (__nv_bfloat16*)y,
            (__nv_bfloat16*)y2, (const __nv_bfloat16*)addend, addend_rows, mean, rstd, rows, split_row, gamma2, beta2);
    else
        shg::launch(layernorm_fwd_kernel<6>, blocks, kRowThreads, 0, s,
            (const __nv_bfloat16*)x, (const __nv_bfloat16*)residual, gamma, beta, eps, (__nv_bfloat16*)y,
            (__nv_bfloat16*)y2, (const __nv_bfloat16*)addend, addend_rows, mean, rstd, rows, split_row, gamma2, beta2);
    return check_launch("shg_layernorm_fwd");
}

int shg_layernorm_fwd(const void* x, const void* residual, const float* gamma, const float* beta, float eps, void* y,
                      void* y2, const void* addend, int addend_rows, float* mean, float* rstd, int rows, int n,
                      void* stream) {
    return shg_layernorm_fwd_grouped(x, residual, gamma, beta, eps, y, y2, addend, addend_rows, mean, rstd, rows, n,
                                     0x7fffffff, nullptr, nullptr, stream);
}

int shg_layernorm_bwd_grouped(const void* dy, const void* dy2, const void* x, const void* residual, const float* gamma,
                              const float* mean, const float* rstd, void* dx, float* dgamma, float* dbeta, float* dxsum,
                              int rows, int n, void* dx_drop, float out_p, uint32_t out_salt, float in_p,
                              uint32_t in_salt, const uint32_t* seed, int split_row, const float* gamma2,
                              float* dgamma2, float* dbeta2, float* dxsum2, void* stream) {
    using namespace shg;
    DropCfg od{(seed && dx_drop) ? out_p : 0.f, out_salt, seed}, id{seed ? in_p : 0.f, in_salt, seed};
    SHG_REQUIRE(n == 768 || n == 1536, SHG_ESHAPE, "shg_layernorm_bwd: width %d (768 or 1536 supported)", n);
    if (rows <= 0) return SHG_OK;
    const bool grouped = split_row < rows;
    SHG_REQUIRE(!grouped || (split_row > 0 && gamma2 && dgamma2 && dbeta2 && (dxsum2 != nullptr) == (dxsum != nullptr)),
                SHG_EINVAL, "shg_layernorm_bwd_grouped: second parameter / gradient set missing");
    const int rows_max = grouped ? (split_row > rows - split_row ? split_row : rows - split_row) : rows;
    int blocks = (rows_max + kRowWarps - 1) / kRowWarps;
    const int cap = grouped ? kNumSMs : 2 * kNumSMs;
    if (blocks > cap) blocks = cap;
    dim3 grid(blocks, grouped ? 2 : 1);
    cudaStream_t s = (cudaStream_t)stream;
    if (n == 768)
        shg::launch(layernorm_bwd_kernel<3>, grid, kRowThreads, 0, s,
            (const __nv_bfloat16*)dy, (const __nv_bfloat16*)dy2, (const __nv_bfloat16*)x,
            (const __nv_bfloat16*)residual, gamma, mean, rstd, (__nv_bfloat16*)dx, dgamma, dbeta, dxsum, rows,
            (__nv_bfloat16*)dx_drop, od, id, split_row, gamma2, dgamma2, dbeta2, dxsum2);
    else
        shg::launch(layernorm_bwd_kernel<6>, grid, kRowThreads, 0, s,
            (const __nv_bfloat16*)dy, (const __nv_bfloat16*)dy2, (const __nv_bfloat16*)x,
            (const __nv_bfloat16*)residual, gamma, mean, rstd, (__nv_bfloat16*)dx, dgamma, dbeta, dxsum, rows,
            (__nv_bfloat16*)dx_drop, od, id, split_row, gamma2, dgamma2, dbeta2, dxsum2);
    return check_launch("shg_layernorm_bwd");
}

int shg_layernorm_bwd(const void* dy, const void* dy2, const void* x, const void* residual, const float* gamma,
                      const float* mean, const float* rstd, void* dx, float* dgamma, float* dbeta, float* dxsum,
                      int rows, int n, void* dx_drop, float out_p, uint32_t out_salt, float in_p, uint32_t in_salt,
                      const uint32_t* seed, void* stream) {
    return shg_layernorm_bwd_grouped(dy, dy2, x, residual, gamma, mean, rstd, dx, dgamma, dbeta, dxsum, rows, n, dx_drop,
                                     out_p, out_salt, in_p, in_salt, seed, 0x7fffffff, nullptr, nullptr, nullptr,
                                     nullptr, stream);
}

int shg_colsum_accum(const void* x, int64_t ld, int rows, int n, float* out, void* stream) {
    using namespace shg;
    SHG_REQUIRE(ld % 8 == 0 && ld >= (n + 7) / 8 * 8, SHG_ESHAPE,
                "shg_colsum_accum: ld=%lld must be a multiple of 8 covering n=%d rounded up to 8", (long long)ld, n);
    if (rows <= 0) return SHG_OK;
    const int col_blocks = ((n + 7) / 8 + 31) / 32;
    int row_blocks = (4 * kNumSMs + col_blocks - 1) / col_blocks;
    int rpb = (rows + row_blocks - 1) / row_blocks;
    if (rpb < 64) rpb = 64;
    row_blocks = (rows + rpb - 1) / rpb;
    shg::launch(colsum_kernel, dim3(col_blocks, row_blocks), 256, 0, (cudaStream_t)stream, (const __nv_bfloat16*)x, ld, rows,
                                                                                  n, out, rpb);
    return check_launch("shg_colsum_accum");
}

int shg_embed_ln_fwd(const float* word, const int64_t* word_ids, const float* pos, const float* type,
                     const int64_t* type_ids, const float* gamma, const float* beta, float eps, void* y, void* pre,
                     float* mean, float* rstd, int rows, int seq_len, int n, float drop_p, uint32_t drop_salt,
                     const uint32_t* seed, void* stream) {
    using namespace shg;
    SHG_REQUIRE(n == 768, SHG_ESHAPE, "shg_embed_ln_fwd: width %d (768 supported)", n);
    SHG_REQUIRE((mean == nullptr) == (rstd == nullptr), SHG_EINVAL, "shg_embed_ln_fwd: mean and rstd go together");
    if (rows <= 0) return SHG_OK;
    const int blocks = (rows + kRowWarps - 1) / kRowWarps;
    shg::launch(embed_ln_fwd_kernel<3>, blocks, kRowThreads, 0, (cudaStream_t)stream, 
        word, word_ids, pos, type, type_ids, gamma, beta, eps, (__nv_bfloat16*)y, (__nv_bfloat16*)pre, mean, rstd,
        rows, seq_len, DropCfg{seed ? drop_p : 0.f, drop_salt, seed});
    return check_launch("shg_embed_ln_fwd");
}

int shg_embed_scatter_bwd(const void* de, const int64_t* word_ids, const int64_t* type_ids, float* dword,
                          int word_pad, float* dpos, int pos_pad, float* dtype, int type_pad, int rows, int seq_len,
                          int n, void* stream) {
    using namespace shg;
    SHG_REQUIRE(n == 768, SHG_ESHAPE, "shg_embed_scatter_bwd: width %d (768 supported)", n);
    if (rows <= 0) return SHG_OK;
    const int blocks = (rows + kRowWarps - 1) / kRowWarps;
    shg::launch(embed_scatter_bwd_kernel<3>, blocks, kRowThreads, 0, (cudaStream_t)stream, 
        (const __nv_bfloat16*)de, word_ids, type_ids, dword, word_pad, dpos, pos_pad, dtype, type_pad, rows, seq_len);
    return check_launch("shg_embed_scatter_bwd");
}

int shg_add_rows(const void* x, const void* add, int add_rows, void* y, int rows, int n, void* stream) {
    using namespace shg;
    SHG_REQUIRE(n % 8 == 0 && add_rows > 0, SHG_ESHAPE, "shg_add_rows: n=%d add_rows=%d", n, add_rows);
    if (rows <= 0) return SHG_OK;
    const int64_t total8 = (int64_t)rows * (n / 8);
    int blocks = (int)((total8 + 255) / 256);
    if (blocks > 8 * kNumSMs) blocks = 8 * kNumSMs;
    shg::launch(add_rows_kernel, blocks, 256, 0, (cudaStream_t)stream, (const __nv_bfloat16*)x, (const __nv_bfloat16*)add,
                                                              add_rows, (__nv_bfloat16*)y, total8, n / 8);
    return check_launch("shg_add_rows");
}

int shg_copy3d(const void* src, void* dst, const void* add, int n0, int n1, int chunk, int64_t s0, int64_t s1,
               int64_t d0, int64_t d1, int accumulate, void* stream) {
    using namespace shg;
    SHG_REQUIRE(chunk % 8 == 0 && s0 % 8 == 0 && s1 % 8 == 0 && d0 % 8 == 0 && d1 % 8 == 0, SHG_ESHAPE,
                "shg_copy3d: chunk and strides must be multiples of 8 elements");
    if (n0 <= 0 || n1 <= 0 || chunk <= 0) return SHG_OK;
    const int64_t total = (int64_t)n0 * n1 * (chunk / 8);
    int64_t blocks = (total + 255) / 256;
    if (blocks > 8 * kNumSMs) blocks = 8 * kNumSMs;
    shg::launch(copy3d_kernel, (int)blocks, 256, 0, (cudaStream_t)stream, (const __nv_bfloat16*)src, (__nv_bfloat16*)dst,
                                                                 (const __nv_bfloat16*)add, n0, n1, chunk / 8, s0, s1,
                                                                 d0, d1, accumulate);
    return check_launch("shg_copy3d");
}

int shg_act_bwd(const void* dy, const void* aux, void* dx, int64_t n, int act, void* stream) {
    using namespace shg;
    SHG_REQUIRE(n % 8 == 0, SHG_ESHAPE, "shg_act_bwd: n=%lld must be a multiple of 8", (long long)n);
    SHG_REQUIRE(act >= SHG_ACT_GELU_ERF && act <= SHG_ACT_TANH, SHG_EINVAL, "shg_act_bwd: act=%d", act);
    if (n <= 0) return SHG_OK;
    int64_t blocks = (n / 8 + 255) / 256;
    if (blocks > 8 * kNumSMs) blocks = 8 * kNumSMs;
    shg::launch(act_bwd_kernel, (int)blocks, 256, 0, (cudaStream_t)stream, (const __nv_bfloat16*)dy, (const __nv_bfloat16*)aux,
                                                                  (__nv_bfloat16*)dx, n / 8, act);
    return check_launch("shg_act_bwd");
}

int shg_bce_logits(const float* x, int64_t ldx, const float* target, int rows, int cols, float grad_scale, float* out,
                   void* dx, int64_t lddx, void* stream) {
    using namespace shg;
    if (rows <= 0 || cols <= 0) return SHG_OK;
    int blocks = (int)(((int64_t)rows * cols + 255) / 256);
    if (blocks > kNumSMs) blocks = kNumSMs;
    shg::launch(bce_logits_kernel, blocks, 256, 0, (cudaStream_t)stream, x, ldx, target, rows, cols, grad_scale, out,
                                                                (__nv_bfloat16*)dx, lddx);
    return check_launch("shg_bce_logits");
}

int shg_f32_to_bf16(const float* x, void* y, int64_t n, void* stream) {
    using namespace shg;
    if (n <= 0) return SHG_OK;
    SHG_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(y) & 15) == 0, SHG_ESHAPE,
                "shg_f32_to_bf16: pointers must be 16-byte aligned");
    int64_t blocks = (n / 8 + 255) / 256;
    if (blocks > 8 * kNumSMs) blocks = 8 * kNumSMs;
    if (blocks < 1) blocks = 1;
    shg::launch(f32_to_bf16_kernel, (int)blocks, 256, 0, (cudaStream_t)stream, x, (__nv_bfloat16*)y, n);
    return check_launch("shg_f32_to_bf16");
}
}
