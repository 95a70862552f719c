// Batched per-frame Hungarian matching and the set-prediction loss, on device.
//
// Replaces the CPU path of the reference's HungarianMatcher (lxrt/matcher.py:59-80: softmax -> gather ->
// .cpu() -> 2*B*16 scipy.optimize.linear_sum_assignment calls) and the scatter / weighted-CE glue of
// AGQA.loss_labels (tasks/agqaHGQA.py:203-229).
//
// Solver: scipy's rectangular_lsap (shortest augmenting path) with the exact tie-break of its column scan.
// One CTA per frame; the cost block is staged in shared memory; ONE WARP walks the (inherently sequential)
// augmentations with lane j owning column j's state (v, shortest-path cost, path, row4col, position in
// scipy's `remaining` list) and lane i owning row i's (u, col4row).  Every arithmetic step is the same
// fp64 expression, in the same association order, as the CPU code, so assignments are bit-identical on
// identical cost matrices.
#include <math.h>

#include "shg_common.cuh"

namespace shg {

constexpr int kMaxDim = 32;

__device__ __forceinline__ double warp_min_f64(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ int warp_max_i32(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// sC: working matrix, nr <= nc, element (i, j) at sC[i * kMaxDim + j].  Called by one full warp.
// Returns 0 or -1 (infeasible).  On return lane l holds col4row (row l) and row4col (column l).
__device__ int solve_warp(const float* sC, int nr, int nc, int lane, int& col4row_out, int& row4col_out) {
    const double kInf = INFINITY;
    double u = 0.0, v = 0.0;
    int col4row = -1, row4col = -1, path = -1;
    for (int cur = 0; cur < nr; ++cur) {
        // scipy: remaining[it] = nc - it - 1  =>  column j sits at position nc-1-j
        int pos = nc - 1 - lane;
        bool in_rem = lane < nc;
        int n_rem = nc;
        bool SR = false, SC = false;
        double spc = kInf, min_val = 0.0;
        int i = cur, sink = -1;
        while (sink == -1) {
            if (lane == i) SR = true;
            const double u_i = __shfl_sync(0xffffffffu, u, i);
            if (in_rem) {
                const double r = ((min_val + (double)sC[i * kMaxDim + lane]) - u_i) - v;
                if (r < spc) {
                    path = i;
                    spc = r;
                }
            }
            const double lowest = warp_min_f64(in_rem ? spc : kInf);
            if (lowest == kInf) return -1;
            // scan-order tie-break: the LAST unassigned column at the minimum wins, else the FIRST column
            int key = -1;
            if (in_rem && spc == lowest) key = (row4col == -1) ? (64 + pos) : (63 - pos);
            const int best = warp_max_i32(key);
            const unsigned who = __ballot_sync(0xffffffffu, key == best);
            const int j = __ffs(who) - 1;
            min_val = lowest;
            const int r4c_j = __shfl_sync(0xffffffffu, row4col, j);
            const int pos_j = __shfl_sync(0xffffffffu, pos, j);
            if (r4c_j == -1) sink = j;
            else i = r4c_j;
            if (lane == j) {
                SC = true;
                in_rem = false;
            }
            // remaining[index] = remaining[--n_rem]
            --n_rem;
            if (in_rem && pos == n_rem) pos = pos_j;
        }
        const double spc_of_my_col = __shfl_sync(0xffffffffu, spc, col4row >= 0 ? col4row : 0);
        if (lane == cur) u += min_val;
        else if (SR) u += min_val - spc_of_my_col;
        if (SC) v -= min_val - spc;
        int j = sink;
        for (;;) {
            const int pi = __shfl_sync(0xffffffffu, path, j);
            if (lane == j) row4col = pi;
            const int t = __shfl_sync(0xffffffffu, col4row, pi);
            if (lane == pi) col4row = j;
            j = t;
            if (pi == cur) break;
        }
    }
    col4row_out = col4row;
    row4col_out = row4col;
    return 0;
}

// Writes the frame's result. `transposed`: working rows are targets, working columns are queries.
__device__ void emit_warp(int rc, bool transposed, int nr, int nc, int q_f, int lane, int col4row, int row4col,
                          int32_t* rows, int32_t* cols, int32_t* n_match, int32_t* status, const int32_t* labels,
                          int32_t* target_classes) {
    for (int t = lane; t < q_f; t += 32) {
        rows[t] = -1;
        cols[t] = -1;
        if (target_classes) target_classes[t] = 0;
    }
    __syncwarp();
    if (rc != 0) {
        if (lane == 0) {
            *n_match = 0;
            *status = rc;
        }
        return;
    }
    if (transposed) {
        const bool has = lane < nc && row4col != -1;
        const unsigned m = __ballot_sync(0xffffffffu, has);
        if (has) {
            const int rank = __popc(m & ((1u << lane) - 1));
            rows[rank] = lane;
            cols[rank] = row4col;
            if (target_classes) target_classes[lane] = labels[row4col];
        }
    } else if (lane < nr) {
        rows[lane] = lane;
        cols[lane] = col4row;
        if (target_classes) target_classes[lane] = labels[col4row];
    }
    if (lane == 0) {
        *n_match = nr;
        *status = 0;
    }
}

__global__ void __launch_bounds__(32) lsap_frames_kernel(const float* __restrict__ cost, int q_f, int ld,
                                                        const int32_t* __restrict__ col_off, int32_t* rows,
                                                        int32_t* cols, int32_t* n_match, int32_t* status) {
    __shared__ float sC[kMaxDim * kMaxDim];
    const int f = blockIdx.x, lane = threadIdx.x;
    const int c0 = col_off[f], n = col_off[f + 1] - c0;
    const float* src = cost + (size_t)f * q_f * ld + c0;
    const bool transposed = n < q_f;
    const int nr = transposed ? n : q_f, nc = transposed ? q_f : n;
    bool bad = false;
    // coalesced along the target columns; stored as the working (possibly transposed) matrix
    for (int e = lane; e < q_f * n; e += 32) {
        const int q = e / n, t = e - q * n;
        const float c = src[(size_t)q * ld + t];
        bad |= isnan(c) || c == -INFINITY;
        if (transposed) sC[t * kMaxDim + q] = c;
        else sC[q * kMaxDim + t] = c;
    }
    bad = __any_sync(0xffffffffu, bad);
    __syncwarp();
    int col4row = -1, row4col = -1, rc = 0;
    if (bad) rc = -2;
    else if (n > kMaxDim) rc = -3;
    else if (nr > 0) rc = solve_warp(sC, nr, nc, lane, col4row, row4col);
    emit_warp(rc, transposed, nr, nc, q_f, lane, col4row, row4col, rows + (size_t)f * q_f, cols + (size_t)f * q_f,
              n_match + f, status + f, nullptr, nullptr);
}

template <typename T>
__device__ __forceinline__ float ldf(const T* p);
template <>
__device__ __forceinline__ float ldf<float>(const float* p) { return __ldg(p); }
template <>
__device__ __forceinline__ float ldf<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }

// Fused: per query row softmax statistics (one warp per row, coalesced over classes), cost = -p[label],
// then warp 0 solves the frame and scatters the matched labels.
template <typename T>
__global__ void __launch_bounds__(256) match_from_logits_kernel(
    const T* __restrict__ logits, int q_f, int n_cls, int ld, const int32_t* __restrict__ labels,
    const int32_t* __restrict__ label_off, int32_t* rows, int32_t* cols, int32_t* n_match, int32_t* status,
    float* row_max, float* row_lse, int32_t* row_argmax, int32_t* target_classes) {
    __shared__ float sC[kMaxDim * kMaxDim];
    __shared__ int s_bad;
    const int f = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
    const int l0 = label_off[f], n = label_off[f + 1] - l0;
    const bool transposed = n < q_f;
    if (threadIdx.x == 0) s_bad = 0;
    __syncthreads();
    for (int q = warp; q < q_f; q += n_warps) {
        const T* x = logits + ((size_t)f * q_f + q) * ld;
        float mx = -INFINITY;
        int arg = 0;
        for (int c = lane; c < n_cls; c += 32) {
            const float v = ldf(x + c);
            if (v > mx) {
                mx = v;
                arg = c;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float om = __shfl_xor_sync(0xffffffffu, mx, o);
            const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
            if (om > mx || (om == mx && oa < arg)) {
                mx = om;
                arg = oa;
            }
        }
        float sum = 0.f;
        for (int c = lane; c < n_cls; c += 32) sum += expf(ldf(x + c) - mx);
        sum = warp_sum(sum);
        if (lane == 0) {
            row_max[(size_t)f * q_f + q] = mx;
            row_lse[(size_t)f * q_f + q] = mx + logf(sum);
            if (row_argmax) row_argmax[(size_t)f * q_f + q] = arg;
        }
        if (lane < n && n <= kMaxDim) {
            const int lab = labels[l0 + lane];
            // a label outside [0, n_cls) is an IndexError in the reference (prob[:, tgt_ids], lxrt/matcher.py:73):
            // never dereferenced here, the frame is flagged (status -4) and left unmatched
            const bool lab_ok = lab >= 0 && lab < n_cls;
            const float c = lab_ok ? -(expf(ldf(x + lab) - mx) / sum) : 0.f;
            if (!lab_ok) atomicMax(&s_bad, 2);
            else if (isnan(c) || c == -INFINITY) atomicMax(&s_bad, 1);
            if (transposed) sC[lane * kMaxDim + q] = c;
            else sC[q * kMaxDim + lane] = c;
        }
    }
    __syncthreads();
    if (warp != 0) return;
    const int nr = transposed ? n : q_f, nc = transposed ? q_f : n;
    int col4row = -1, row4col = -1, rc = 0;
    if (s_bad == 2) rc = -4;
    else if (s_bad) rc = -2;
    else if (n > kMaxDim) rc = -3;
    else if (nr > 0) rc = solve_warp(sC, nr, nc, lane, col4row, row4col);
    emit_warp(rc, transposed, nr, nc, q_f, lane, col4row, row4col, rows + (size_t)f * q_f, cols + (size_t)f * q_f,
              n_match + f, status + f, labels + l0, target_classes ? target_classes + (size_t)f * q_f : nullptr);
}

// out[0] += sum w*nll, out[1] += sum w, out[2] += #matched, out[3] += #matched with argmax == label
template <typename T>
__global__ void __launch_bounds__(1024) set_loss_fwd_kernel(const T* __restrict__ logits, int n_rows, int ld,
                                                          const int32_t* __restrict__ tgt,
                                                          const float* __restrict__ row_lse,
                                                          const int32_t* __restrict__ row_argmax, float eos_coef,
                                                          float* out, int n_cls, const int32_t* __restrict__ status,
                                                          int n_frames, float* err_out) {
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    int n_err = 0;
    // frames the matcher refused (NaN / -inf costs, too many targets, label out of range: the reference raises
    // there) and target classes outside [0, n_cls) are counted into *err_out, which the host turns into an exception
    // when it reads the loss sums — no extra synchronisation
    if (status)
        for (int f = blockIdx.x * blockDim.x + threadIdx.x; f < n_frames; f += gridDim.x * blockDim.x)
            n_err += status[f] != 0;
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < n_rows; r += gridDim.x * blockDim.x) {
        const int t = tgt[r];
        if (t < 0 || t >= n_cls) {
            ++n_err;
            continue;
        }
        const float w = t == 0 ? eos_coef : 1.0f;
        a0 += w * (row_lse[r] - ldf(logits + (size_t)r * ld + t));
        a1 += w;
        if (t != 0) {
            a2 += 1.f;
            if (row_argmax && row_argmax[r] == t) a3 += 1.f;
        }
    }
    __shared__ float red[4][32];
    a0 = warp_sum(a0), a1 = warp_sum(a1), a2 = warp_sum(a2), a3 = warp_sum(a3);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) red[0][warp] = a0, red[1][warp] = a1, red[2][warp] = a2, red[3][warp] = a3;
    __syncthreads();
    if (threadIdx.x < 4) {
        float s = 0.f;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[threadIdx.x][w];
        atomicAdd(out + threadIdx.x, s);
    }
    if (err_out && n_err) atomicAdd(err_out, (float)n_err);
}

template <typename T, typename TO>
__global__ void __launch_bounds__(256) set_loss_bwd_kernel(const T* __restrict__ logits, int n_rows, int n_cls,
                                                          int ld, const int32_t* __restrict__ tgt,
                                                          const float* __restrict__ row_lse, float eos_coef,
                                                          const float* __restrict__ out, float grad_scale,
                                                          TO* __restrict__ dlogits, int ldo) {
    const int lane = threadIdx.x & 31;
    const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= n_rows) return;
    const int t = tgt[r];
    // rows with an out-of-range class were excluded from the loss (and flagged) by the forward: zero gradient
    const bool t_ok = t >= 0 && t < n_cls;
    const float s = t_ok ? grad_scale * (t == 0 ? eos_coef : 1.0f) / out[1] : 0.f;
    const float lse = row_lse[r];
    const T* x = logits + (size_t)r * ld;
    TO* g = dlogits + (size_t)r * ldo;
    for (int c = lane; c < n_cls; c += 32) {
        const float p = expf(ldf(x + c) - lse);
        const float gv = s * (p - (c == t ? 1.0f : 0.0f));
        if constexpr (sizeof(TO) == 2) g[c] = __float2bfloat16(gv);
        else g[c] = gv;
    }
}

}  // namespace shg

extern "C" {

int shg_lsap_batched(const float* cost, int n_frames, int q_f, int ld, const int32_t* col_off, int32_t* rows,
                     int32_t* cols, int32_t* n_match, int32_t* status, void* stream) {
    SHG_REQUIRE(q_f >= 1 && q_f <= shg::kMaxDim, SHG_ESHAPE, "shg_lsap_batched: q_f=%d outside [1,32]", q_f);
    SHG_REQUIRE(n_frames >= 0 && ld >= 0, SHG_EINVAL, "shg_lsap_batched: negative size");
    if (n_frames == 0) return SHG_OK;
    shg::launch(shg::lsap_frames_kernel, n_frames, 32, 0, (cudaStream_t)stream, cost, q_f, ld, col_off, rows, cols, n_match,
                                                                      status);
    return shg::check_launch("shg_lsap_batched");
}

int shg_match_from_logits(const void* logits, int logits_bf16, int n_frames, int q_f, int n_cls, int ld_logits,
                          const int32_t* labels, const int32_t* label_off, int32_t* rows, int32_t* cols,
                          int32_t* n_match, int32_t* status, float* row_max, float* row_lse, int32_t* row_argmax,
                          int32_t* target_classes, void* stream) {
    SHG_REQUIRE(q_f >= 1 && q_f <= shg::kMaxDim, SHG_ESHAPE, "shg_match_from_logits: q_f=%d outside [1,32]", q_f);
    SHG_REQUIRE(n_cls >= 1 && ld_logits >= n_cls, SHG_EINVAL, "shg_match_from_logits: bad class count / ld");
    if (n_frames == 0) return SHG_OK;
    const int threads = 32 * (q_f < 8 ? q_f : 8);
    if (logits_bf16)
        shg::launch(shg::match_from_logits_kernel<__nv_bfloat16>, n_frames, threads, 0, (cudaStream_t)stream, 
            (const __nv_bfloat16*)logits, q_f, n_cls, ld_logits, labels, label_off, rows, cols, n_match, status,
            row_max, row_lse, row_argmax, target_classes);
    else
        shg::launch(shg::match_from_logits_kernel<float>, n_frames, threads, 0, (cudaStream_t)stream, 
            (const float*)logits, q_f, n_cls, ld_logits, labels, label_off, rows, cols, n_match, status, row_max,
            row_lse, row_argmax, target_classes);
    return shg::check_launch("shg_match_from_logits");
}

int shg_set_loss_fwd(const void* logits, int logits_bf16, int n_rows, int n_cls, int ld_logits,
                     const int32_t* target_classes, const float* row_lse, const int32_t* row_argmax, float eos_coef,
                     float* out, const int32_t* status, int n_frames, float* err_out, void* stream) {
    if (n_rows == 0) return SHG_OK;
    // up to 32768 rows (256 clips of 128 queries) ONE block sums in a fixed order: the weight sum out[1] divides every
    // logit gradient, and a last-bit difference between two runs of the same step (atomic order over blocks) is
    // amplified by the bf16 roundings of the backward into 1e-4-level gradient differences.  Larger inputs fall back
    // to per-block partial sums combined with atomics.
    const bool one_block = n_rows <= 32768;
    const int threads = one_block ? 1024 : 256;
    const int blocks = one_block ? 1 : ((n_rows + 255) / 256 < 148 ? (n_rows + 255) / 256 : 148);
    if (logits_bf16)
        shg::launch(shg::set_loss_fwd_kernel<__nv_bfloat16>, blocks, threads, 0, (cudaStream_t)stream, 
            (const __nv_bfloat16*)logits, n_rows, ld_logits, target_classes, row_lse, row_argmax, eos_coef, out,
            n_cls, status, n_frames, err_out);
    else
        shg::launch(shg::set_loss_fwd_kernel<float>, blocks, threads, 0, (cudaStream_t)stream, 
            (const float*)logits, n_rows, ld_logits, target_classes, row_lse, row_argmax, eos_coef, out, n_cls, status,
            n_frames, err_out);
    return shg::check_launch("shg_set_loss_fwd");
}

int shg_set_loss_bwd(const void* logits, int logits_bf16, int n_rows, int n_cls, int ld_logits,
                     const int32_t* target_classes, const float* row_lse, float eos_coef, const float* out,
                     float grad_scale, void* dlogits, int dlogits_bf16, int ld_dlogits, void* stream) {
    if (n_rows == 0) return SHG_OK;
    const int blocks = (n_rows + 7) / 8;
    cudaStream_t s = (cudaStream_t)stream;
#define SHG_LAUNCH(T, TO)                                                                                          \
    shg::launch(shg::set_loss_bwd_kernel<T, TO>, blocks, 256, 0, s, (const T*)logits, n_rows, n_cls, ld_logits,             \
                                                           target_classes, row_lse, eos_coef, out, grad_scale,     \
                                                           (TO*)dlogits, ld_dlogits)
    if (logits_bf16 && dlogits_bf16) SHG_LAUNCH(__nv_bfloat16, __nv_bfloat16);
    else if (logits_bf16) SHG_LAUNCH(__nv_bfloat16, float);
    else if (dlogits_bf16) SHG_LAUNCH(float, __nv_bfloat16);
    else SHG_LAUNCH(float, float);
#undef SHG_LAUNCH
    return shg::check_launch("shg_set_loss_bwd");
}
}
