// Gradient-norm clipping and BertAdam over the flat parameter buffers.
//
// Replaces nn.utils.clip_grad_norm_(model.parameters(), 5.) + BertAdam.step() (tasks/agqaHGQA.py:391-392,
// lxrt/optimization.py:101-180): a Python loop over 576 tensors x ~8 elementwise kernels becomes two launches
// that stream the flat fp32 gradient / moment / master buffers once and rewrite the bf16 shadow the tensor
// cores read.  Only "live" segments (parameters that receive a gradient; the reference skips p.grad is None)
// are touched — dead-but-serialised branches keep their values, weight decay included.
#include "shg_common.cuh"

namespace shg {

// sum of squares over the live segments -> out[0].  DETERMINISTIC: per-block partials go to `ws`, the last block to
// arrive (counter in ws[kSumsqWs - 1]) adds them in a fixed order.  Data-parallel ranks hold identical all-reduced
// gradients and must derive the identical clip coefficient from them — fp32 atomics across blocks would make the
// norm, hence every moment and weight, differ between ranks in the last bits (measured: 60 M of 191 M moments after
// three steps), and replicas would drift apart.
constexpr int kSumsqWs = 2048;
__global__ void __launch_bounds__(256) grad_sumsq_kernel(const float* __restrict__ g,
                                                        const int64_t* __restrict__ seg_start,
                                                        const int64_t* __restrict__ seg_prefix, int n_seg,
                                                        int64_t total4, float* __restrict__ out, float* __restrict__ ws) {
    float acc = 0.f;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total4; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t e = i * 4;
        int lo = 0, hi = n_seg - 1;           // last segment whose prefix <= e
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (seg_prefix[mid] <= e) lo = mid;
            else hi = mid - 1;
        }
        const float4 v = *reinterpret_cast<const float4*>(g + seg_start[lo] + (e - seg_prefix[lo]));
        acc += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    }
    acc = warp_sum(acc);
    __shared__ float red[8];
    __shared__ int s_last;
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
        for (int w = 0; w < 8; ++w) s += red[w];
        ws[blockIdx.x] = s;
        __threadfence();
        unsigned int* counter = reinterpret_cast<unsigned int*>(ws + kSumsqWs - 1);
        s_last = atomicAdd(counter, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    float t = 0.f;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) t += __ldcg(ws + i);      // fixed order per thread
    t = warp_sum(t);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = t;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
        for (int w = 0; w < 8; ++w) s += red[w];
        out[0] = s;
        *reinterpret_cast<unsigned int*>(ws + kSumsqWs - 1) = 0u;       // ready for the next launch / graph replay
    }
}

// clip_coef = min(1, max_norm / (sqrt(sumsq) + 1e-6));  g *= clip_coef
// m = b1 m + (1-b1) g ; v = b2 v + (1-b2) g^2 ; p -= lr * (m / (sqrt(v) + eps) + wd * p)   (no bias correction)
__global__ void __launch_bounds__(256) bert_adam_kernel(float* __restrict__ p, float* __restrict__ g,
                                                       float* __restrict__ m, float* __restrict__ v,
                                                       __nv_bfloat16* __restrict__ shadow, int64_t n_shadow,
                                                       const int64_t* __restrict__ seg_start,
                                                       const int64_t* __restrict__ seg_prefix, int n_seg,
                                                       int64_t total4, const float* __restrict__ sumsq,
                                                       float max_norm, float lr, const float* __restrict__ lr_dev,
                                                       float b1, float b2, float eps, float wd, float grad_scale,
                                                       int zero_grad) {
    if (lr_dev) lr = lr_dev[0];     // device-resident learning rate: the launch stays valid inside a CUDA graph
    float clip = 1.0f;
    if (max_norm > 0.f) {
        const float norm = sqrtf(sumsq[0]) * grad_scale;
        clip = fminf(1.0f, max_norm / (norm + 1e-6f));
    }
    const float gs = clip * grad_scale;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total4; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t e = i * 4;
        int lo = 0, hi = n_seg - 1;
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (seg_prefix[mid] <= e) lo = mid;
            else hi = mid - 1;
        }
        const int64_t off = seg_start[lo] + (e - seg_prefix[lo]);
        float4 pv = *reinterpret_cast<float4*>(p + off);
        float4 gv = *reinterpret_cast<float4*>(g + off);
        float4 mv = *reinterpret_cast<float4*>(m + off);
        float4 vv = *reinterpret_cast<float4*>(v + off);
        float* pp = &pv.x; float* gp = &gv.x; float* mp = &mv.x; float* vp = &vv.x;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float gk = gp[k] * gs;
            mp[k] = b1 * mp[k] + (1.0f - b1) * gk;
            vp[k] = b2 * vp[k] + (1.0f - b2) * gk * gk;
            const float upd = mp[k] / (sqrtf(vp[k]) + eps) + wd * pp[k];
            pp[k] -= lr * upd;
        }
        *reinterpret_cast<float4*>(p + off) = pv;
        *reinterpret_cast<float4*>(m + off) = mv;
        *reinterpret_cast<float4*>(v + off) = vv;
        if (zero_grad) *reinterpret_cast<float4*>(g + off) = make_float4(0.f, 0.f, 0.f, 0.f);
        if (off < n_shadow) {
            uint2 u;
            u.x = pack_bf16x2(pv.x, pv.y), u.y = pack_bf16x2(pv.z, pv.w);
            *reinterpret_cast<uint2*>(shadow + off) = u;
        }
    }
}

}  // namespace shg

extern "C" {

int shg_grad_sumsq(const float* grad, const int64_t* seg_start, const int64_t* seg_prefix, int n_seg,
                   int64_t total, float* out, float* workspace, void* stream) {
    using namespace shg;
    SHG_REQUIRE(total % 4 == 0 && n_seg >= 1, SHG_ESHAPE, "shg_grad_sumsq: total=%lld must be a multiple of 4",
                (long long)total);
    SHG_REQUIRE(workspace && out, SHG_EINVAL, "shg_grad_sumsq: workspace (2048 zero-initialised floats) and out are required");
    int64_t blocks = (total / 4 + 255) / 256;
    if (blocks > 8 * kNumSMs) blocks = 8 * kNumSMs;
    static_assert(8 * kNumSMs < kSumsqWs - 1, "partials must fit the workspace");
    shg::launch(grad_sumsq_kernel, (int)blocks, 256, 0, (cudaStream_t)stream, grad, seg_start, seg_prefix, n_seg, total / 4,
                                                                     out, workspace);
    return check_launch("shg_grad_sumsq");
}

int shg_bert_adam_step(float* param, float* grad, float* exp_avg, float* exp_avg_sq, void* shadow, int64_t n_shadow,
                       const int64_t* seg_start, const int64_t* seg_prefix, int n_seg, int64_t total,
                       const float* sumsq, float max_norm, float lr, const float* lr_dev, float b1, float b2, float eps,
                       float weight_decay, float grad_scale, int zero_grad, void* stream) {
    using namespace shg;
    SHG_REQUIRE(total % 4 == 0 && n_seg >= 1, SHG_ESHAPE, "shg_bert_adam_step: total=%lld must be a multiple of 4",
                (long long)total);
    SHG_REQUIRE(max_norm <= 0.f || sumsq, SHG_EINVAL, "shg_bert_adam_step: clipping needs the sum of squares");
    int64_t blocks = (total / 4 + 255) / 256;
    if (blocks > 8 * kNumSMs) blocks = 8 * kNumSMs;
    shg::launch(bert_adam_kernel, (int)blocks, 256, 0, (cudaStream_t)stream, 
        param, grad, exp_avg, exp_avg_sq, (__nv_bfloat16*)shadow, shadow ? n_shadow : 0, seg_start, seg_prefix, n_seg,
        total / 4, sumsq, max_norm, lr, lr_dev, b1, b2, eps, weight_decay, grad_scale, zero_grad);
    return check_launch("shg_bert_adam_step");
}
}
