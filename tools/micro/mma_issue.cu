// Microbenchmark: what does ONE thread pay to issue small tcgen05.mma instructions, and how long do they take?
// M = 128, K = 16 (one bf16 k-step), N in {16..256}; A from shared memory (SS) or tensor memory (TS); one or two
// CTAs per SM issuing at the same time.  Prints cycles per MMA for the issue loop alone and including completion.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_issue mma_issue.cu && ./mma_issue
#include <cstdio>
#include <cuda_runtime.h>
#include "../../shg_vqa_b200/csrc/shg_common.cuh"
using namespace shg;
namespace shg { void set_error(const char*, ...) {} int check_launch(const char*) { return 0; } }

template <int MODE>   // 0: SS (A, B K-major)  1: TS  2: SS, A and B MN-major (the merged backward's dK / dV)  3: SS, A K-major, B MN-major (dQ)
__global__ void __launch_bounds__(128) k(int n, int iters, unsigned long long* out) {
    constexpr bool TS = MODE == 1;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint32_t slot;
    __shared__ uint64_t bar;
    for (int i = threadIdx.x; i < 49152 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (threadIdx.x < 32) tmem_alloc(&slot, 256);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = slot;
    if (threadIdx.x == 0) {
        const uint32_t a = smem_u32(smem), b = smem_u32(smem + 16384);
        const uint32_t idesc = make_idesc_bf16(128, n, MODE == 2 ? 1 : 0, MODE >= 2 ? 1 : 0);
        for (int rep = 0; rep < 2; ++rep) {
            const unsigned long long t0 = clock64();
            for (int it = 0; it < iters; ++it) {
                const int ks = it & 3;
                if (TS) umma_bf16_ts(tmem, tmem + 128 + ks * 8, make_smem_desc(b + ks * 32, 16, 1024), idesc, it > 0);
                else if (MODE == 2) umma_bf16(tmem, make_smem_desc(a + ks * 2048, 16384, 1024), make_smem_desc(b + ks * 2048, 8192, 1024), idesc, it > 0);
                else if (MODE == 3) umma_bf16(tmem, make_smem_desc(a + ks * 32, 16, 1024), make_smem_desc(b + ks * 2048, 8192, 1024), idesc, it > 0);
                else umma_bf16(tmem, make_smem_desc(a + ks * 32, 16, 1024), make_smem_desc(b + ks * 32, 16, 1024), idesc, it > 0);
            }
            const unsigned long long t1 = clock64();
            tc_commit(&bar);
            mbar_wait(&bar, rep & 1);
            const unsigned long long t2 = clock64();
            if (rep == 1) out[blockIdx.x * 2] = t1 - t0, out[blockIdx.x * 2 + 1] = t2 - t0;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tmem, 256); }
}

int main() {
    unsigned long long* d;
    cudaMalloc(&d, 296 * 16);
    cudaFuncSetAttribute(k<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 60 * 1024);
    cudaFuncSetAttribute(k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 60 * 1024);
    cudaFuncSetAttribute(k<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 60 * 1024);
    cudaFuncSetAttribute(k<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 60 * 1024);
    const int iters = 512;
    for (int ctas : {148, 296})
        for (int ts = 0; ts < 4; ++ts)
            for (int n : {16, 32, 64, 128, 256}) {
                if (ts >= 2 && n != 64) continue;
                if (ts == 1) k<1><<<ctas, 128, 50 * 1024>>>(n, iters, d);
                else if (ts == 2) k<2><<<ctas, 128, 50 * 1024>>>(n, iters, d);
                else if (ts == 3) k<3><<<ctas, 128, 50 * 1024>>>(n, iters, d);
                else k<0><<<ctas, 128, 50 * 1024>>>(n, iters, d);
                cudaError_t e = cudaDeviceSynchronize();
                unsigned long long h[592];
                cudaMemcpy(h, d, ctas * 16, cudaMemcpyDeviceToHost);
                double is = 0, tot = 0;
                for (int i = 0; i < ctas; ++i) is += h[2 * i], tot += h[2 * i + 1];
                printf("%d CTA/SM %s N=%3d: issue %.1f clk/MMA, issue+complete %.1f clk/MMA (ideal compute %d clk) %s\n", ctas / 148,
                       ts == 0 ? "SS" : (ts == 1 ? "TS" : (ts == 2 ? "SS A,B MN-major" : "SS B MN-major")), n, is / ctas / iters, tot / ctas / iters, n / 2, cudaGetErrorString(e));
            }
    return 0;
}
