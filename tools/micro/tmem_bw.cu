// Microbenchmark: tcgen05.ld throughput per SM (how fast can softmax threads pull fp32 scores out of TMEM?).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_bw tmem_bw.cu && ./tmem_bw
#include <cstdio>
#include <cuda_runtime.h>
#include "../../shg_vqa_b200/csrc/shg_common.cuh"
using namespace shg;
namespace shg { void set_error(const char*, ...) {} int check_launch(const char*) { return 0; } }

template <int X>
__global__ void __launch_bounds__(512) k(int iters, int n_warps, unsigned long long* out, float* sink) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) tmem_alloc(&slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = slot;
    const uint32_t t_row = tmem + (uint32_t((warp & 3) * 32) << 16);
    float acc = 0.f;
    unsigned long long t0 = clock64();
    if (warp < n_warps) {
        for (int it = 0; it < iters; ++it) {
            uint32_t v[32];
            if (X == 32) tmem_ld32(t_row + ((it * 32 + (warp >> 2) * 64) & 255), v);
            else { uint32_t w[16]; tmem_ld16(t_row + ((it * 16) & 255), w);
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = w[i]; }
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < (X == 32 ? 32 : 16); ++i) acc += __uint_as_float(v[i]);
        }
    }
    __syncthreads();
    unsigned long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
    if (acc == 123.456f) sink[0] = acc;
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

int main() {
    unsigned long long* d; float* s;
    cudaMalloc(&d, 148 * 8); cudaMalloc(&s, 4);
    const int iters = 2000;
    for (int nw : {1, 4, 8, 16}) {
        k<32><<<148, 512>>>(iters, nw, d, s);
        cudaDeviceSynchronize();
        k<32><<<148, 512>>>(iters, nw, d, s);
        cudaError_t e = cudaDeviceSynchronize();
        unsigned long long h[148];
        cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
        double clk = (double)h[0];
        double bytes = (double)nw * iters * 32 * 32 * 4;
        printf("x32: %2d warps: %8.0f clk for %d loads/warp -> %.1f B/clk/SM (%.1f clk per warp-load) %s\n", nw, clk, iters,
               bytes / clk, clk / iters, cudaGetErrorString(e));
    }
    return 0;
}
