// Error reporting, device check and the host-side TMA descriptor cache.
#include <cudaTypedefs.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <unordered_map>
#include <vector>

#include "shg_common.cuh"

namespace shg {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return SHG_ECUDA;
    }
    return SHG_OK;
}

struct TmapKey {
    uint64_t w[20];
    bool operator==(const TmapKey& o) const { return memcmp(w, o.w, sizeof(w)) == 0; }
};
struct TmapHash {
    size_t operator()(const TmapKey& k) const {
        uint64_t h = 1469598103934665603ull;
        for (int i = 0; i < 20; ++i) {
            h ^= k.w[i];
            h *= 1099511628211ull;
        }
        return (size_t)h;
    }
};

static PFN_cuTensorMapEncodeTiled_v12000 encode_fn() {
    static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
    }
    return fn;
}

const CUtensorMap* get_tmap(const void* ptr, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                            const uint32_t* box, int dtype) {
    static std::mutex mu;
    static std::unordered_map<TmapKey, CUtensorMap*, TmapHash> cache;
    int dev = 0;
    cudaGetDevice(&dev);
    if (rank < 1 || rank > 5) {
        set_error("tensor map rank %d", rank);
        return nullptr;
    }
    TmapKey key;
    memset(&key, 0, sizeof(key));
    key.w[0] = reinterpret_cast<uint64_t>(ptr);
    key.w[1] = (uint64_t)rank | ((uint64_t)dev << 8) | ((uint64_t)dtype << 16);
    for (int i = 0; i < rank; ++i) {
        key.w[2 + i] = dims[i];
        key.w[7 + i] = i ? strides_bytes[i - 1] : 0;
        key.w[12 + i] = box[i];
    }
    std::lock_guard<std::mutex> lock(mu);
    auto it = cache.find(key);
    if (it != cache.end()) return it->second;
    auto fn = encode_fn();
    if (!fn) {
        set_error("cuTensorMapEncodeTiled entry point not found");
        return nullptr;
    }
    if ((reinterpret_cast<uint64_t>(ptr) & 15) != 0) {
        set_error("TMA operand base %p is not 16-byte aligned", ptr);
        return nullptr;
    }
    for (int i = 0; i + 1 < rank; ++i)
        if (strides_bytes[i] % 16 != 0) {
            set_error("TMA operand stride %llu B is not a multiple of 16", (unsigned long long)strides_bytes[i]);
            return nullptr;
        }
    CUtensorMap* m = new CUtensorMap;
    cuuint64_t gdim[5] = {1, 1, 1, 1, 1};
    cuuint64_t gstr[4] = {0, 0, 0, 0};
    cuuint32_t bx[5] = {1, 1, 1, 1, 1};
    cuuint32_t es[5] = {1, 1, 1, 1, 1};
    for (int i = 0; i < rank; ++i) {
        gdim[i] = dims[i];
        bx[i] = box[i];
    }
    for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides_bytes[i];
    CUresult r = fn(m, dtype == kTmapF32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(ptr), gdim, gstr, bx, es,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with CUresult %d (rank %d dims %llu,%llu,%llu box %u,%u,%u)", (int)r,
                  rank, (unsigned long long)gdim[0], (unsigned long long)gdim[1], (unsigned long long)gdim[2], bx[0],
                  bx[1], bx[2]);
        delete m;
        return nullptr;
    }
    cache.emplace(key, m);
    return m;
}

}  // namespace shg

extern "C" {

const char* shg_last_error(void) { return shg::g_err; }

int shg_version(void) { return 200; }

int shg_device_check(int device) {
    cudaDeviceProp p;
    cudaError_t e = cudaGetDeviceProperties(&p, device);
    if (e != cudaSuccess) {
        shg::set_error("cudaGetDeviceProperties(%d): %s", device, cudaGetErrorString(e));
        return SHG_ECUDA;
    }
    if (p.major != 10) {
        shg::set_error("device %d is sm_%d%d; libshgvqa only runs on sm_100 (B200) and has no fallback", device,
                       p.major, p.minor);
        return SHG_EDEVICE;
    }
    return SHG_OK;
}
}
