// flic_device.cuh -- device-only helpers shared by the kernels.
#pragma once
#include "flic_core.cuh"

#include <atomic>

namespace flic {

// glibc's exp2f table; each TU keeps its own 256-byte constant copy and stages it into shared
// memory once per CTA (lanes index it divergently, which constant memory would serialise).
static __constant__ uint64_t c_exp2f_tab[32] = {FLIC_EXP2F_TABLE};

// s_tab must be declared __align__(256) (exp_tab_entry() ORs the entry offset into the address).
// Returns the table's shared-space address, made opaque so that it stays in one register instead
// of being recomputed from the shared-window base at every lookup.
__device__ __forceinline__ ExpTab stage_exp_table(uint64_t* s_tab) {
    if (threadIdx.x < 32) s_tab[threadIdx.x] = c_exp2f_tab[threadIdx.x];
    __syncthreads();
    uint32_t a = (uint32_t)__cvta_generic_to_shared(s_tab);
    asm volatile("" : "+r"(a));
    return a;
}

// CTA-wide barrier that may be reached from different code paths of a warp-specialised kernel
// (every thread of the CTA executes the same NUMBER of barriers; bar.sync counts arrivals).
__device__ __forceinline__ void cta_sync() { asm volatile("bar.sync 0;" ::: "memory"); }

__device__ __forceinline__ int64_t shfl_i64(int64_t v, int src) {
    int lo = __shfl_sync(0xffffffffu, (int)(uint32_t)(uint64_t)v, src);
    int hi = __shfl_sync(0xffffffffu, (int)(uint32_t)((uint64_t)v >> 32), src);
    return (int64_t)(((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo);
}

__device__ __forceinline__ int64_t warp_max_i64(int64_t v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        const int64_t o = shfl_i64(v, (threadIdx.x & 31) ^ d);
        v = o > v ? o : v;
    }
    return v;
}

// SM count of the CURRENT device, cached per device (a process may drive several GPUs from
// several threads; the cache entries are written once each with the same value, so a relaxed
// atomic per slot is enough).
inline int sm_count() {
    constexpr int kMaxDevices = 64;
    static std::atomic<int> cache[kMaxDevices] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) dev = 0;
    int sms = cache[dev].load(std::memory_order_relaxed);
    if (sms == 0) {
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
        cache[dev].store(sms, std::memory_order_relaxed);
    }
    return sms;
}

}  // namespace flic
