// K1 -- fused CDF evaluation + 24-bit frequency quantisation.
//
// Replaces encode pass 1 of the reference (rans/rans.pyx:49-56, generated C rans/rans.cpp:1668-1755)
// plus the two CDF() calls per symbol (rans.pyx:31-35).  One thread per symbol, grid-stride,
// fully coalesced 4-byte loads of x / mean / scale and 4-byte stores of start / freq
// (algorithmic traffic 12 B in + 8 B out per symbol).  The arithmetic is FP64-heavy (two
// correctly-rounded reciprocals, two glibc-expf evaluations), so this kernel is bound by the
// FP64 / conversion pipes rather than by HBM; see DESIGN.md for the roofline.
#include "flic_device.cuh"
#include "flic_kernels.cuh"

namespace flic {

__global__ void __launch_bounds__(256)
cdf_tables_kernel(const float* __restrict__ x, const float* __restrict__ mean,
                  const float* __restrict__ scale, int64_t n, uint32_t* __restrict__ start,
                  uint32_t* __restrict__ freq, int32_t* __restrict__ status_word) {
    __shared__ uint64_t s_tab[32];
    stage_exp_table(s_tab);
    int32_t flags = 0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    // software pipeline: the next symbol's parameters are in flight while this one is evaluated,
    // so HBM latency never sits in front of the arithmetic
    float xv = __ldg(x + i), mv = __ldg(mean + i), sv = __ldg(scale + i);
    for (; i < n; i += stride) {
        const int64_t nx = i + stride;
        float xn = 0.0f, mn = 0.0f, sn = 1.0f;
        if (nx < n) { xn = __ldg(x + nx); mn = __ldg(mean + nx); sn = __ldg(scale + nx); }
        const SymbolTable t = make_table(xv, mv, sv, s_tab, flags);
        start[i] = t.start;
        freq[i] = t.freq;
        xv = xn; mv = mn; sv = sn;
    }
    if (flags) atomicOr(status_word, flags);
}

// Diagnostic: the device expf restatement on its own, so that tests can sweep it against the host
// libm over the whole reachable domain.
__global__ void __launch_bounds__(256)
debug_expf_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t n) {
    __shared__ uint64_t s_tab[32];
    stage_exp_table(s_tab);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        y[i] = expf_glibc(x[i], s_tab);
}

// Diagnostic: part1 as a function of the float argument alone (the quotient stage replaced by a
// plain widening), for the exhaustive sweep against the reference arithmetic.
__global__ void __launch_bounds__(256)
debug_part1_kernel(const float* __restrict__ arg, int32_t* __restrict__ y, int64_t n) {
    __shared__ uint64_t s_tab[32];
    stage_exp_table(s_tab);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        y[i] = part1_from_arg(arg_from_quotient((double)arg[i]), s_tab);
}

cudaError_t launch_debug_part1(const float* arg, int32_t* y, int64_t n, cudaStream_t stream) {
    if (n <= 0) return cudaSuccess;
    int64_t blocks = (n + 255) / 256;
    const int64_t cap = (int64_t)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    debug_part1_kernel<<<(unsigned)blocks, 256, 0, stream>>>(arg, y, n);
    return cudaGetLastError();
}

cudaError_t launch_debug_expf(const float* x, float* y, int64_t n, cudaStream_t stream) {
    if (n <= 0) return cudaSuccess;
    int64_t blocks = (n + 255) / 256;
    const int64_t cap = (int64_t)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    debug_expf_kernel<<<(unsigned)blocks, 256, 0, stream>>>(x, y, n);
    return cudaGetLastError();
}

cudaError_t launch_cdf_tables(const float* x, const float* mean, const float* scale, int64_t n,
                              uint32_t* start, uint32_t* freq, int32_t* status_word,
                              cudaStream_t stream) {
    if (n <= 0) return cudaSuccess;
    const int sms = sm_count();
    const int threads = 256;
    int64_t blocks = (n + threads - 1) / threads;
    const int64_t cap = (int64_t)sms * 8;  // 8 resident CTAs of 256 threads per SM, grid-stride beyond
    if (blocks > cap) blocks = cap;
    cdf_tables_kernel<<<(unsigned)blocks, threads, 0, stream>>>(x, mean, scale, n, start, freq, status_word);
    return cudaGetLastError();
}

}  // namespace flic
