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

// VEC = 4: each thread takes four consecutive symbols with 16-byte loads and stores (arrays 16-byte
// aligned); the four evaluations are independent, which gives the scheduler eight FP64 chains per
// thread.  VEC = 1: any alignment, and the tail of a vector launch.
template <int VEC>
__global__ void __launch_bounds__(256)
cdf_tables_kernel(const float* __restrict__ x, const float* __restrict__ mean,
                  const float* __restrict__ scale, int64_t n, uint32_t* __restrict__ start,
                  uint32_t* __restrict__ freq, int32_t* __restrict__ status_word) {
    __shared__ __align__(256) uint64_t s_tab[32];
    const ExpTab tab = stage_exp_table(s_tab);
    int32_t flags = 0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t groups = n / VEC;
    if (i < groups) {
        // software pipeline: the next group's parameters are in flight while this one is evaluated,
        // so HBM latency never sits in front of the arithmetic
        float xv[VEC], mv[VEC], sv[VEC];
        auto load = [&](int64_t g, float* a, float* b, float* c) {
            if (VEC == 4) {
                const float4 p = __ldg(reinterpret_cast<const float4*>(x) + g);
                const float4 q = __ldg(reinterpret_cast<const float4*>(mean) + g);
                const float4 r = __ldg(reinterpret_cast<const float4*>(scale) + g);
                a[0] = p.x; a[1] = p.y; a[2] = p.z; a[3] = p.w;
                b[0] = q.x; b[1] = q.y; b[2] = q.z; b[3] = q.w;
                c[0] = r.x; c[1] = r.y; c[2] = r.z; c[3] = r.w;
            } else {
                a[0] = __ldg(x + g); b[0] = __ldg(mean + g); c[0] = __ldg(scale + g);
            }
        };
        load(i, xv, mv, sv);
        for (; i < groups; i += stride) {
            float xn[VEC], mn[VEC], sn[VEC];
#pragma unroll
            for (int k = 0; k < VEC; ++k) { xn[k] = 0.0f; mn[k] = 0.0f; sn[k] = 1.0f; }
            if (i + stride < groups) load(i + stride, xn, mn, sn);
            uint32_t st[VEC], fr[VEC];
#pragma unroll
            for (int k = 0; k < VEC; ++k) {
                const SymbolTable t = make_table(xv[k], mv[k], sv[k], tab, flags);
                st[k] = t.start;
                fr[k] = t.freq;
            }
            if (VEC == 4) {
                reinterpret_cast<uint4*>(start)[i] = make_uint4(st[0], st[1], st[2], st[3]);
                reinterpret_cast<uint4*>(freq)[i] = make_uint4(fr[0], fr[1], fr[2], fr[3]);
            } else {
                start[i] = st[0];
                freq[i] = fr[0];
            }
#pragma unroll
            for (int k = 0; k < VEC; ++k) { xv[k] = xn[k]; mv[k] = mn[k]; sv[k] = sn[k]; }
        }
    }
    if (flags) atomicOr(status_word, flags);
}

// Diagnostic: the device expf restatement on its own, so that tests can sweep it against the host
// libm over the whole reachable domain.
__global__ void __launch_bounds__(256)
debug_expf_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t n) {
    __shared__ __align__(256) uint64_t s_tab[32];
    const ExpTab tab = stage_exp_table(s_tab);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        y[i] = expf_glibc(x[i], tab);
}

// Diagnostic: part1 as a function of the float argument alone (the quotient stage replaced by a
// plain widening), for the exhaustive sweep against the reference arithmetic.
__global__ void __launch_bounds__(256)
debug_part1_kernel(const float* __restrict__ arg, int32_t* __restrict__ y, int64_t n) {
    __shared__ __align__(256) uint64_t s_tab[32];
    const ExpTab tab = stage_exp_table(s_tab);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        y[i] = part1_from_arg(arg_from_quotient((double)arg[i]), tab);
}

cudaError_t launch_debug_part1(const float* arg, int32_t* y, int64_t n, cudaStream_t stream) {
    if (n <= 0) return cudaSuccess;
    int64_t blocks = (n + 255) / 256;
    const int64_t cap = (int64_t)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    debug_part1_kernel<<<(unsigned)blocks, 256, 0, stream>>>(arg, y, n);
    return cudaGetLastError();
}

// Diagnostics: the two reciprocal-based divisions of the coder against the hardware's own exact
// division, on operands drawn the way the coder produces them (counter-based generator, so a
// test can cover 2^32 cases in a second).  mismatches is a device counter.
__device__ __forceinline__ uint64_t splitmix64(uint64_t z) {
    z += 0x9e3779b97f4a7c15ull;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}

// mode 0: realistic magnitudes (|mean| <= 4, scale in [2^-37, 2^22], symbol inside its window);
// mode 1: any finite positive float scale (subnormals included), any float mean with |mean| <= 16384,
//         any symbol index below 2^22.
__global__ void __launch_bounds__(256)
debug_div_check_kernel(int64_t n, uint64_t seed, int mode, unsigned long long* __restrict__ mismatches) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    unsigned long long bad = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint64_t r1 = splitmix64(seed + 3 * (uint64_t)i), r2 = splitmix64(seed + 3 * (uint64_t)i + 1),
                       r3 = splitmix64(seed + 3 * (uint64_t)i + 2);
        float mean, scale;
        int k;
        if (mode == 0) {
            mean = (float)((double)(int64_t)(r1 % 2000001) / 1000000.0 - 1.0) * 4.0f;
            const uint32_t bits = ((90u + (uint32_t)(r2 % 60)) << 23) | (uint32_t)(r2 >> 41);
            scale = __uint_as_float(bits);
            k = (int)(r3 % 4097) - 2048 + (int)(mean * 256.0f);
        } else {
            mean = __uint_as_float((uint32_t)r1);
            if (!(fabsf(mean) <= 16384.0f)) mean = 0.25f;
            uint32_t bits = (uint32_t)(r2 >> 33);
            if (bits == 0 || bits >= 0x7f800000u) bits = 0x3f800001u;
            scale = __uint_as_float(bits);
            k = (int)(r3 % 8388608) - 4194304;
        }
        const SymbolModel m = make_model(mean, scale);
        const double a = dsub(half_bin_point(k), m.mean_d);
        const double q = div_by_scale(a, m);
        const double want = __ddiv_rn(a, (double)scale);
        bad += (__double_as_longlong(q) != __double_as_longlong(want));
    }
    if (bad) atomicAdd(mismatches, bad);
}

// rans_push against 64-bit integer division over (state, freq) pairs that satisfy the coder's
// invariants (2^32 <= state < 2^64, 1 <= freq <= 2^24, start + freq <= 2^24).
__global__ void __launch_bounds__(256)
debug_push_check_kernel(int64_t n, uint64_t seed, unsigned long long* __restrict__ mismatches) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    unsigned long long bad = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint64_t r1 = splitmix64(seed + 3 * (uint64_t)i), r2 = splitmix64(seed + 3 * (uint64_t)i + 1),
                       r3 = splitmix64(seed + 3 * (uint64_t)i + 2);
        uint32_t freq = (uint32_t)(r1 % 16777216) + 1;                    // 1 .. 2^24
        if ((r3 & 7) == 0) freq = (uint32_t)(r3 >> 8) % 64 + 1;           // small frequencies too
        if ((r3 & 0xf00) == 0) freq = 1u << ((r3 >> 12) % 25);            // powers of two
        uint64_t state = r2 | 0x100000000ull;                             // any state >= 2^32
        if ((r3 & 0x30) == 0) state = ((uint64_t)freq << 40) - 1 - (r2 & 0xffff);   // just below a renormalisation
        if ((r3 & 0xc0) == 0) state = ((uint64_t)freq << 40) + (r2 & 0xffff);       // just above
        const uint32_t start = (uint32_t)((r3 >> 40) % (16777216u - freq + 1));
        // what the reference does (rans.pyx:62-65)
        uint64_t ref = state;
        uint32_t ref_word = 0;
        bool ref_emit = false;
        if (ref >= ((uint64_t)freq << 40)) { ref_word = (uint32_t)ref; ref >>= 32; ref_emit = true; }
        ref = ((ref / freq) << 24) + (ref % freq) + start;
        uint64_t st = state;
        uint32_t word = 0;
        const bool emit = rans_push(st, start, freq, word);
        bad += (st != ref) || (emit != ref_emit) || (emit && word != ref_word);
    }
    if (bad) atomicAdd(mismatches, bad);
}

cudaError_t launch_debug_div_check(int64_t n, uint64_t seed, int mode, unsigned long long* mismatches, cudaStream_t stream) {
    if (n <= 0) return cudaSuccess;
    debug_div_check_kernel<<<sm_count() * 8, 256, 0, stream>>>(n, seed, mode, mismatches);
    return cudaGetLastError();
}

cudaError_t launch_debug_push_check(int64_t n, uint64_t seed, unsigned long long* mismatches, cudaStream_t stream) {
    if (n <= 0) return cudaSuccess;
    debug_push_check_kernel<<<sm_count() * 8, 256, 0, stream>>>(n, seed, mismatches);
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
    const int64_t cap = (int64_t)sms * 8;  // 8 resident CTAs of 256 threads per SM, grid-stride beyond
    const bool aligned = (((uintptr_t)x | (uintptr_t)mean | (uintptr_t)scale | (uintptr_t)start | (uintptr_t)freq) & 15) == 0;
    const int64_t n4 = aligned ? n / 4 * 4 : 0;
    if (n4 > 0) {
        int64_t blocks = (n4 / 4 + threads - 1) / threads;
        if (blocks > cap) blocks = cap;
        cdf_tables_kernel<4><<<(unsigned)blocks, threads, 0, stream>>>(x, mean, scale, n4, start, freq, status_word);
    }
    if (n > n4) {   // unaligned arrays, or the last 1-3 symbols
        const int64_t m = n - n4;
        int64_t blocks = (m + threads - 1) / threads;
        if (blocks > cap) blocks = cap;
        cdf_tables_kernel<1><<<(unsigned)blocks, threads, 0, stream>>>(x + n4, mean + n4, scale + n4, m, start + n4,
                                                                       freq + n4, status_word);
    }
    return cudaGetLastError();
}

}  // namespace flic
