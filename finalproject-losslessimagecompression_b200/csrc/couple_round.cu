// K5 -- fused integer additive-coupling add-and-round, plus the uint8 <-> grid conversions.
//
// Replaces, in one pass over the b-channels only,
//   couplelib.py:49-52   txa = round(dense(xa)); zb = xb + txa; z = cat([za, zb])      (forward)
//   couplelib.py:58-60   txa = round(dense(za)); xb = zb - txa; x = cat([xa, xb])      (backward)
//   roundlib.py:18-38    Round(t) = STE(rint(t * 2^nbits)) / 2^nbits, torch.round = ties-to-even
// The reference materialises round(t), the sum and the concatenation as three to four elementwise
// passes over (B, C, H, W); here the b-channel slab of x is updated in place (the a-channels are
// untouched, so no cat): 4 B of t + 4 B of xb read, 4 B written per b-channel element.
//
// Bit-exactness: t*2^nbits and /2^nbits are exact scalings; rintf is ties-to-even like
// torch.round; the straight-through form x + (rint(x) - x) of BaseRound.forward is reproduced
// literally (it differs from rint(x) only in the sign of zero) with non-contracted adds.
#include "flic_device.cuh"
#include "flic_kernels.cuh"

namespace flic {

__device__ __forceinline__ float round_nbits(float t, float bins, float inv_bins) {
    const float xs = __fmul_rn(t, bins);
    const float y = rintf(xs);
    const float ste = __fadd_rn(xs, __fsub_rn(y, xs));  // roundlib.py:23-24
    return __fmul_rn(ste, inv_bins);                    // division by a power of two
}

// x viewed as (batch, channels*hw); the updated slab of image b is
// [b*channels*hw + a_ch*hw, (b+1)*channels*hw), t is (batch, slab) contiguous.
template <int VEC>
__global__ void __launch_bounds__(256)
couple_add_round_kernel(float* __restrict__ x, const float* __restrict__ t, int64_t batch,
                        int64_t img_stride, int64_t slab_off, int64_t slab, float sign, float bins,
                        float inv_bins) {
    const int64_t per_img = slab / VEC;
    const int64_t total = batch * per_img;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int64_t b = i / per_img;
        const int64_t k = (i - b * per_img) * VEC;
        float* xp = x + b * img_stride + slab_off + k;
        const float* tp = t + b * slab + k;
        if (VEC == 4) {
            const float4 tv = __ldg(reinterpret_cast<const float4*>(tp));
            float4 xv = *reinterpret_cast<float4*>(xp);
            xv.x = __fadd_rn(xv.x, sign * round_nbits(tv.x, bins, inv_bins));
            xv.y = __fadd_rn(xv.y, sign * round_nbits(tv.y, bins, inv_bins));
            xv.z = __fadd_rn(xv.z, sign * round_nbits(tv.z, bins, inv_bins));
            xv.w = __fadd_rn(xv.w, sign * round_nbits(tv.w, bins, inv_bins));
            *reinterpret_cast<float4*>(xp) = xv;
        } else {
            *xp = __fadd_rn(*xp, sign * round_nbits(__ldg(tp), bins, inv_bins));
        }
    }
}

cudaError_t launch_couple_add_round(float* x, const float* t, int64_t batch, int64_t channels,
                                    int64_t a_ch, int64_t hw, float sign, int nbits,
                                    cudaStream_t stream) {
    const int64_t slab = (channels - a_ch) * hw;
    if (batch <= 0 || slab <= 0) return cudaSuccess;
    const int64_t img_stride = channels * hw, slab_off = a_ch * hw;
    const float bins = (float)(1 << nbits), inv_bins = 1.0f / bins;
    const bool vec4 = (slab % 4 == 0) && (img_stride % 4 == 0) && (slab_off % 4 == 0) &&
                      ((uintptr_t)x % 16 == 0) && ((uintptr_t)t % 16 == 0);
    const int64_t work = batch * (vec4 ? slab / 4 : slab);
    int64_t blocks = (work + 255) / 256;
    const int64_t cap = (int64_t)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    if (vec4)
        couple_add_round_kernel<4><<<(unsigned)blocks, 256, 0, stream>>>(x, t, batch, img_stride, slab_off, slab, sign, bins, inv_bins);
    else
        couple_add_round_kernel<1><<<(unsigned)blocks, 256, 0, stream>>>(x, t, batch, img_stride, slab_off, slab, sign, bins, inv_bins);
    return cudaGetLastError();
}

// trainer.py:61,72: ToTensor (k / 255) then Round(nbits=8) = rint(k/255*256)/256 = (k + [k>=128])/256.
__global__ void __launch_bounds__(256)
u8_to_grid_kernel(const uint8_t* __restrict__ src, float* __restrict__ dst, int64_t n) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float v = __fdiv_rn((float)src[i], 255.0f);
        dst[i] = round_nbits(v, 256.0f, 0.00390625f);
    }
}

// Inverse of the above on the 256 reachable levels {0..127, 129..256}/256; anything else is
// reported (a decoded image that is not on the input grid means a corrupt stream or model drift).
__global__ void __launch_bounds__(256)
grid_to_u8_kernel(const float* __restrict__ src, uint8_t* __restrict__ dst, int64_t n,
                  int32_t* __restrict__ status_word) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int32_t bad = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float xs = src[i] * 256.0f;
        const int j = (int)rintf(xs);
        if ((float)j != xs || j < 0 || j > 256 || j == 128) bad = ST_OUT_OF_WINDOW;
        int k = j > 128 ? j - 1 : j;
        k = k < 0 ? 0 : (k > 255 ? 255 : k);
        dst[i] = (uint8_t)k;
    }
    if (bad) atomicOr(status_word, bad);
}

// Vector bodies.  uint8 -> float: four pixels per thread and step (a 4-byte load, a 16-byte store; both
// sides of a warp's access are contiguous), four steps in flight.  The 256 possible values of the
// quantisation are tabulated per CTA with the very formula of the scalar kernel, so the two agree by
// construction.  float -> uint8: 16 pixels per thread (four 16-byte loads, one 16-byte store).
__global__ void __launch_bounds__(256)
u8_to_grid_vec_kernel(const uint8_t* __restrict__ src, float* __restrict__ dst, int64_t n4) {
    __shared__ float s_lut[256];
    s_lut[threadIdx.x] = round_nbits(__fdiv_rn((float)threadIdx.x, 255.0f), 256.0f, 0.00390625f);
    __syncthreads();
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const uint32_t* s4 = reinterpret_cast<const uint32_t*>(src);
    float4* d4 = reinterpret_cast<float4*>(dst);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += 4 * stride) {
        uint32_t w[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) w[q] = i + q * stride < n4 ? __ldg(s4 + i + q * stride) : 0u;
#pragma unroll
        for (int q = 0; q < 4; ++q)
            if (i + q * stride < n4)
                d4[i + q * stride] = make_float4(s_lut[w[q] & 255u], s_lut[(w[q] >> 8) & 255u], s_lut[(w[q] >> 16) & 255u], s_lut[w[q] >> 24]);
    }
}

__global__ void __launch_bounds__(256)
grid_to_u8_vec_kernel(const float* __restrict__ src, uint8_t* __restrict__ dst, int64_t n16,
                      int32_t* __restrict__ status_word) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int32_t bad = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) {
        const float4* sp = reinterpret_cast<const float4*>(src) + 4 * i;
        uint32_t w[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float4 f = __ldg(sp + q);
            const float e[4] = {f.x, f.y, f.z, f.w};
            uint32_t word = 0;
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const float xs = e[t] * 256.0f;
                const int j = (int)rintf(xs);
                if ((float)j != xs || j < 0 || j > 256 || j == 128) bad = ST_OUT_OF_WINDOW;
                int k = j > 128 ? j - 1 : j;
                k = k < 0 ? 0 : (k > 255 ? 255 : k);
                word |= (uint32_t)k << (8 * t);
            }
            w[q] = word;
        }
        reinterpret_cast<uint4*>(dst)[i] = make_uint4(w[0], w[1], w[2], w[3]);
    }
    if (bad) atomicOr(status_word, bad);
}

cudaError_t launch_u8_to_grid(const uint8_t* src, float* dst, int64_t n, cudaStream_t stream) {
    if (n <= 0) return cudaSuccess;
    const int64_t cap = (int64_t)sm_count() * 8;
    int64_t done = 0;
    if ((uintptr_t)src % 16 == 0 && (uintptr_t)dst % 16 == 0 && n >= 16) {
        const int64_t n4 = n / 4;
        int64_t blocks = (n4 + 1023) / 1024;
        if (blocks > cap) blocks = cap;
        u8_to_grid_vec_kernel<<<(unsigned)blocks, 256, 0, stream>>>(src, dst, n4);
        done = n4 * 4;
    }
    if (done < n) {                                  // unaligned tensors, and the last pixels
        int64_t blocks = (n - done + 255) / 256;
        if (blocks > cap) blocks = cap;
        u8_to_grid_kernel<<<(unsigned)blocks, 256, 0, stream>>>(src + done, dst + done, n - done);
    }
    return cudaGetLastError();
}

cudaError_t launch_grid_to_u8(const float* src, uint8_t* dst, int64_t n, int32_t* status_word,
                              cudaStream_t stream) {
    if (n <= 0) return cudaSuccess;
    const int64_t cap = (int64_t)sm_count() * 8;
    int64_t done = 0;
    if ((uintptr_t)src % 16 == 0 && (uintptr_t)dst % 16 == 0 && n >= 16) {
        const int64_t n16 = n / 16;
        int64_t blocks = (n16 + 255) / 256;
        if (blocks > cap) blocks = cap;
        grid_to_u8_vec_kernel<<<(unsigned)blocks, 256, 0, stream>>>(src, dst, n16, status_word);
        done = n16 * 16;
    }
    if (done < n) {
        int64_t blocks = (n - done + 255) / 256;
        if (blocks > cap) blocks = cap;
        grid_to_u8_kernel<<<(unsigned)blocks, 256, 0, stream>>>(src + done, dst + done, n - done, status_word);
    }
    return cudaGetLastError();
}

}  // namespace flic
