// N1 -- the flow's pure index shuffles as gather kernels.
//
//   Permute.forward / .backward   invertible.py:38-48   NCHW->NHWC copy, F.linear with a dim x dim
//                                                       permutation matrix, NHWC->NCHW copy
//   ExtendDim.forward / .backward extenddim.py:23-37    space-to-depth / depth-to-space
//
// A permutation matrix applied by F.linear is a channel gather: out[:, i] = x[:, ids[i]]
// (P[i, ids[i]] = 1, invertible.py:34).  Both ops move every float exactly once:
// 4 B read + 4 B written per element, against the reference's three passes plus a
// C x C matmul per Permute.  Values are copied bit-for-bit (the matmul form turns -0.0 into +0.0;
// torch.equal does not distinguish them and every downstream op is sign-of-zero agnostic).
#include "flic_device.cuh"
#include "flic_kernels.cuh"

namespace flic {

// dst[b, i, p] = src[b, perm[i], p];  one thread per float4 (or float) of a channel plane.
template <int VEC>
__global__ void __launch_bounds__(256)
permute_channels_kernel(const float* __restrict__ src, float* __restrict__ dst,
                        const int32_t* __restrict__ perm, int64_t batch, int64_t channels, int64_t hw) {
    const int64_t per_plane = hw / VEC;
    const int64_t total = batch * channels * per_plane;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int64_t plane = i / per_plane;
        const int64_t k = (i - plane * per_plane) * VEC;
        const int64_t b = plane / channels;
        const int64_t c = plane - b * channels;
        const int64_t sc = __ldg(perm + c);
        const float* sp = src + (b * channels + sc) * hw + k;
        float* dp = dst + plane * hw + k;
        if (VEC == 4) *reinterpret_cast<float4*>(dp) = __ldg(reinterpret_cast<const float4*>(sp));
        else *dp = __ldg(sp);
    }
}

cudaError_t launch_permute_channels(const float* src, float* dst, const int32_t* perm, int64_t batch,
                                    int64_t channels, int64_t hw, cudaStream_t stream) {
    const int64_t n = batch * channels * hw;
    if (n <= 0) return cudaSuccess;
    const bool vec4 = (hw % 4 == 0) && ((uintptr_t)src % 16 == 0) && ((uintptr_t)dst % 16 == 0);
    const int64_t work = vec4 ? n / 4 : n;
    int64_t blocks = (work + 255) / 256;
    const int64_t cap = (int64_t)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    if (vec4) permute_channels_kernel<4><<<(unsigned)blocks, 256, 0, stream>>>(src, dst, perm, batch, channels, hw);
    else permute_channels_kernel<1><<<(unsigned)blocks, 256, 0, stream>>>(src, dst, perm, batch, channels, hw);
    return cudaGetLastError();
}

// forward (direction=+1):  dst[b, c*s*s + dy*s + dx, h, w] = src[b, c, h*s + dy, w*s + dx]
//   src (B, C, H, W) -> dst (B, C*s*s, H/s, W/s)                       extenddim.py:23-29
// backward (direction=-1): the inverse, src (B, C*s*s, H/s, W/s) -> dst (B, C, H, W)   :31-37
// One thread per element of the LARGE-plane tensor (the (B,C,H,W) side), so that side is
// coalesced; the other side is a stride-s access that the sectors of neighbouring threads cover.
__global__ void __launch_bounds__(256)
squeeze_kernel(const float* __restrict__ src, float* __restrict__ dst, int64_t batch, int64_t C,
               int64_t H, int64_t W, int s, int direction) {
    const int64_t total = batch * C * H * W;
    const int64_t h2 = H / s, w2 = W / s;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int64_t x = i % W;
        const int64_t y = (i / W) % H;
        const int64_t c = (i / (W * H)) % C;
        const int64_t b = i / (W * H * C);
        const int64_t dy = y % s, dx = x % s;
        const int64_t small = ((b * C * s * s + c * s * s + dy * s + dx) * h2 + y / s) * w2 + x / s;
        if (direction > 0) dst[small] = __ldg(src + i);
        else dst[i] = __ldg(src + small);
    }
}

// s == 2, W a multiple of 8, 16-byte aligned tensors: one thread per 8 consecutive floats of a row of
// the large tensor (two 16-byte accesses), which are 4 + 4 consecutive floats of the dx = 0 and
// dx = 1 planes of the small one (one 16-byte access each).  Every access is a full-width vector
// and every sector is touched once.
__global__ void __launch_bounds__(256)
squeeze2_vec_kernel(const float* __restrict__ src, float* __restrict__ dst, int64_t batch, int64_t C,
                    int64_t H, int64_t W, int direction) {
    const int64_t w8 = W >> 3;                       // threads per row of the large tensor
    const int64_t total = batch * C * H * w8;
    const int64_t h2 = H >> 1, w2 = W >> 1;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int64_t xq = i % w8;
        const int64_t row = i / w8;                  // (b C + c) H + y
        const int64_t y = row % H;
        const int64_t bc = row / H;                  // b C + c
        const int64_t big = row * W + xq * 8;
        // plane (b, c 4 + dy 2 + dx) of the small tensor, row y / 2, columns 4 xq .. 4 xq + 3
        const int64_t small0 = ((bc * 4 + (y & 1) * 2) * h2 + (y >> 1)) * w2 + xq * 4;
        const int64_t small1 = small0 + h2 * w2;
        if (direction > 0) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(src + big));
            const float4 b = __ldg(reinterpret_cast<const float4*>(src + big + 4));
            *reinterpret_cast<float4*>(dst + small0) = make_float4(a.x, a.z, b.x, b.z);
            *reinterpret_cast<float4*>(dst + small1) = make_float4(a.y, a.w, b.y, b.w);
        } else {
            const float4 e = __ldg(reinterpret_cast<const float4*>(src + small0));
            const float4 o = __ldg(reinterpret_cast<const float4*>(src + small1));
            *reinterpret_cast<float4*>(dst + big) = make_float4(e.x, o.x, e.y, o.y);
            *reinterpret_cast<float4*>(dst + big + 4) = make_float4(e.z, o.z, e.w, o.w);
        }
    }
}

cudaError_t launch_squeeze(const float* src, float* dst, int64_t batch, int64_t C, int64_t H,
                           int64_t W, int scale, int direction, cudaStream_t stream) {
    const int64_t n = batch * C * H * W;
    if (n <= 0) return cudaSuccess;
    const int64_t cap = (int64_t)sm_count() * 8;
    if (scale == 2 && W % 8 == 0 && H % 2 == 0 && (uintptr_t)src % 16 == 0 && (uintptr_t)dst % 16 == 0) {
        int64_t blocks = (n / 8 + 255) / 256;
        if (blocks > cap) blocks = cap;
        squeeze2_vec_kernel<<<(unsigned)blocks, 256, 0, stream>>>(src, dst, batch, C, H, W, direction);
        return cudaGetLastError();
    }
    int64_t blocks = (n + 255) / 256;
    if (blocks > cap) blocks = cap;
    squeeze_kernel<<<(unsigned)blocks, 256, 0, stream>>>(src, dst, batch, C, H, W, scale, direction);
    return cudaGetLastError();
}

}  // namespace flic
