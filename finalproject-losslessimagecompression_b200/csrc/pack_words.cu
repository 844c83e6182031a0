// K4 -- per-stream word counts -> exclusive prefix sum -> dense concatenation of the streams.
//
// The reference appends each level's word list in Python (trainer.py:315,321); with many
// independent streams the equivalent is a scan over the per-stream counts followed by a gather
// from the encoder's worst-case scratch regions into one dense buffer, so that a compressed
// batch is (word_offsets, final_states, packed words).  Traffic is ~b/8 bytes per symbol each
// way (b = coded bits/symbol), a few percent of the encoder's.
//
// Scan: three small launches (per-block totals, scan of the totals by one CTA, per-block
// exclusive scan + carry).  n_streams is at most a few million, so this is launch-latency sized.
#include "flic_device.cuh"
#include "flic_kernels.cuh"

namespace flic {

constexpr int kScanBlock = 1024;

__device__ __forceinline__ int64_t warp_inclusive_scan(int64_t v, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int64_t o = shfl_i64(v, lane - d < 0 ? lane : lane - d);
        if (lane >= d) v += o;
    }
    return v;
}

// Inclusive scan across a 1024-thread CTA; returns this thread's inclusive value, *total the sum.
__device__ __forceinline__ int64_t block_inclusive_scan(int64_t v, int64_t* s_warp, int64_t* total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    v = warp_inclusive_scan(v, lane);
    if (lane == 31) s_warp[warp] = v;
    __syncthreads();
    if (warp == 0) {
        int64_t w = s_warp[lane];
        w = warp_inclusive_scan(w, lane);
        s_warp[lane] = w;
    }
    __syncthreads();
    const int64_t carry = warp ? s_warp[warp - 1] : 0;
    *total = s_warp[31];
    __syncthreads();
    return v + carry;
}

__global__ void __launch_bounds__(kScanBlock)
scan_block_totals_kernel(const int64_t* __restrict__ counts, int64_t n, int64_t* __restrict__ block_totals) {
    __shared__ int64_t s_warp[32];
    const int64_t i = (int64_t)blockIdx.x * kScanBlock + threadIdx.x;
    int64_t total;
    block_inclusive_scan(i < n ? counts[i] : 0, s_warp, &total);
    if (threadIdx.x == 0) block_totals[blockIdx.x] = total;
}

// One CTA: exclusive scan of block_totals in place (n_blocks may exceed 1024: chunked with carry).
__global__ void __launch_bounds__(kScanBlock)
scan_totals_kernel(int64_t* __restrict__ block_totals, int64_t n_blocks) {
    __shared__ int64_t s_warp[32];
    int64_t carry = 0;
    for (int64_t base = 0; base < n_blocks; base += kScanBlock) {
        const int64_t i = base + threadIdx.x;
        const int64_t v = i < n_blocks ? block_totals[i] : 0;
        int64_t total;
        const int64_t inc = block_inclusive_scan(v, s_warp, &total);
        if (i < n_blocks) block_totals[i] = carry + inc - v;
        carry += total;
    }
}

__global__ void __launch_bounds__(kScanBlock)
scan_apply_kernel(const int64_t* __restrict__ counts, int64_t n, const int64_t* __restrict__ block_totals,
                  int64_t* __restrict__ word_offsets) {
    __shared__ int64_t s_warp[32];
    const int64_t i = (int64_t)blockIdx.x * kScanBlock + threadIdx.x;
    const int64_t v = i < n ? counts[i] : 0;
    int64_t total;
    const int64_t inc = block_inclusive_scan(v, s_warp, &total);
    const int64_t base = block_totals[blockIdx.x];
    if (i < n) {
        word_offsets[i] = base + inc - v;
        if (i == n - 1) word_offsets[n] = base + inc;
    }
}

int64_t scan_tmp_elems(int64_t n_streams) {
    return (n_streams + kScanBlock - 1) / kScanBlock + 1;
}

cudaError_t launch_scan_counts(const int64_t* counts, int64_t n_streams, int64_t* word_offsets,
                               int64_t* scan_tmp, cudaStream_t stream) {
    if (n_streams <= 0) {
        return cudaMemsetAsync(word_offsets, 0, sizeof(int64_t), stream);
    }
    const int64_t blocks = (n_streams + kScanBlock - 1) / kScanBlock;
    scan_block_totals_kernel<<<(unsigned)blocks, kScanBlock, 0, stream>>>(counts, n_streams, scan_tmp);
    scan_totals_kernel<<<1, kScanBlock, 0, stream>>>(scan_tmp, blocks);
    scan_apply_kernel<<<(unsigned)blocks, kScanBlock, 0, stream>>>(counts, n_streams, scan_tmp, word_offsets);
    return cudaGetLastError();
}

// Gather: one warp per stream copies its words from scratch[offsets[s] ...] to
// packed[word_offsets[s] ...]; consecutive lanes move consecutive words.
__global__ void __launch_bounds__(256)
pack_words_kernel(const uint32_t* __restrict__ scratch, const int64_t* __restrict__ offsets,
                  const int64_t* __restrict__ word_offsets, int64_t n_streams,
                  uint32_t* __restrict__ packed, int64_t capacity, int32_t* __restrict__ status) {
    const int lane = threadIdx.x & 31;
    const int64_t warps_total = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t s = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); s < n_streams; s += warps_total) {
        const int64_t src = offsets[s];
        const int64_t dst = word_offsets[s];
        const int64_t cnt = word_offsets[s + 1] - dst;
        if (dst + cnt > capacity) {  // caller's buffer too small: report, copy nothing
            if (lane == 0) atomicOr(status + s, ST_UNDERRUN);
            continue;
        }
        for (int64_t k = lane; k < cnt; k += 32) packed[dst + k] = scratch[src + k];
    }
}

cudaError_t launch_pack_words(const uint32_t* scratch, const int64_t* offsets,
                              const int64_t* word_offsets, int64_t n_streams, uint32_t* packed,
                              int64_t packed_capacity, int32_t* status, cudaStream_t stream) {
    if (n_streams <= 0) return cudaSuccess;
    const int warps_per_cta = 8;
    int64_t blocks = (n_streams + warps_per_cta - 1) / warps_per_cta;
    const int64_t cap = (int64_t)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    pack_words_kernel<<<(unsigned)blocks, warps_per_cta * 32, 0, stream>>>(
        scratch, offsets, word_offsets, n_streams, packed, packed_capacity, status);
    return cudaGetLastError();
}

// One warp per stream, like pack_words_kernel, with the destination given per stream.
__global__ void __launch_bounds__(256)
gather_words_kernel(const uint32_t* __restrict__ src, const int64_t* __restrict__ src_offsets,
                    const int64_t* __restrict__ dst_starts, int64_t n_streams, uint32_t* __restrict__ dst,
                    int64_t capacity, int32_t* __restrict__ status) {
    const int lane = threadIdx.x & 31;
    const int64_t warps_total = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t s = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); s < n_streams; s += warps_total) {
        const int64_t a = src_offsets[s], cnt = src_offsets[s + 1] - a, d = dst_starts[s];
        if (d < 0 || d + cnt > capacity) {
            if (lane == 0 && status) atomicOr(status + s, ST_UNDERRUN);
            continue;
        }
        for (int64_t k = lane; k < cnt; k += 32) dst[d + k] = src[a + k];
    }
}

cudaError_t launch_gather_words(const uint32_t* src, const int64_t* src_offsets, const int64_t* dst_starts,
                                int64_t n_streams, uint32_t* dst, int64_t dst_capacity, int32_t* status,
                                cudaStream_t stream) {
    if (n_streams <= 0) return cudaSuccess;
    const int warps_per_cta = 8;
    int64_t blocks = (n_streams + warps_per_cta - 1) / warps_per_cta;
    const int64_t cap = (int64_t)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    gather_words_kernel<<<(unsigned)blocks, warps_per_cta * 32, 0, stream>>>(src, src_offsets, dst_starts, n_streams,
                                                                           dst, dst_capacity, status);
    return cudaGetLastError();
}

// dst[i] = src[n - 1 - i]: the reference's callers hand decode() its arrays reversed and get the
// message back reversed (trainer.py:317-318); the single-stream drop-in undoes that on the device.
__global__ void __launch_bounds__(256)
reverse_u32_kernel(const uint32_t* __restrict__ src, uint32_t* __restrict__ dst, int64_t n) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = src[n - 1 - i];
}

cudaError_t launch_reverse_u32(const uint32_t* src, uint32_t* dst, int64_t n, cudaStream_t stream) {
    if (n <= 0) return cudaSuccess;
    int64_t blocks = (n + 255) / 256;
    const int64_t cap = (int64_t)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    reverse_u32_kernel<<<(unsigned)blocks, 256, 0, stream>>>(src, dst, n);
    return cudaGetLastError();
}

}  // namespace flic
