// flic_kernels.cuh -- launch-side declarations shared by the .cu files and the C-ABI (capi.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace flic {

// Streams per warp / symbols per stream per tile in the coder kernels.  One lane owns one
// stream; a warp stages a 32-stream x 32-symbol tile through shared memory so that global
// traffic is coalesced (128 B per stream row) while the rANS recurrence runs lane-per-stream.
constexpr int kLanes = 32;
constexpr int kTile = 32;
constexpr int kCoderWarps = 4;  // warps per CTA in encode/decode kernels

// Name of the coder kernel the last launch_rans_encode (which = 0) / launch_rans_decode (1) chose.
const char* last_coder_kernel(int which);
void note_coder_kernel(int which, const char* name);
// Decode kernel choice: -1 by stream count, 0 lane-per-stream, 1 CTA-per-stream, 2 / 4 / 8 a cluster of
// that many CTAs per stream.  Returns the old value.
int set_decode_kernel(int which);

// K1  (x, mean, scale) -> (start, freq)                     rans/rans.pyx:49-56
cudaError_t launch_cdf_tables(const float* x, const float* mean, const float* scale, int64_t n,
                              uint32_t* start, uint32_t* freq, int32_t* status_word,
                              cudaStream_t stream);

cudaError_t launch_debug_expf(const float* x, float* y, int64_t n, cudaStream_t stream);
cudaError_t launch_debug_part1(const float* arg, int32_t* y, int64_t n, cudaStream_t stream);
cudaError_t launch_debug_div_check(int64_t n, uint64_t seed, int mode, unsigned long long* mismatches, cudaStream_t stream);
cudaError_t launch_debug_push_check(int64_t n, uint64_t seed, unsigned long long* mismatches, cudaStream_t stream);

// K1+K2 fused: per-stream rANS encode into worst-case scratch regions.
//   scratch[offsets[s] .. offsets[s] + counts[s])  = words of stream s in emission order
cudaError_t launch_rans_encode(const float* x, const float* mean, const float* scale,
                               const int64_t* offsets, int64_t n_streams,
                               const uint64_t* init_states, uint32_t* scratch, int64_t* counts,
                               uint64_t* states, int32_t* status, cudaStream_t stream);

// Copy every stream's words to a destination of its own: dst[dst_starts[s] + i] = src[src_offsets[s] + i]
// for i < src_offsets[s + 1] - src_offsets[s] (concatenating the levels of a chained stream).
cudaError_t launch_gather_words(const uint32_t* src, const int64_t* src_offsets, const int64_t* dst_starts,
                                int64_t n_streams, uint32_t* dst, int64_t dst_capacity, int32_t* status,
                                cudaStream_t stream);

// dst[i] = src[n - 1 - i] (32-bit elements; src != dst)
cudaError_t launch_reverse_u32(const uint32_t* src, uint32_t* dst, int64_t n, cudaStream_t stream);

// K4: exclusive scan of counts -> word_offsets[n_streams + 1]; gather scratch -> packed.
// scan_tmp needs scan_tmp_elems(n_streams) int64 elements.
int64_t scan_tmp_elems(int64_t n_streams);
cudaError_t launch_scan_counts(const int64_t* counts, int64_t n_streams, int64_t* word_offsets,
                               int64_t* scan_tmp, cudaStream_t stream);
cudaError_t launch_pack_words(const uint32_t* scratch, const int64_t* offsets,
                              const int64_t* word_offsets, int64_t n_streams, uint32_t* packed,
                              int64_t packed_capacity, int32_t* status, cudaStream_t stream);

// Continuation of a decode across calls (coder.py:29-38 chains one state through the levels):
// stream s has in[s] unread words at packed[word_offsets[s] ...] when the call starts (null: all
// of word_offsets[s + 1] - word_offsets[s]) and out[s] when it returns (null: not reported).
struct WordsLeft {
    const int64_t* in;
    int64_t* out;
};

// K3: per-stream rANS decode (search + state update), symbols written in forward order.
cudaError_t launch_rans_decode(const uint32_t* packed, const int64_t* word_offsets,
                               const uint64_t* states, const float* mean, const float* scale,
                               const int64_t* offsets, int64_t n_streams, float* x_out,
                               uint64_t* end_states, int32_t* status, int check_end,
                               WordsLeft left, cudaStream_t stream);

// K3c / K3d: the same for few streams -- one CTA (cluster = 1) or one thread-block cluster of 2, 4
// or 8 CTAs per stream, exact CDF windows tabulated ahead of the serial chain by producer warps
// (rans_decode_coop.cu).  launch_rans_decode picks kernel and cluster size by stream count.
cudaError_t launch_rans_decode_coop(const uint32_t* packed, const int64_t* word_offsets,
                                    const uint64_t* states, const float* mean, const float* scale,
                                    const int64_t* offsets, int64_t n_streams, float* x_out,
                                    uint64_t* end_states, int32_t* status, int check_end,
                                    WordsLeft left, int cluster, cudaStream_t stream);
// Clusters of `cluster` (2, 4, 8) CTAs of that kernel the current device holds at once; 0: none.
int64_t coop_cluster_capacity(int cluster);

// K5: x[:, a_ch:, :, :] += sign * Round_nbits(t)            couplelib.py:49-51,58-59; roundlib.py:18-38
//   x: (batch, channels, hw) contiguous;  t: (batch, channels - a_ch, hw) contiguous
cudaError_t launch_couple_add_round(float* x, const float* t, int64_t batch, int64_t channels,
                                    int64_t a_ch, int64_t hw, float sign, int nbits,
                                    cudaStream_t stream);

// uint8 pixels -> grid floats (trainer.py:61,72) and back.
cudaError_t launch_u8_to_grid(const uint8_t* src, float* dst, int64_t n, cudaStream_t stream);
cudaError_t launch_grid_to_u8(const float* src, uint8_t* dst, int64_t n, int32_t* status_word,
                              cudaStream_t stream);

// N1: channel permutation (invertible.py:38-48) and space-to-depth (extenddim.py:23-37) as gathers.
cudaError_t launch_permute_channels(const float* src, float* dst, const int32_t* perm, int64_t batch,
                                    int64_t channels, int64_t hw, cudaStream_t stream);
cudaError_t launch_squeeze(const float* src, float* dst, int64_t batch, int64_t C, int64_t H,
                           int64_t W, int scale, int direction, cudaStream_t stream);

// N4: DLogistic.log_prob with the per-image sum fused (distlib.py:40-55, flows.py:154-169) and
// DLogistic.sample from given uniforms (distlib.py:57-70).  logp_out / sum_out may be null.
cudaError_t launch_dlogistic_log_prob(const float* x, const float* mean, const float* logscale,
                                      int64_t batch, int64_t per_item, int nbits, float eps,
                                      float* logp_out, float* sum_out, cudaStream_t stream);
cudaError_t launch_dlogistic_sample(const float* u, const float* mean, const float* logscale, int64_t n,
                                    int nbits, float* out, cudaStream_t stream);

}  // namespace flic
