// K1+K2 fused -- table evaluation and 64-bit-state / 32-bit-word rANS encode, many streams.
//
// Replaces rans.encode (rans/rans.pyx:37-67): pass 1 (per-symbol start/freq, :49-56) and pass 2
// (the serial renorm-then-update recurrence, :58-67), for every stream of a partition at once.
// Each stream starts at state 1<<32 like the reference's call site (trainer.py:310) and its
// words come out in emission order, so (final_state, words) per stream is what the reference's
// encode() returns for the same slice.
//
// Mapping.  One lane owns one stream (the recurrence is serial per stream); a warp owns 32
// consecutive streams and walks them in tiles of 32 symbols:
//   phase A  for each of the 32 streams in turn the whole warp loads 32 consecutive symbols
//            (x, mean, scale: three coalesced 128-byte rows), evaluates (start, freq) -- the
//            expensive FP64 part, perfectly parallel -- and parks them in shared memory;
//   phase B  every lane replays its own stream's 32 table entries through the rANS recurrence,
//            state in registers, and appends emitted words to its stream's scratch region.
// The (start, freq) tables therefore never touch HBM: algorithmic traffic is the 12 B/symbol of
// inputs plus the coded bits.  Shared memory holds a [32][33] uint2 tile per warp (pitch 33 so
// that phase B's lane-per-row reads are bank-conflict free) and the 256-byte exp2 table.
#include "flic_device.cuh"
#include "flic_kernels.cuh"

namespace flic {

template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
rans_encode_kernel(const float* __restrict__ x, const float* __restrict__ mean,
                   const float* __restrict__ scale, const int64_t* __restrict__ offsets,
                   int64_t n_streams, const uint64_t* __restrict__ init_states,
                   uint32_t* __restrict__ scratch, int64_t* __restrict__ counts,
                   uint64_t* __restrict__ states, int32_t* __restrict__ status) {
    __shared__ uint64_t s_tab[32];
    __shared__ uint2 s_tile[WARPS][kLanes][kTile + 1];
    stage_exp_table(s_tab);

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int64_t first = ((int64_t)blockIdx.x * WARPS + warp) * kLanes;
    if (first >= n_streams) return;  // whole warp leaves together
    const int64_t stream = first + lane;
    const bool live = stream < n_streams;

    const int64_t beg = live ? offsets[stream] : 0;
    const int64_t len = live ? offsets[stream + 1] - beg : 0;
    const int64_t max_len = warp_max_i64(len);

    uint64_t state = (live && init_states) ? init_states[stream] : kRansL;
    int64_t wpos = beg;  // scratch region of this stream starts at its first symbol index
    int32_t flags = 0;
    uint2(*tile)[kTile + 1] = s_tile[warp];

    for (int64_t t0 = 0; t0 < max_len; t0 += kTile) {
        // ---- phase A: tables for a 32-stream x 32-symbol tile, one coalesced row at a time
#pragma unroll 1
        for (int r = 0; r < kLanes; ++r) {
            const int64_t b_r = shfl_i64(beg, r);
            const int64_t l_r = shfl_i64(len, r);
            if (t0 >= l_r) continue;  // warp-uniform
            const int64_t i = t0 + lane;
            int32_t f = 0;
            if (i < l_r) {
                const int64_t g = b_r + i;
                const SymbolTable e = make_table(__ldg(x + g), __ldg(mean + g), __ldg(scale + g), s_tab, f);
                tile[r][lane] = make_uint2(e.start, e.freq);
            }
            if (__any_sync(0xffffffffu, f != 0)) {  // rare: hand the row's status to its owner lane
                const int32_t row_flags = __reduce_or_sync(0xffffffffu, (unsigned)f);
                if (lane == r) flags |= row_flags;
            }
        }
        __syncwarp();
        // ---- phase B: lane-per-stream recurrence over the tile
        const int64_t rem = len - t0;
        const int cnt = rem >= kTile ? kTile : (rem > 0 ? (int)rem : 0);
#pragma unroll 4
        for (int j = 0; j < kTile; ++j) {
            if (j < cnt) {
                const uint2 e = tile[lane][j];
                uint32_t word;
                if (rans_push(state, e.x, e.y, word)) scratch[wpos++] = word;
            }
        }
        __syncwarp();
    }
    if (live) {
        counts[stream] = wpos - beg;
        states[stream] = state;
        status[stream] = flags;
    }
}

cudaError_t launch_rans_encode(const float* x, const float* mean, const float* scale,
                               const int64_t* offsets, int64_t n_streams,
                               const uint64_t* init_states, uint32_t* scratch, int64_t* counts, uint64_t* states, int32_t* status,
                               cudaStream_t stream) {
    if (n_streams <= 0) return cudaSuccess;
    // Few streams: one warp per CTA so that the warps spread over all SMs; otherwise 4 warps per
    // CTA (34 KB of shared memory each, 6 CTAs resident per SM).
    const int64_t warps = (n_streams + kLanes - 1) / kLanes;
    if (warps <= (int64_t)sm_count() * 8) {
        rans_encode_kernel<1><<<(unsigned)warps, 32, 0, stream>>>(
            x, mean, scale, offsets, n_streams, init_states, scratch, counts, states, status);
    } else {
        const int64_t blocks = (warps + kCoderWarps - 1) / kCoderWarps;
        rans_encode_kernel<kCoderWarps><<<(unsigned)blocks, kCoderWarps * 32, 0, stream>>>(
            x, mean, scale, offsets, n_streams, init_states, scratch, counts, states, status);
    }
    return cudaGetLastError();
}

}  // namespace flic
