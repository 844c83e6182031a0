// K3 -- rANS decode: renormalise, find the symbol, undo the state update; many streams.
//
// Replaces rans.decode (rans/rans.pyx:69-110).  The reference's caller hands it the word buffer
// and the mean/scale lists REVERSED and gets the symbols back reversed (trainer.py:317-318);
// here every stream is simply walked from its last symbol to its first, consuming its words
// from the last emitted to the first, and the symbols are stored in forward order.
//
// The reference finds the symbol with an 11-12 step binary search over the 2048-bin window,
// calling CDF() at every step (rans.pyx:96-104) and twice more for (start, freq) (:106-107).
// Because CDF(s) is non-decreasing in s any search that returns the smallest
// in-window s with CDF(s) > mod is bit-identical.  This kernel solves the continuous logistic
// model for s in float (two Newton steps, guess_symbol()), then evaluates the exact CDF at the
// guess and its left neighbour -- which are the (end, start) pair the state update needs anyway.
// On the reference's own test distributions that is 2.00 exact CDF evaluations per symbol
// instead of 13-14; a galloping / bisecting bracket takes over when the guess is off.
//
// Mapping: one lane per stream (every CDF evaluation depends on the stream's current state, so
// unlike the encoder nothing can be hoisted to helper warps); a warp stages
// [32 streams][8 symbols] tiles of (mean, scale) through shared memory with cp.async, and the
// decoded symbols go back through the same tile so the x store is coalesced too.  Neither the
// parameter loads nor the bitstream words (prefetched one renormalisation ahead) put global
// latency on the per-stream dependency chain.
#include "flic_device.cuh"
#include "flic_kernels.cuh"

namespace flic {

// Symbols per stream per tile.  Tiles are double-buffered: while the lanes decode tile q the
// parameters of tile q-1 arrive through cp.async (LDGSTS), so no global-load latency sits on the
// serial decode chain.  8 symbols x 2 buffers is 4.6 KB per warp: registers, not shared memory,
// limit residency.
constexpr int kDecTile = 8;
constexpr int kRowsPerPass = kLanes / kDecTile;  // tile rows one warp-wide copy covers

__device__ __forceinline__ void cp_async_f32(float* smem_dst, const float* gmem_src) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

#ifndef FLIC_DEC_MIN_BLOCKS
#define FLIC_DEC_MIN_BLOCKS 7
#endif
template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32, WARPS == 4 ? FLIC_DEC_MIN_BLOCKS : 16)
rans_decode_kernel(const uint32_t* __restrict__ packed, const int64_t* __restrict__ word_offsets,
                   const uint64_t* __restrict__ states, const float* __restrict__ mean,
                   const float* __restrict__ scale, const int64_t* __restrict__ offsets,
                   int64_t n_streams, float* __restrict__ x_out, uint64_t* __restrict__ end_states,
                   int32_t* __restrict__ status, int check_end) {
    __shared__ uint64_t s_tab[32];
    __shared__ float2 s_tile[WARPS][2][kLanes][kDecTile + 1];
    stage_exp_table(s_tab);

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int64_t first = ((int64_t)blockIdx.x * WARPS + warp) * kLanes;
    if (first >= n_streams) return;
    const int64_t stream = first + lane;
    const bool live = stream < n_streams;

    const int64_t beg = live ? offsets[stream] : 0;
    int64_t len = live ? offsets[stream + 1] - beg : 0;
    const int64_t wbeg = live ? word_offsets[stream] : 0;
    const int64_t wcount = live ? word_offsets[stream + 1] - wbeg : 0;
    const bool too_long = wcount > 0xffffffffll;
    if (too_long) len = 0;
    const int64_t max_len = warp_max_i64(len);
    // words are consumed from the last emitted to the first: wrem counts the unread ones, and the
    // next word to pull is kept in a register, loaded one renormalisation ahead
    const uint32_t* wp = packed + wbeg;
    uint32_t wrem = too_long ? 0u : (uint32_t)wcount;
    uint64_t state = live ? states[stream] : kRansL;
    uint32_t next_word = wrem ? __ldg(wp + (wrem - 1)) : 0u;
    int32_t flags = too_long ? ST_TOO_LONG : 0;
    const int sub = lane / kDecTile;   // which row of a pass this lane copies
    const int col = lane % kDecTile;   // symbol within the tile row

    // stage (mean, scale) of tile q into buffer q & 1: every 8-lane group copies one 32-byte row
    auto prefetch = [&](int64_t q) {
        float2(*tile)[kDecTile + 1] = s_tile[warp][q & 1];
        const int64_t i = q * kDecTile + col;
#pragma unroll
        for (int p = 0; p < kLanes / kRowsPerPass; ++p) {
            const int r = p * kRowsPerPass + sub;
            const int64_t b_r = shfl_i64(beg, r);
            const int64_t l_r = shfl_i64(len, r);
            if (i < l_r) {
                cp_async_f32(&tile[r][col].x, mean + b_r + i);
                cp_async_f32(&tile[r][col].y, scale + b_r + i);
            }
        }
        cp_async_commit();
    };

    const int64_t n_tiles = (max_len + kDecTile - 1) / kDecTile;
    if (n_tiles > 0) prefetch(n_tiles - 1);
    for (int64_t q = n_tiles - 1; q >= 0; --q) {
        if (q > 0) {
            prefetch(q - 1);
            cp_async_wait<1>();  // tile q has landed; tile q-1 may still be in flight
        } else {
            cp_async_wait<0>();
        }
        __syncwarp();
        float2(*tile)[kDecTile + 1] = s_tile[warp][q & 1];
        const int64_t t0 = q * kDecTile;
        // ---- lane-per-stream, last symbol of the tile first
        const int64_t rem = len - t0;
        const int cnt = rem >= kDecTile ? kDecTile : (rem > 0 ? (int)rem : 0);
#pragma unroll 1
        for (int j = kDecTile - 1; j >= 0; --j) {
            if (j < cnt) {
                if (state < kRansL) {  // rans.pyx:87-89
                    if (wrem) {
                        state = (state << 32) | next_word;
                        --wrem;
                        if (wrem) next_word = __ldg(wp + (wrem - 1));
                    } else {
                        flags |= ST_UNDERRUN;
                    }
                }
                const float2 ms = tile[lane][j];
                const int s = decode_symbol(state, ms.x, ms.y, s_tab, flags);
                tile[lane][j].x = (float)s * 0.00390625f;  // s / 256., exact
            }
        }
        __syncwarp();
        // ---- coalesced store of the decoded symbols (32-byte row segments)
#pragma unroll
        for (int p = 0; p < kLanes / kRowsPerPass; ++p) {
            const int r = p * kRowsPerPass + sub;
            const int64_t b_r = shfl_i64(beg, r);
            const int64_t l_r = shfl_i64(len, r);
            const int64_t i = t0 + col;
            if (i < l_r) x_out[b_r + i] = tile[r][col].x;
        }
        __syncwarp();  // buffer q & 1 is overwritten by the prefetch of tile q-2 next iteration
    }
    if (live) {
        if (check_end && !too_long && (state != kRansL || wrem != 0)) flags |= ST_BAD_END_STATE;
        end_states[stream] = state;
        status[stream] = flags;
    }
}

// ---- fast path: every lane stages its own stream ----------------------------------------------
// The kernel above moves (mean, scale) tiles with warp-wide copies, which costs four 64-bit
// shuffles and two address computations per 32 elements each way -- about 40 of its ~250
// instructions per symbol.  Here each lane copies 32-byte blocks of ITS OWN stream: two 16-byte
// cp.async per array per 8 symbols, and the 8 decoded symbols leave as two 16-byte stores.
// A 32-byte block is one DRAM sector, so the traffic is still exactly the algorithmic bytes.
//
// Blocks are cut on the 32-byte grid of the arrays' addresses: with shift = (address / 4) mod 8,
// block T of a lane holds the symbols i with (i + shift) >> 3 == T.  The launcher uses this kernel
// only when mean, scale and x_out share the same shift (true whenever they are allocations of
// their own, or equal slices of such).  A stream's first and last block may be partial; those
// go element by element.
#ifndef FLIC_LANE_BLOCK
#define FLIC_LANE_BLOCK 8
#endif
constexpr int kBlk = FLIC_LANE_BLOCK;   // symbols per lane per block: 8 (one 32-byte sector) or 4
constexpr int kBlkShift = kBlk == 8 ? 3 : 2;
// floats per lane row in shared memory: 8-symbol rows are padded to 48 B so that the 16-byte
// accesses of a quarter-warp fall on disjoint banks; 4-symbol rows (16 B) are conflict-free as is
constexpr int kBlkPitch = kBlk == 8 ? 12 : 4;

__device__ __forceinline__ void cp_async_16(float* smem_dst, const float* gmem_src) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}

template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32, WARPS == 4 ? FLIC_DEC_MIN_BLOCKS : 16)
rans_decode_lane_kernel(const uint32_t* __restrict__ packed, const int64_t* __restrict__ word_offsets,
                        const uint64_t* __restrict__ states, const float* __restrict__ mean,
                        const float* __restrict__ scale, const int64_t* __restrict__ offsets,
                        int64_t n_streams, float* __restrict__ x_out, uint64_t* __restrict__ end_states,
                        int32_t* __restrict__ status, int check_end, int shift) {
    __shared__ uint64_t s_tab[32];
    __shared__ __align__(16) float s_mean[WARPS][2][kLanes][kBlkPitch];
    __shared__ __align__(16) float s_scale[WARPS][2][kLanes][kBlkPitch];
    stage_exp_table(s_tab);

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int64_t first = ((int64_t)blockIdx.x * WARPS + warp) * kLanes;
    if (first >= n_streams) return;
    const int64_t stream = first + lane;
    const bool live = stream < n_streams;

    const int64_t beg = live ? offsets[stream] : 0;
    int64_t len = live ? offsets[stream + 1] - beg : 0;
    const int64_t wbeg = live ? word_offsets[stream] : 0;
    const int64_t wcount = live ? word_offsets[stream + 1] - wbeg : 0;
    const bool too_long = wcount > 0xffffffffll;
    if (too_long) len = 0;
    const int64_t end = beg + len;
    // blocks of this lane, last to first: T_hi, T_hi - 1, ..., T_lo
    const int64_t t_hi = (end - 1 + shift) >> kBlkShift, t_lo = (beg + shift) >> kBlkShift;
    const int64_t my_blocks = len > 0 ? t_hi - t_lo + 1 : 0;
    const int64_t n_iter = warp_max_i64(my_blocks);

    uint32_t wrem = too_long ? 0u : (uint32_t)wcount;
    const uint32_t* wptr = packed + wbeg + wrem;   // one past the word held in next_word
    uint64_t state = live ? states[stream] : kRansL;
    uint32_t next_word = wrem ? __ldg(wptr - 1) : 0u;
    int32_t flags = too_long ? ST_TOO_LONG : 0;

    // this lane's rows in the two buffers
    float* const row_mean = s_mean[warp][0][lane];
    float* const row_scale = s_scale[warp][0][lane];
    constexpr int kBufStride = kLanes * kBlkPitch;

    // stage block number q (counted from the stream's end) into buffer q & 1
    auto prefetch = [&](int64_t q) {
        if (q < my_blocks) {
            const int64_t i0 = ((t_hi - q) << kBlkShift) - shift;       // first symbol index of the block
            float* dm = row_mean + (int)(q & 1) * kBufStride;
            float* ds = row_scale + (int)(q & 1) * kBufStride;
            if (i0 >= beg && i0 + kBlk <= end) {                // whole block inside the stream
                cp_async_16(dm, mean + i0);
                cp_async_16(ds, scale + i0);
                if (kBlk == 8) {
                    cp_async_16(dm + 4, mean + i0 + 4);
                    cp_async_16(ds + 4, scale + i0 + 4);
                }
            } else {
#pragma unroll
                for (int j = 0; j < kBlk; ++j)
                    if (i0 + j >= beg && i0 + j < end) {
                        cp_async_f32(dm + j, mean + i0 + j);
                        cp_async_f32(ds + j, scale + i0 + j);
                    }
            }
        }
        cp_async_commit();
    };

    if (n_iter > 0) prefetch(0);
    for (int64_t q = 0; q < n_iter; ++q) {
        if (q + 1 < n_iter) {
            prefetch(q + 1);
            cp_async_wait<1>();   // block q has landed; block q+1 may still be in flight
        } else {
            cp_async_wait<0>();
        }
        // each lane reads only what it copied itself: no warp barrier needed
        if (q < my_blocks) {
            const int64_t i0 = ((t_hi - q) << kBlkShift) - shift;
            float* bm = row_mean + (int)(q & 1) * kBufStride;
            const float* bs = row_scale + (int)(q & 1) * kBufStride;
            const int j_lo = i0 >= beg ? 0 : (int)(beg - i0);
            const int j_hi = i0 + kBlk <= end ? kBlk : (int)(end - i0);
            auto pull = [&]() {  // rans.pyx:87-89
                if (state < kRansL) {
                    if (wrem) {
                        state = (state << 32) | next_word;
                        --wrem;
                        --wptr;
                        if (wrem) next_word = __ldg(wptr - 1);
                    } else {
                        flags |= ST_UNDERRUN;
                    }
                }
            };
            if (kBlk == 8 && j_lo == 0 && j_hi == kBlk) {
                // whole block: parameters come out of shared memory as 16-byte vectors (the row
                // pitch keeps those conflict-free), four symbols are decoded from registers, last
                // first, and leave as one 16-byte store
#pragma unroll 1
                for (int h = 1; h >= 0; --h) {
                    const float4 mv = *reinterpret_cast<const float4*>(bm + 4 * h);
                    const float4 sv = *reinterpret_cast<const float4*>(bs + 4 * h);
                    const float ms[4] = {mv.x, mv.y, mv.z, mv.w};
                    const float ss[4] = {sv.x, sv.y, sv.z, sv.w};
                    float xo[4];
#pragma unroll
                    for (int k = 3; k >= 0; --k) {
                        pull();
                        xo[k] = (float)decode_symbol(state, ms[k], ss[k], s_tab, flags) * 0.00390625f;  // s / 256., exact
                    }
                    const float4 o = make_float4(xo[0], xo[1], xo[2], xo[3]);
                    if (h == 1) {
                        // parked in the (now free) upper half of the mean row, so that the whole
                        // 32-byte sector is written at once
                        *reinterpret_cast<float4*>(bm + 4) = o;
                    } else {
                        *reinterpret_cast<float4*>(x_out + i0) = o;
                        *reinterpret_cast<float4*>(x_out + i0 + 4) = *reinterpret_cast<const float4*>(bm + 4);
                    }
                }
            } else {
                for (int j = j_hi - 1; j >= j_lo; --j) {
                    pull();
                    x_out[i0 + j] = (float)decode_symbol(state, bm[j], bs[j], s_tab, flags) * 0.00390625f;
                }
            }
        }
    }
    if (live) {
        if (check_end && !too_long && (state != kRansL || wrem != 0)) flags |= ST_BAD_END_STATE;
        end_states[stream] = state;
        status[stream] = flags;
    }
}

cudaError_t launch_rans_decode(const uint32_t* packed, const int64_t* word_offsets,
                               const uint64_t* states, const float* mean, const float* scale,
                               const int64_t* offsets, int64_t n_streams, float* x_out,
                               uint64_t* end_states, int32_t* status, int check_end,
                               cudaStream_t stream) {
    if (n_streams <= 0) return cudaSuccess;
    const int64_t warps = (n_streams + kLanes - 1) / kLanes;
    const bool small = warps <= (int64_t)sm_count() * 16;
    const int64_t blocks = (warps + kCoderWarps - 1) / kCoderWarps;
    // the lane-staged kernel needs the three symbol arrays on the same 32-byte phase
    const int sh_m = (int)(((uintptr_t)mean >> 2) & (kBlk - 1)), sh_s = (int)(((uintptr_t)scale >> 2) & (kBlk - 1));
    const int sh_x = (int)(((uintptr_t)x_out >> 2) & (kBlk - 1));
    const bool lane_staged = sh_m == sh_s && sh_m == sh_x && ((uintptr_t)mean & 3) == 0 &&
                             ((uintptr_t)scale & 3) == 0 && ((uintptr_t)x_out & 3) == 0;
    note_coder_kernel(1, lane_staged ? "rans_decode_lane_kernel" : "rans_decode_kernel");
    if (lane_staged) {
        if (small)
            rans_decode_lane_kernel<1><<<(unsigned)warps, 32, 0, stream>>>(
                packed, word_offsets, states, mean, scale, offsets, n_streams, x_out, end_states, status, check_end, sh_m);
        else
            rans_decode_lane_kernel<kCoderWarps><<<(unsigned)blocks, kCoderWarps * 32, 0, stream>>>(
                packed, word_offsets, states, mean, scale, offsets, n_streams, x_out, end_states, status, check_end, sh_m);
    } else if (small) {
        rans_decode_kernel<1><<<(unsigned)warps, 32, 0, stream>>>(
            packed, word_offsets, states, mean, scale, offsets, n_streams, x_out, end_states, status, check_end);
    } else {
        rans_decode_kernel<kCoderWarps><<<(unsigned)blocks, kCoderWarps * 32, 0, stream>>>(
            packed, word_offsets, states, mean, scale, offsets, n_streams, x_out, end_states, status, check_end);
    }
    return cudaGetLastError();
}

}  // namespace flic
