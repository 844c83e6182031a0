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

#include <atomic>
#include <cstdint>
#include <cstdlib>

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
                   int32_t* __restrict__ status, int check_end, WordsLeft left) {
    __shared__ __align__(256) uint64_t s_tab[32];
    __shared__ float2 s_tile[WARPS][2][kLanes][kDecTile + 1];
    const ExpTab tab = stage_exp_table(s_tab);

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int64_t first = ((int64_t)blockIdx.x * WARPS + warp) * kLanes;
    if (first >= n_streams) return;
    const int64_t stream = first + lane;
    const bool live = stream < n_streams;

    const int64_t beg = live ? offsets[stream] : 0;
    int64_t len = live ? offsets[stream + 1] - beg : 0;
    const int64_t wbeg = live ? word_offsets[stream] : 0;
    int64_t wcount = live ? word_offsets[stream + 1] - wbeg : 0;
    if (live && left.in) wcount = left.in[stream] < 0 ? 0 : (left.in[stream] < wcount ? left.in[stream] : wcount);
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
                const int s = decode_symbol(state, ms.x, ms.y, tab, flags);
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
        if (left.out) left.out[stream] = (int64_t)wrem;
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

// Bookkeeping is per lane and 32-bit: a lane's blocks are numbered q = 0 (the stream's last
// symbols) .. nb - 1 (its first); only block 0 and block nb - 1 can be partial, every other one is
// a whole 32-byte sector per array, so the loop carries three running pointers and a counter
// instead of 64-bit symbol indices.  The word pull is predicated (select + conditional load, no
// divergent branch): with 32 lanes, some lane renormalises at nearly every symbol, and a branch
// would make the whole warp walk its body each time.
template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32, WARPS == 4 ? FLIC_DEC_MIN_BLOCKS : 16)
rans_decode_lane_kernel(const uint32_t* __restrict__ packed, const int64_t* __restrict__ word_offsets,
                        const uint64_t* __restrict__ states, const float* __restrict__ mean,
                        const float* __restrict__ scale, const int64_t* __restrict__ offsets,
                        int64_t n_streams, float* __restrict__ x_out, uint64_t* __restrict__ end_states,
                        int32_t* __restrict__ status, int check_end, WordsLeft left, int shift) {
    __shared__ __align__(256) uint64_t s_tab[32];
    // [buffer][mean, scale][lane][kBlkPitch]
    __shared__ __align__(16) float s_par[WARPS][2][2][kLanes][kBlkPitch];
    const ExpTab tab = stage_exp_table(s_tab);

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int64_t first = ((int64_t)blockIdx.x * WARPS + warp) * kLanes;
    if (first >= n_streams) return;
    const int64_t stream = first + lane;
    const bool live = stream < n_streams;

    const int64_t beg = live ? offsets[stream] : 0;
    int64_t len = live ? offsets[stream + 1] - beg : 0;
    const int64_t wbeg = live ? word_offsets[stream] : 0;
    int64_t wcount = live ? word_offsets[stream + 1] - wbeg : 0;
    if (live && left.in) wcount = left.in[stream] < 0 ? 0 : (left.in[stream] < wcount ? left.in[stream] : wcount);
    // 32-bit counters: a single stream of 2^31 symbols (or words) is not supported
    const bool too_long = wcount > 0x7fffffffll || len > 0x7fffffffll;
    if (too_long) len = 0;
    const int64_t end = beg + len;
    const int64_t t_hi = (end - 1 + shift) >> kBlkShift, t_lo = (beg + shift) >> kBlkShift;
    const int nb = len > 0 ? (int)(t_hi - t_lo + 1) : 0;
    int n_iter = nb;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) n_iter = max(n_iter, __shfl_xor_sync(0xffffffffu, n_iter, d));
    // symbols [j_head, kBlk) of block nb - 1 and [0, j_tail) of block 0 belong to the stream
    const int j_head = (int)((beg + shift) & (kBlk - 1));
    const int j_tail = (int)((end - 1 + shift) & (kBlk - 1)) + 1;
    const int j_ends = j_head | (j_tail << 4);

    const uint32_t* const wbase = packed + wbeg;
    int wrem = too_long ? 0 : (int)wcount;          // unread words; negative once the stream has under-run
    const uint64_t st0 = live ? states[stream] : kRansL;
    uint32_t hi = (uint32_t)(st0 >> 32), lo = (uint32_t)st0;
    uint32_t next_word = wrem > 0 ? __ldg(wbase + (wrem - 1)) : 0u;
    int32_t flags = too_long ? ST_TOO_LONG : 0;
    ParamGuard guard = guard_init();

    // element index of the first slot of block q is i_stage (while staging) / i_dec (while decoding);
    // the three arrays share the phase `shift`, so one running index serves them all
    int64_t i_stage = (t_hi << kBlkShift) - shift;
    int64_t i_dec = i_stage;
    float* const row = s_par[warp][0][0][lane];
    constexpr int kArr = kLanes * kBlkPitch, kBuf = 2 * kArr;

    auto stage = [&](int q) {   // (mean, scale) of block q -> buffer q & 1
        if (q < nb) {
            float* d = row + (q & 1) * kBuf;
            const float* pm = mean + i_stage;
            const float* ps = scale + i_stage;
            const int jl = q == nb - 1 ? (j_ends & 15) : 0, jh = q == 0 ? (j_ends >> 4) : kBlk;
            if (jl == 0 && jh == kBlk) {
                cp_async_16(d, pm);
                cp_async_16(d + kArr, ps);
                if (kBlk == 8) {
                    cp_async_16(d + 4, pm + 4);
                    cp_async_16(d + kArr + 4, ps + 4);
                }
            } else {
#pragma unroll
                for (int j = 0; j < kBlk; ++j)
                    if (j >= jl && j < jh) {
                        cp_async_f32(d + j, pm + j);
                        cp_async_f32(d + kArr + j, ps + j);
                    }
            }
            i_stage -= kBlk;
        }
        cp_async_commit();
    };
    auto pull = [&]() {  // rans.pyx:87-89
        // state < 2^32 iff the high word is zero: then (hi, lo) <- (lo, next_word), and
        // { --wrem; if (wrem > 0) next_word = wbase[wrem - 1]; }, all predicated: no divergent
        // branch (with 32 lanes, some lane renormalises at nearly every symbol), one comparison
        // serving the selects and the decrement; the address is formed unconditionally (two
        // shift-adds against selects under a predicate -- a wide multiply-add compiles to four
        // instructions -- and zero extension is as good as sign extension for an address that is
        // only used when wrem > 0)
        asm volatile("{\n\t.reg .pred p, q;\n\t.reg .u64 a;\n\t"
                     "setp.eq.u32 p, %0, 0;\n\t"
                     "selp.b32 %0, %1, %0, p;\n\t"
                     "selp.b32 %1, %2, %1, p;\n\t"
                     "@p add.s32 %3, %3, -1;\n\t"
                     "setp.gt.and.s32 q, %3, 0, p;\n\t"
                     "cvt.u64.u32 a, %3;\n\t"
                     "shl.b64 a, a, 2;\n\t"
                     "add.s64 a, a, %4;\n\t"
                     "@q ld.global.nc.u32 %2, [a+-4];\n\t}"
                     : "+r"(hi), "+r"(lo), "+r"(next_word), "+r"(wrem) : "l"(wbase));
    };

    if (n_iter > 0) stage(0);
    for (int q = 0; q < n_iter; ++q) {
        if (q + 1 < n_iter) {
            stage(q + 1);
            cp_async_wait<1>();   // block q has landed; block q+1 may still be in flight
        } else {
            cp_async_wait<0>();
        }
        // each lane reads only what it copied itself: no warp barrier needed
        if (q < nb) {
            float* bm = row + (q & 1) * kBuf;
            const float* bs = bm + kArr;
            float* px = x_out + i_dec;
            const int jl = q == nb - 1 ? (j_ends & 15) : 0, jh = q == 0 ? (j_ends >> 4) : kBlk;
            if (kBlk == 8 && jl == 0 && jh == kBlk) {
                // whole block: parameters come out of shared memory as 16-byte vectors (the row
                // pitch keeps those conflict-free), four symbols at a time are decoded from
                // registers, last first; the upper four are parked in the (now free) upper half
                // of the mean row so that the whole 32-byte sector is written at once
#pragma unroll
                for (int h = 1; h >= 0; --h) {
                    const float4 mv = *reinterpret_cast<const float4*>(bm + 4 * h);
                    const float4 sv = *reinterpret_cast<const float4*>(bs + 4 * h);
                    const float ms[4] = {mv.x, mv.y, mv.z, mv.w};
                    const float ss[4] = {sv.x, sv.y, sv.z, sv.w};
                    float xo[4];
#pragma unroll
                    for (int k = 3; k >= 0; --k) {
                        pull();
                        xo[k] = (float)decode_symbol_lean(hi, lo, ms[k], ss[k], tab, guard, flags) * 0.00390625f;  // s / 256., exact
                    }
                    const float4 o = make_float4(xo[0], xo[1], xo[2], xo[3]);
                    if (h == 1) {
                        *reinterpret_cast<float4*>(bm + 4) = o;
                    } else {
                        *reinterpret_cast<float4*>(px) = o;
                        *reinterpret_cast<float4*>(px + 4) = *reinterpret_cast<const float4*>(bm + 4);
                    }
                }
            } else {
                for (int j = jh - 1; j >= jl; --j) {
                    pull();
                    px[j] = (float)decode_symbol_lean(hi, lo, bm[j], bs[j], tab, guard, flags) * 0.00390625f;
                }
            }
            i_dec -= kBlk;
        }
    }
    if (live) {
        flags |= guard_flags(guard);
        if (wrem < 0) flags |= ST_UNDERRUN;
        const uint64_t state = ((uint64_t)hi << 32) | lo;
        if (check_end && !too_long && (state != kRansL || wrem != 0)) flags |= ST_BAD_END_STATE;
        end_states[stream] = state;
        status[stream] = flags;
        if (left.out) left.out[stream] = wrem > 0 ? (int64_t)wrem : 0;
    }
}

// Kernel choice by stream count.
//   * up to the number of 8-, 4-, 2-CTA clusters the device holds at once: a cluster per stream
//     (rans_decode_coop.cu, K3d) -- the whole wave of streams runs concurrently and every stream has
//     16 to 112 producer warps on other SMs tabulating for it;
//   * up to two CTAs per SM: a CTA per stream (K3c); within that one wave every stream decodes 1.3x
//     (rans/test.py's mixture of distributions) to 3x (narrow distributions) faster than a lane of the
//     lane-per-stream kernel; a second wave would cost more than it gains (measured crossover:
//     296 -> 444 streams);
//   * more: a lane per stream.
// FLIC_DEC_COOP_MAX_STREAMS overrides the CTA-per-stream count (0 disables both cooperative
// kernels); flic_set_decode_kernel() forces a kernel.
static std::atomic<int> g_decode_kernel{-1};
int set_decode_kernel(int which) { return g_decode_kernel.exchange(which); }

// 0: lane per stream; 1: CTA per stream; 2, 4, 8: cluster per stream
static int decode_kernel_for(int64_t n_streams) {
    const int forced = g_decode_kernel.load(std::memory_order_relaxed);
    if (forced == 0 || forced == 1) return forced;
    if (forced == 2 || forced == 4 || forced == 8) return coop_cluster_capacity(forced) > 0 ? forced : 1;
    static const int64_t env = [] {
        const char* e = getenv("FLIC_DEC_COOP_MAX_STREAMS");
        return e ? (int64_t)atoll(e) : (int64_t)-1;
    }();
    if (env == 0) return 0;
    for (int c = 8; c >= 2; c >>= 1)
        if (n_streams <= coop_cluster_capacity(c)) return c;
    return n_streams <= (env >= 0 ? env : 2 * (int64_t)sm_count()) ? 1 : 0;
}

cudaError_t launch_rans_decode(const uint32_t* packed, const int64_t* word_offsets,
                               const uint64_t* states, const float* mean, const float* scale,
                               const int64_t* offsets, int64_t n_streams, float* x_out,
                               uint64_t* end_states, int32_t* status, int check_end,
                               WordsLeft left, cudaStream_t stream) {
    if (n_streams <= 0) return cudaSuccess;
    if (const int k = decode_kernel_for(n_streams)) {
        note_coder_kernel(1, k == 1 ? "rans_decode_coop_kernel" : k == 2 ? "rans_decode_coop_kernel<cluster 2>"
                                     : k == 4 ? "rans_decode_coop_kernel<cluster 4>" : "rans_decode_coop_kernel<cluster 8>");
        return launch_rans_decode_coop(packed, word_offsets, states, mean, scale, offsets, n_streams, x_out,
                                       end_states, status, check_end, left, k, stream);
    }
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
                packed, word_offsets, states, mean, scale, offsets, n_streams, x_out, end_states, status, check_end, left, sh_m);
        else
            rans_decode_lane_kernel<kCoderWarps><<<(unsigned)blocks, kCoderWarps * 32, 0, stream>>>(
                packed, word_offsets, states, mean, scale, offsets, n_streams, x_out, end_states, status, check_end, left, sh_m);
    } else if (small) {
        rans_decode_kernel<1><<<(unsigned)warps, 32, 0, stream>>>(
            packed, word_offsets, states, mean, scale, offsets, n_streams, x_out, end_states, status, check_end, left);
    } else {
        rans_decode_kernel<kCoderWarps><<<(unsigned)blocks, kCoderWarps * 32, 0, stream>>>(
            packed, word_offsets, states, mean, scale, offsets, n_streams, x_out, end_states, status, check_end, left);
    }
    return cudaGetLastError();
}

}  // namespace flic
