// K1+K2 fused -- table evaluation and 64-bit-state / 32-bit-word rANS encode, many streams.
//
// Replaces rans.encode (rans/rans.pyx:37-67): pass 1 (per-symbol start/freq, :49-56) and pass 2
// (the serial renorm-then-update recurrence, :58-67), for every stream of a partition at once.
// Each stream starts at state 1<<32 like the reference's call site (trainer.py:310), or at
// init_states[s] (coder.py:25 chains states), and its words come out in emission order, so
// (final_state, words) per stream is what the reference's encode() returns for the same slice.
//
// Mapping.  A CTA owns 32 consecutive streams and walks them in tiles of 32 symbols per stream.
// The work splits by what is parallel and what is not:
//   producers (4 or 8 warps)  pass 1.  For each of the tile's 32 rows (one stream's 32 consecutive
//       symbols) a warp loads x, mean, scale as three coalesced 128-byte rows and evaluates
//       (start, freq) -- the expensive FP64 part, perfectly parallel over symbols -- into a
//       shared-memory tile.  Rows are taken 4 at a time, their loads issued up front.
//   consumer (1 warp)    pass 2.  Lane l replays row l of the finished tile through the serial
//       rANS recurrence, state in registers, and appends the emitted words to the stream's
//       scratch region.
// Tiles are double-buffered: while the consumer runs tile k the producers fill tile k+1, with one
// CTA barrier per tile.  Nine warps are resident per 32 streams, which keeps the SMs busy at
// stream counts where a lane-per-stream kernel would leave them idle, and a single long stream
// still gets 8 warps' worth of table evaluation.
// The (start, freq) tables never touch HBM: algorithmic traffic is the 12 B/symbol of inputs plus
// the coded bits.  Tile pitch is 33 entries so the consumer's lane-per-row reads spread over the banks;
// an entry carries the reciprocal the push needs, computed by the producers (it does not depend on
// the state), so the consumer's chain is convert -> multiply -> fix-up.
#include "flic_device.cuh"
#include "flic_kernels.cuh"

#include <atomic>

namespace flic {

// Names are string literals; the pointers are atomics so that threads driving different GPUs
// never read a torn value (the last writer wins, which is all a profiling aid promises).
static std::atomic<const char*> g_last_kernel[2] = {{""}, {""}};
const char* last_coder_kernel(int which) { return g_last_kernel[which & 1].load(std::memory_order_relaxed); }
void note_coder_kernel(int which, const char* name) { g_last_kernel[which & 1].store(name, std::memory_order_relaxed); }

// Rows a producer warp loads (and then evaluates) back to back.
constexpr int kRowBatch = 4;

// Stream count (in warps of 32 streams per SM) from which the lane-per-stream kernel is used.
// Measured with ImageNet64-shaped streams (tools/variants.py, -DFLIC_ENC_LANE_MIN_WARPS=...): the
// tile kernel wins up to 5.2 warps per SM (94.8 against 77.6 G symbols/s), the two tie at 10.4, the
// lane kernel wins from 15.6 (143 against 121) and by 36 % at 20.8 (159.5 against 117.7).
#ifndef FLIC_ENC_LANE_MIN_WARPS
#define FLIC_ENC_LANE_MIN_WARPS 12
#endif
constexpr int kLaneKernelMinWarpsPerSm = FLIC_ENC_LANE_MIN_WARPS;

// PRODUCERS per CTA.  The consumer issues ~1440 instructions per tile (32 steps x ~45) and a
// producer 6400 / PRODUCERS; every resident warp gets the same share of its scheduler, so a CTA's
// tile takes as long as its busiest warp.  4 producers balance the two roles (1600 vs 1440) and
// keep the issue slots full when many CTAs are resident; 8 producers halve the latency of a
// tile when only a few CTAs exist (few streams).
template <int PRODUCERS>
__global__ void __launch_bounds__((PRODUCERS + 1) * 32)
rans_encode_kernel(const float* __restrict__ x, const float* __restrict__ mean,
                   const float* __restrict__ scale, const int64_t* __restrict__ offsets,
                   int64_t n_streams, const uint64_t* __restrict__ init_states,
                   uint32_t* __restrict__ scratch, int64_t* __restrict__ counts,
                   uint64_t* __restrict__ states, int32_t* __restrict__ status) {
    __shared__ __align__(256) uint64_t s_tab[32];
    // (start, freq) and the biased reciprocal of freq the push divides with (flic_core.cuh: rans_push_rf)
    __shared__ uint4 s_tile[2][kLanes][kTile + 1];
    __shared__ int64_t s_beg[kLanes], s_len[kLanes];
    __shared__ int32_t s_flags[kLanes];
    __shared__ int64_t s_max_len;

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int64_t first = (int64_t)blockIdx.x * kLanes;

    if (warp == 0) {
        const int64_t stream = first + lane;
        const bool live = stream < n_streams;
        const int64_t beg = live ? offsets[stream] : 0;
        const int64_t len = live ? offsets[stream + 1] - beg : 0;
        s_beg[lane] = beg;
        s_len[lane] = len;
        s_flags[lane] = 0;
        const int64_t mx = warp_max_i64(len);
        if (lane == 0) s_max_len = mx;
    }
    const ExpTab tab = stage_exp_table(s_tab);  // CTA barrier inside
    const int64_t n_tiles = (s_max_len + kTile - 1) / kTile;

    if (warp > 0) {
        // ------------------------------------------------------------------ producers
        const int pw = warp - 1;
        for (int64_t k = 0; k < n_tiles; ++k) {
            uint4(*tile)[kTile + 1] = s_tile[k & 1];
            const int64_t i = k * kTile + lane;
#pragma unroll 1
            for (int r0 = pw * kRowBatch; r0 < kLanes; r0 += PRODUCERS * kRowBatch) {
                float xv[kRowBatch], mv[kRowBatch], sv[kRowBatch];
                bool on[kRowBatch];
#pragma unroll
                for (int q = 0; q < kRowBatch; ++q) {  // all loads of the batch first
                    const int r = r0 + q;
                    on[q] = i < s_len[r];
                    xv[q] = mv[q] = 0.0f;
                    sv[q] = 1.0f;
                    if (on[q]) {
                        const int64_t g = s_beg[r] + i;
                        xv[q] = __ldg(x + g);
                        mv[q] = __ldg(mean + g);
                        sv[q] = __ldg(scale + g);
                    }
                }
#pragma unroll
                for (int q = 0; q < kRowBatch; ++q) {
                    const int r = r0 + q;
                    if (on[q]) {
                        int32_t f = 0;
                        const SymbolTable e = make_table(xv[q], mv[q], sv[q], tab, f);
                        const double rf = push_reciprocal(e.freq);
                        tile[r][lane] = make_uint4(e.start, e.freq, (uint32_t)__double2loint(rf), (uint32_t)__double2hiint(rf));
                        if (f) atomicOr(&s_flags[r], f);  // rare
                    }
                }
            }
            cta_sync();  // tile k complete; the consumer has finished tile k-1, so buffer (k+1)&1 is free
        }
        cta_sync();      // pairs with the consumer's barrier after its last tile
    } else {
        // ------------------------------------------------------------------ consumer
        const int64_t stream = first + lane;
        const bool live = stream < n_streams;
        const int64_t beg = s_beg[lane], len = s_len[lane];
        uint64_t state = (live && init_states) ? init_states[stream] : kRansL;
        uint32_t* wp = scratch + beg;  // the stream's scratch region starts at its first symbol index
        cta_sync();                    // tile 0 complete
        for (int64_t k = 0; k < n_tiles; ++k) {
            uint4(*tile)[kTile + 1] = s_tile[k & 1];
            const int64_t rem = len - k * kTile;
            const int cnt = rem >= kTile ? kTile : (rem > 0 ? (int)rem : 0);
            // the tile's words are counted in 32 bits against the (opaque) pointer the tile starts at:
            // an emit is address + predicated store + predicated increment, where a running 64-bit
            // pointer advanced by a 0 / 1 select costs the lone consumer warp eight to ten
            // instructions per symbol
            uint32_t* wt = wp;
            asm volatile("" : "+l"(wt));
            uint32_t wn = 0;
            if (cnt == kTile) {
                // whole row: no per-symbol test, and the entries are fetched four symbols ahead of
                // their use, so the shared-memory latency is not on the state's dependency chain --
                // what is left per symbol is renorm select -> convert -> multiply -> floor -> fix-up
                uint4 e[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) e[u] = tile[lane][u];
#pragma unroll
                for (int j0 = 0; j0 < kTile; j0 += 4) {
                    uint4 nx[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) nx[u] = tile[lane][(j0 + 4 + u) & (kTile - 1)];   // last round: re-reads 0..3, unused
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        uint32_t word;
                        const bool emit = rans_push_rf(state, e[u].x, e[u].y, __hiloint2double((int)e[u].w, (int)e[u].z), word);
                        if (emit) {
                            asm volatile("st.global.u32 [%0], %1;" ::"l"(wt + wn), "r"(word) : "memory");
                            ++wn;
                        }
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) e[u] = nx[u];
                }
            } else {
                for (int j = 0; j < cnt; ++j) {
                    const uint4 e = tile[lane][j];
                    uint32_t word;
                    if (rans_push_rf(state, e.x, e.y, __hiloint2double((int)e.w, (int)e.z), word)) {
                        asm volatile("st.global.u32 [%0], %1;" ::"l"(wt + wn), "r"(word) : "memory");
                        ++wn;
                    }
                }
            }
            wp += wn;
            cta_sync();  // tile k consumed; tile k+1 complete
        }
        if (live) {
            counts[stream] = (int64_t)(wp - (scratch + beg));
            states[stream] = state;
            status[stream] = s_flags[lane];
        }
    }
}

// ---- many streams: every lane codes its own stream ------------------------------------------------
// With enough streams to fill the SMs on their own (a lane each), the producer/consumer split is
// not needed: a lane stages 32-byte blocks of ITS stream's x / mean / scale with 16-byte cp.async
// (one DRAM sector per array per 8 symbols), evaluates the eight table entries and pushes them
// through its rANS state in the same loop.  No CTA barriers, no role imbalance, no tile
// transposes: ~170 instructions per 32 symbols instead of ~215.  Blocks are cut on the 32-byte grid
// of the arrays' addresses (see rans_decode.cu); the launcher picks this kernel only when the
// three arrays share the same phase.
#ifndef FLIC_LANE_BLOCK
#define FLIC_LANE_BLOCK 8
#endif
constexpr int kBlk = FLIC_LANE_BLOCK;   // symbols per lane per block: 8 (one 32-byte sector) or 4
constexpr int kBlkShift = kBlk == 8 ? 3 : 2;
// A lane's row in shared memory is exactly its block (32 B or 16 B).  For 8-symbol rows the two
// 16-byte halves of rows 4..7, 12..15, ... are swapped (XOR swizzle on bit 2 of the column), which
// puts the 16-byte accesses of every quarter-warp on disjoint banks without padding: 24.6 KB per
// CTA instead of 37 KB, so shared memory no longer limits residency.
constexpr int kBlkPitch = kBlk;

__device__ __forceinline__ void cp_async_16(float* smem_dst, const float* gmem_src) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_4(float* smem_dst, const float* gmem_src) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(gmem_src) : "memory");
}

// 4 resident CTAs per SM (120 registers): the eight table evaluations of a block are independent, and
// with the registers to keep all of them in flight the compiler hides the FP64 latencies inside one
// warp, which pays more than the warps it costs -- 175.3 G symbols/s against 166.5 at 7 CTAs
// (72 registers, an 8-byte spill), 170 at 6 (80), 167.6 at 5 (96), 166.8 at 8 (64); 3 and 2 compile
// to the same 120-register code.  Fewer lanes streaming at once also means less DRAM over-fetch: the
// more lanes, the more of the 64-byte DRAM bursts' second halves leave L2 before their lane asks for
// them (ncu: 23.5 / 22.0 / 20.4 GB read at 9 / 7 / 6 CTAs for 19.3 GB of inputs).
// Bookkeeping is 32-bit and per lane as in the decoder (rans_decode.cu): blocks q = 0 .. nb - 1,
// only the first and the last can be partial, one running element index for the three arrays.
#ifndef FLIC_ENC_MIN_BLOCKS
#define FLIC_ENC_MIN_BLOCKS 4
#endif
template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32, FLIC_ENC_MIN_BLOCKS)
rans_encode_lane_kernel(const float* __restrict__ x, const float* __restrict__ mean,
                        const float* __restrict__ scale, const int64_t* __restrict__ offsets,
                        int64_t n_streams, const uint64_t* __restrict__ init_states,
                        uint32_t* __restrict__ scratch, int64_t* __restrict__ counts,
                        uint64_t* __restrict__ states, int32_t* __restrict__ status, int shift) {
    __shared__ __align__(256) uint64_t s_tab[32];
    __shared__ __align__(16) float s_in[WARPS][2][3][kLanes][kBlkPitch];   // [buffer][x, mean, scale]
    const ExpTab tab = stage_exp_table(s_tab);

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int64_t first = ((int64_t)blockIdx.x * WARPS + warp) * kLanes;
    if (first >= n_streams) return;
    const int64_t stream = first + lane;
    const bool live = stream < n_streams;
    const int64_t beg = live ? offsets[stream] : 0;
    int64_t len = live ? offsets[stream + 1] - beg : 0;
    const bool too_long = len > 0x7fffffffll;     // 32-bit counters: one stream of 2^31 symbols is not supported
    if (too_long) len = 0;
    const int64_t end = beg + len;
    const int64_t t_lo = (beg + shift) >> kBlkShift, t_hi = (end - 1 + shift) >> kBlkShift;
    const int nb = len > 0 ? (int)(t_hi - t_lo + 1) : 0;
    int n_iter = nb;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) n_iter = max(n_iter, __shfl_xor_sync(0xffffffffu, n_iter, d));
    // symbols [j_head, kBlk) of block 0 and [0, j_tail) of block nb - 1 belong to the stream
    const int j_ends = (int)((beg + shift) & (kBlk - 1)) | (((int)((end - 1 + shift) & (kBlk - 1)) + 1) << 4);

    uint64_t state = (live && init_states) ? init_states[stream] : kRansL;
    // the stream's scratch region starts at its first symbol index; the words are counted in 32 bits
    // (a stream has fewer than 2^31 symbols and emits at most one word per symbol), so an emit is an
    // address, a predicated store and a predicated increment -- a running 64-bit pointer costs two
    // carries per symbol and the compiler keeps a second copy of it for the final count
    uint32_t* wbase = scratch + beg;
    asm volatile("" : "+l"(wbase));     // one register pair, not re-derived from the parameters at every emit
    uint32_t wn = 0;
    int32_t flags = too_long ? ST_TOO_LONG : 0;
    ParamGuard guard = guard_init();
    constexpr int kArr = kLanes * kBlkPitch, kBuf = 3 * kArr;
    float* const row = s_in[warp][0][0][lane];
    const int swz = kBlk == 8 ? ((lane >> 2) & 1) * 4 : 0;   // column j of this lane lives at j ^ swz
    int64_t i_stage = (t_lo << kBlkShift) - shift;           // first slot of the block staged next

    auto stage = [&](int q) {
        if (q < nb) {
            float* d = row + (q & 1) * kBuf;
            const float *px = x + i_stage, *pm = mean + i_stage, *ps = scale + i_stage;
            const int jl = q == 0 ? (j_ends & 15) : 0, jh = q == nb - 1 ? (j_ends >> 4) : kBlk;
            if (jl == 0 && jh == kBlk) {
                cp_async_16(d + swz, px);
                cp_async_16(d + kArr + swz, pm);
                cp_async_16(d + 2 * kArr + swz, ps);
                if (kBlk == 8) {
                    cp_async_16(d + (4 ^ swz), px + 4);
                    cp_async_16(d + kArr + (4 ^ swz), pm + 4);
                    cp_async_16(d + 2 * kArr + (4 ^ swz), ps + 4);
                }
            } else {
#pragma unroll
                for (int j = 0; j < kBlk; ++j)
                    if (j >= jl && j < jh) {
                        cp_async_4(d + (j ^ swz), px + j);
                        cp_async_4(d + kArr + (j ^ swz), pm + j);
                        cp_async_4(d + 2 * kArr + (j ^ swz), ps + j);
                    }
            }
            i_stage += kBlk;
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    auto push = [&](const SymbolTable& e) {
        uint32_t word;
        if (rans_push(state, e.start, e.freq, word)) {   // (wbase is opaque: say that it is global memory)
            asm volatile("st.global.u32 [%0], %1;" ::"l"(wbase + wn), "r"(word) : "memory");
            ++wn;
        }
    };

    if (n_iter > 0) stage(0);
    for (int q = 0; q < n_iter; ++q) {
        if (q + 1 < n_iter) {
            stage(q + 1);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        if (q < nb) {
            const float* b = row + (q & 1) * kBuf;
            const int jl = q == 0 ? (j_ends & 15) : 0, jh = q == nb - 1 ? (j_ends >> 4) : kBlk;
            if (kBlk == 8 && jl == 0 && jh == kBlk) {
                // whole block: the row comes out of shared memory as 16-byte vectors (conflict-free
                // by the swizzle) and its eight symbols are coded from registers, fully unrolled, so
                // that the table evaluations of later symbols fill the latencies of earlier pushes
                float xs[8], ms[8], ss[8];
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int c = (4 * h) ^ swz;
                    const float4 xv = *reinterpret_cast<const float4*>(b + c);
                    const float4 mv = *reinterpret_cast<const float4*>(b + kArr + c);
                    const float4 sv = *reinterpret_cast<const float4*>(b + 2 * kArr + c);
                    xs[4 * h] = xv.x; xs[4 * h + 1] = xv.y; xs[4 * h + 2] = xv.z; xs[4 * h + 3] = xv.w;
                    ms[4 * h] = mv.x; ms[4 * h + 1] = mv.y; ms[4 * h + 2] = mv.z; ms[4 * h + 3] = mv.w;
                    ss[4 * h] = sv.x; ss[4 * h + 1] = sv.y; ss[4 * h + 2] = sv.z; ss[4 * h + 3] = sv.w;
                }
#pragma unroll
                for (int k = 0; k < 8; ++k) push(make_table_lean(xs[k], ms[k], ss[k], tab, guard, flags));
            } else {
                for (int j = jl; j < jh; ++j) {
                    const int c = j ^ swz;
                    push(make_table_lean(b[c], b[kArr + c], b[2 * kArr + c], tab, guard, flags));
                }
            }
        }
    }
    if (live) {
        counts[stream] = (int64_t)wn;
        states[stream] = state;
        status[stream] = flags | guard_flags(guard);
    }
}

cudaError_t launch_rans_encode(const float* x, const float* mean, const float* scale,
                               const int64_t* offsets, int64_t n_streams,
                               const uint64_t* init_states, uint32_t* scratch, int64_t* counts,
                               uint64_t* states, int32_t* status, cudaStream_t stream) {
    if (n_streams <= 0) return cudaSuccess;
    const int64_t blocks = (n_streams + kLanes - 1) / kLanes;
    const int sh_x = (int)(((uintptr_t)x >> 2) & (kBlk - 1)), sh_m = (int)(((uintptr_t)mean >> 2) & (kBlk - 1));
    const int sh_s = (int)(((uintptr_t)scale >> 2) & (kBlk - 1));
    const bool same_phase = sh_x == sh_m && sh_x == sh_s && (((uintptr_t)x | (uintptr_t)mean | (uintptr_t)scale) & 3) == 0;
    // a lane per stream pays from ~12 warps of streams per SM upwards
    if (same_phase && blocks >= (int64_t)sm_count() * kLaneKernelMinWarpsPerSm) {
        const int64_t ctas = (blocks + kCoderWarps - 1) / kCoderWarps;
        rans_encode_lane_kernel<kCoderWarps><<<(unsigned)ctas, kCoderWarps * 32, 0, stream>>>(
            x, mean, scale, offsets, n_streams, init_states, scratch, counts, states, status, sh_x);
        note_coder_kernel(0, "rans_encode_lane_kernel");
    } else if (blocks >= (int64_t)sm_count() * 6) {
        rans_encode_kernel<4><<<(unsigned)blocks, 5 * 32, 0, stream>>>(
            x, mean, scale, offsets, n_streams, init_states, scratch, counts, states, status);
        note_coder_kernel(0, "rans_encode_kernel");
    } else {
        rans_encode_kernel<8><<<(unsigned)blocks, 9 * 32, 0, stream>>>(
            x, mean, scale, offsets, n_streams, init_states, scratch, counts, states, status);
        note_coder_kernel(0, "rans_encode_kernel");
    }
    return cudaGetLastError();
}

}  // namespace flic
