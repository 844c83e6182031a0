// K3c -- rANS decode for FEW streams: one CTA per stream, exact CDF windows evaluated ahead of
// the serial chain.
//
// Replaces rans.decode (rans/rans.pyx:69-110) where the reference actually calls it: one stream
// per latent level per batch (trainer.py:308-318), one 50 M-symbol stream (rans/test.py:6-22).
// A stream is a serial chain -- symbol i's slot `mod` is known only after symbol i+1 has been
// popped -- and in the lane-per-stream kernel (rans_decode.cu) every link of that chain carries
// the symbol guess and two exact CDF evaluations: ~1000 cycles per symbol for a lone warp, so a
// handful of streams decode at 2-3 M symbols/s each while the rest of the GPU idles.
//
// What does NOT depend on the state is the distribution: mean and scale of every symbol are
// inputs.  So the exact CDF of symbol i can be tabulated before its turn comes, for the bins
// where the symbol is likely to be: a window of 5.1 scale units either side of the mode (98.8 % of
// a logistic's mass).  The CTA is split by role:
//   producers            for the next group of 32 symbols: per-symbol model, window size and
//       placement; then one exact CDF evaluation per lane per 32-entry chunk of every window
//       (cdf_at(), the same function every other kernel uses).  A chunk is built in registers --
//       entry e = (CDF(b - 1), CDF(b)) of its bin b, the left value taken from the lane below --
//       and stored with one 8-byte store per lane; consecutive chunks overlap by one bin, so no
//       chunk needs a value another warp computes.  Chunks of all 32 symbols form one task list
//       that the producer warps share round-robin.
//   consumer (1 warp)    the chain.  All lanes hold the same state; per symbol: pull a word if
//       needed, mod = state & 0xffffff, each lane tests its tabulated entry and computes the pop
//       AS IF the entry were the symbol's, and an OR-reduction of the one lane that is right
//       broadcasts the new state -- the chain is select, mask, subtract, multiply-add, select,
//       reduce: under a hundred cycles instead of a thousand.  The stream's words wait in
//       registers (64 at a time, one per lane, fetched by shuffle), so renormalisation puts no
//       address arithmetic or load into the instruction stream of the chain.
// A symbol whose window was not tabulated (group over capacity, bad parameters) or that falls
// outside it is decoded on the chain by decode_symbol_model() exactly as in the lane kernel, by
// all lanes at once.  Results are bit-identical by construction: the window holds exact CDF values
// and the search returns the smallest in-window s with CDF(s) > mod (CDF is non-decreasing,
// SURVEY A.2).
//
// K3d -- the same with a thread-block CLUSTER per stream (a handful of streams: the reference's
// own partition is one stream per latent level, trainer.py:308-318, and one stream in
// rans/test.py).  With one CTA the producers keep up only with narrow distributions, so symbols
// wider than 12 bins per scale unit (a quarter of rans/test.py's mixture) are decoded on the chain
// at ~1000 cycles each.  A cluster of 2, 4 or 8 CTAs puts 16 producer warps on each of the other
// SMs of the cluster; they write their chunks straight into the home CTA's shared memory
// (st.shared::cluster through a mapa-translated address: distributed shared memory), and the
// window grows to 32 chunks (992 bins: 5.1 scale units either side up to a scale of 97 bins),
// which the chain searches in two ballots (chunk index, then entry).  Groups cycle through three
// buffers and are handed over with the cluster barrier split into its arrive and wait halves: the
// consumer arrives for group g + 1 BEFORE it decodes group g and waits after, so the barrier's
// latency (about a microsecond across SMs) is spent while the chain is busy.  The home CTA keeps
// the chain's scheduler free: its own producers are the warps of the other three sub-partitions.
#include "flic_device.cuh"
#include "flic_kernels.cuh"

#include <type_traits>
#ifndef FLIC_COOP_STATS
#define FLIC_COOP_STATS 0      // 1: the consumer counts the paths its symbols took and the cycles it waited (printf at the end)
#endif
#if FLIC_COOP_STATS
#include <cstdio>
#define COOP_STAT(...) __VA_ARGS__
#else
#define COOP_STAT(...)
#endif

namespace flic {

#ifndef FLIC_COOP_UNROLL
#define FLIC_COOP_UNROLL 4
#endif
#ifndef FLIC_CLUSTER_HOME_PRODUCERS
#define FLIC_CLUSTER_HOME_PRODUCERS 1
#endif
#ifndef FLIC_COOP1_WARPS
#define FLIC_COOP1_WARPS 8      // CTA-per-stream kernel: warps per CTA (producers + the consumer)
#endif
#ifndef FLIC_COOP1_SLOTS
#define FLIC_COOP1_SLOTS 160    // and 32-entry slots per group buffer
#endif
constexpr int kCoopUnroll = FLIC_COOP_UNROLL;   // evaluations a producer warp interleaves
constexpr unsigned kFull = 0xffffffffu;
constexpr int kCoopSlotBase = 2560;     // bytes of shared memory in front of the slots
constexpr int kNever = 0x7fffffff;      // p = v = kNever: an empty interval no mod falls into
constexpr int kChunkBins = 31;          // bins a chunk adds to its window (entry 0 only carries the left value)

// Geometry by cluster size C (CTAs per stream).
template <int C>
struct CoopCfg {
    // C == 1: warps 0..6 produce, warp 7 is the consumer (more producers take issue slots from the
    // chain: 15 cost it 17 %); two CTAs fit an SM.  C > 1: 16 warps per CTA, the last one of the home
    // CTA is the consumer, alone on its sub-partition.
    static constexpr int kWarps = C == 1 ? FLIC_COOP1_WARPS : 16;
    static constexpr int kConsumerWarp = kWarps - 1;
    static constexpr int kHomeProducers = C == 1 ? FLIC_COOP1_WARPS - 1 : (FLIC_CLUSTER_HOME_PRODUCERS ? 12 : 0);
    static constexpr int kProducers = kHomeProducers + (C - 1) * kWarps;   // per stream
    static constexpr int kBuffers = C == 1 ? 2 : 3;       // group buffers in flight
    static constexpr int kSlots = C == 1 ? FLIC_COOP1_SLOTS : 296;     // 32-entry slots per group buffer
    static constexpr int kMaxChunks = C == 1 ? 4 : 32;    // 32-bin chunks per wide window
    static constexpr int kRows = (kSlots - 32) * 2;       // 128-byte rows (one chunk of a wide window each) behind the 32 head slots
    static constexpr size_t kSmemBytes = kCoopSlotBase + 8u * kBuffers * kSlots * 32;
};

// ---- cluster plumbing (C > 1) ------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
// shared::cluster address of `local_shared_addr` in the CTA of rank `rank`
__device__ __forceinline__ uint32_t cluster_map(uint32_t local_shared_addr, uint32_t rank) {
    uint32_t a;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(a) : "r"(local_shared_addr), "r"(rank));
    return a;
}
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
// The consumer's arrival only says "I have finished READING a buffer": nothing it wrote is for the
// producers, so it needs no release fence (which would wait for its decoded-symbol stores to drain).
__device__ __forceinline__ void cluster_arrive_relaxed() { asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory"); }
// one tabulated entry (8 bytes) into the home CTA's slot memory
template <int C>
__device__ __forceinline__ void st_entry(uint32_t a, int p, int v) {
    if constexpr (C == 1) asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(a), "r"(p), "r"(v) : "memory");
    else asm volatile("st.shared::cluster.v2.b32 [%0], {%1, %2};" ::"r"(a), "r"(p), "r"(v) : "memory");
}

template <int C>
__device__ __forceinline__ void st_value(uint32_t a, int v) {
    if constexpr (C == 1) asm volatile("st.shared.b32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
    else asm volatile("st.shared::cluster.b32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}

// One tabulated bin: v = CDF(bin), p = CDF(bin - 1).  The symbol is the one entry with
// p <= mod < v, a test each lane makes on its own; (p, v - p) are the (start, freq) of the pop.
struct __align__(8) CoopEntry {
    int p, v;
};

__device__ __forceinline__ double shfl_f64(double v, int src) {
    const int lo = __shfl_sync(kFull, __double2loint(v), src);
    const int hi = __shfl_sync(kFull, __double2hiint(v), src);
    return __hiloint2double(hi, lo);
}

// Per-lane description of one symbol of the current group (lane j <-> symbol j of the group).
// Slot j of the group buffer is symbol j's HEAD: its one chunk (n == 1), or an index of its chunks
// (n > 1), which then live in slots >= 32.  The chain therefore knows where to look first without
// reading any metadata.
struct CoopSymbol {
    SymbolModel m;
    float mean, scale;
    int n;       // chunks tabulated; 0: none (decoded on the chain); -1: bad parameters (on the chain, which flags them)
    int t_off;   // n > 1: first row
    int w0;      // window origin (see coop_describe)
};
struct CoopGroup {
    unsigned ones;   // symbols with n == 1
    unsigned multis; // symbols with n > 1
    int n_ones;      // their number: tasks [0, n_ones) evaluate them, tasks from n_ones on the rows of the wide windows
    int total;       // tasks of the group
};

// Every warp of the CTA (of the cluster) computes the same description from the same inputs, so no
// metadata has to cross warps.
//
// Window: one chunk (15 bins either side of the mode) up to a logistic scale of 3 bins (1.3 % of
// a logistic's mass lies outside, and a symbol there costs a few hundred cycles on the chain);
// above that 5.1 scale units either side, in 2 to kMaxChunks chunks, which the chain searches in
// two steps.  One CTA: up to 12 bins per scale unit (4 chunks); wider distributions are not
// tabulated there: at ~45 cycles per chunk (7 producer warps) the tabulation would take as long as
// the lane kernel's step, which is what those symbols get.  A cluster tabulates every width (32
// chunks cover 5.1 scale units up to a scale of 97 bins, and 3.3 at the 148 bins that end
// rans/test.py's range).
// MODEL: also the FP64 model the evaluations need (producers); the consumer only needs the layout.
template <int C, bool MODEL>
__device__ __forceinline__ CoopGroup coop_describe(float mean, float scale, bool valid, int lane, CoopSymbol& d) {
    using Cfg = CoopCfg<C>;
    d.mean = mean;
    d.scale = scale;
    if constexpr (MODEL) d.m = make_model(mean, scale);
    else d.m.lower = lower_of(mean);
    const float cb = scale * 256.0f;                // logistic scale in bins
    const bool ok = valid && params_ok(mean, scale);
    int n = 0;
    if (ok && cb <= 3.0f) n = 1;
    else if (ok && (C > 1 || cb <= 12.0f)) {
        const int half = (int)(fminf(cb, 128.0f) * 5.1f) + 2;       // bins either side of the mode
        n = (2 * half + 5 + 31) >> 5;                               // two more bins either side: the guess may be one off
        n = n < Cfg::kMaxChunks ? n : Cfg::kMaxChunks;
        n = n < 2 ? 2 : n;
    }
    const int many = n > 1 ? n : 0;
    int incl = many;
#pragma unroll
    for (int s = 1; s < 32; s <<= 1) {
        const int o = __shfl_up_sync(kFull, incl, s);
        if (lane >= s) incl += o;
    }
    if (many && incl > Cfg::kRows) n = 0;           // over capacity: decoded on the chain (narrowing every window
                                                    // of the group in proportion instead was measured: worse on
                                                    // rans/test.py's symbols, which lie uniformly within 5 scale units)
    d.t_off = incl - many;                          // first row
    d.n = ok ? n : -1;
    // centred on the mode (lower + 1024 = round(256 mean)): with at most 512 bins either side the
    // window lies inside the coder's 2048-bin support, so it needs no edge cases.
    //   n == 1: entry e of the head slot stands for bin w0 - 1 + e (e = 1 .. 31: mode - 15 .. mode + 15)
    //   n > 1:  value r of the rows stands for bin w0 + r (r = 0 .. 32 n - 1)
    d.w0 = d.m.lower + 1024 - (n > 1 ? 16 * n : 15);
    CoopGroup grp;
    grp.ones = __ballot_sync(kFull, n == 1);
    grp.multis = __ballot_sync(kFull, n > 1);
    grp.n_ones = __popc(grp.ones);
    // tasks of symbols that went over capacity are not evaluated: the count stops at the last slot in use
    const int used = __reduce_max_sync(kFull, n > 1 ? incl : 0);
    grp.total = grp.n_ones + used;
    return grp;
}

template <int C>
__global__ void __launch_bounds__(CoopCfg<C>::kWarps * 32)
rans_decode_coop_kernel(const uint32_t* __restrict__ packed, const int64_t* __restrict__ word_offsets,
                        const uint64_t* __restrict__ states, const float* __restrict__ mean,
                        const float* __restrict__ scale, const int64_t* __restrict__ offsets,
                        int64_t n_streams, float* __restrict__ x_out, uint64_t* __restrict__ end_states,
                        int32_t* __restrict__ status, int check_end, WordsLeft left) {
    using Cfg = CoopCfg<C>;
    constexpr int kSlots = Cfg::kSlots;
    extern __shared__ __align__(256) unsigned char s_raw[];
    uint64_t* const s_tab = reinterpret_cast<uint64_t*>(s_raw);                              // 256 B
    // s_raw + 512: 32 x 4 B code (chunk offset, or the symbol itself for a scalar step); + 640: 32 x 4 B
    // lane index of the matching entry (both: consumer only, addressed in the shared space)
    // slots: [kBuffers][kSlots][32] CoopEntry from s_raw + kCoopSlotBase, in the HOME CTA (cluster rank 0)
    const ExpTab tab = stage_exp_table(s_tab);

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const uint32_t rank = C == 1 ? 0u : cluster_ctarank();
    const int64_t stream = C == 1 ? (int64_t)blockIdx.x : (int64_t)(blockIdx.x / C);
    const int64_t beg = offsets[stream];
    int64_t len = offsets[stream + 1] - beg;
    const int64_t wbeg = word_offsets[stream];
    int64_t wcount = word_offsets[stream + 1] - wbeg;
    if (left.in) wcount = left.in[stream] < 0 ? 0 : (left.in[stream] < wcount ? left.in[stream] : wcount);
    const bool too_long = wcount > 0x7fffffffll || len > 0x7fffffffll;
    if (too_long) len = 0;
    const int n_groups = (int)((len + 31) >> 5);
    const float* const mean_s = mean + beg;
    const float* const scale_s = scale + beg;
    // every CTA of the cluster is running before anybody writes into the home CTA's shared memory
    if constexpr (C > 1) { cluster_arrive(); cluster_wait(); }

    // group g holds the symbols [len - 32 (g + 1), len - 32 g) that exist; lane j <-> the j-th of them
    auto load_group = [&](int g, float& mu, float& sc) -> bool {
        const int64_t i = len - 32 * (int64_t)(g + 1) + lane;
        const bool valid = g < n_groups && i >= 0;
        mu = valid ? __ldg(mean_s + i) : 0.0f;
        sc = valid ? __ldg(scale_s + i) : 1.0f;
        return valid;
    };

    // roles.  One CTA: warps 0..6 produce.  Cluster: in the home CTA the warps of sub-partitions
    // 0..2 (warp & 3 != 3; the consumer is alone on sub-partition 3), in the others all 16.
    const bool is_consumer = rank == 0 && warp == Cfg::kConsumerWarp;
    int producer = -1;                                  // index among the stream's producer warps
    if constexpr (C == 1) {
        producer = warp < Cfg::kHomeProducers ? warp : -1;
    } else {
        if (rank != 0) producer = Cfg::kHomeProducers + ((int)rank - 1) * Cfg::kWarps + warp;
        else if (Cfg::kHomeProducers && (warp & 3) != 3) producer = (warp >> 2) * 3 + (warp & 3);
    }

    // Hand-over protocol.
    //   C == 1: two buffers, one CTA barrier per group (producers: after tabulating g; consumer:
    //       before decoding g), one more at the end.
    //   C > 1: three buffers, cluster barrier phase g = "group g is tabulated and the consumer has
    //       begun group g - 1" (i.e. finished g - 2, whose buffer group g + 1 reuses).  Producers
    //       and idle warps arrive and wait once per group; the consumer arrives for phase g + 1
    //       before decoding group g and waits for it afterwards.  Every thread of the cluster goes
    //       through n_groups phases.
    if (producer >= 0) {
        // ------------------------------------------------------------------ producers
        uint32_t slot_base = (uint32_t)__cvta_generic_to_shared(s_raw) + (uint32_t)kCoopSlotBase;
        if constexpr (C > 1) slot_base = cluster_map(slot_base, 0u);
        float mu, sc;
        bool valid = load_group(0, mu, sc);
        for (int g = 0; g < n_groups; ++g) {
            CoopSymbol d;
            const CoopGroup grp = coop_describe<C, true>(mu, sc, valid, lane, d);
            valid = load_group(g + 1, mu, sc);          // in flight during the evaluations
            const uint32_t buf = slot_base + (uint32_t)(g % Cfg::kBuffers) * (uint32_t)(kSlots * 256);
            // U tasks per pass: their evaluations are independent and interleave (a lone warp is
            // latency-bound on the ~35-deep dependency chain of one evaluation); a warp whose share
            // of the group is a single task evaluates just that one
            auto run_tasks = [&](auto uc, int t0) {
                constexpr int U = decltype(uc)::value;
                int bins[U], cs[U], ns[U], js[U], slots[U];
                bool on[U];
                SymbolModel mj[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int t = t0 + u * Cfg::kProducers;
                    int j, c, sl;
                    if (t < grp.n_ones) {               // the t-th single-chunk symbol, in its head slot
                        j = (int)__fns(grp.ones, 0, t + 1);
                        c = 0;
                        sl = j;
                    } else {                            // a row of a wide window
                        sl = t - grp.n_ones;
                        const unsigned owners = __ballot_sync(kFull, d.n > 1 && d.t_off <= sl);
                        j = owners ? 31 - __clz(owners) : 0;
                        c = sl - __shfl_sync(kFull, d.t_off, j);
                    }
                    js[u] = j; cs[u] = c; slots[u] = sl;
                    ns[u] = __shfl_sync(kFull, d.n, j);
                    on[u] = t < grp.total && c < ns[u];
                    mj[u].mean_d = shfl_f64(d.m.mean_d, j);
                    mj[u].scale_d = shfl_f64(d.m.scale_d, j);
                    mj[u].rscale = shfl_f64(d.m.rscale, j);
                    mj[u].lower = __shfl_sync(kFull, d.m.lower, j);
                    const int w0 = __shfl_sync(kFull, d.w0, j);
                    bins[u] = ns[u] > 1 ? w0 + 32 * c + lane : w0 - 1 + lane;
                }
                int v[U];
#pragma unroll
                for (int u = 0; u < U; ++u) v[u] = cdf_at(bins[u], mj[u], tab);
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int below = __shfl_up_sync(kFull, v[u], 1);
                    if (!on[u]) continue;
                    if (ns[u] > 1) {
                        // a row of a wide window: the values alone, 4 bytes per bin
                        st_value<C>(buf + 8192u + 128u * (uint32_t)slots[u] + 4u * (uint32_t)lane, v[u]);
                    } else {
                        // head slot: entry e = (value of the lane below, own value); entry 0 has no left
                        // neighbour and is the empty interval p == v
                        st_entry<C>(buf + 256u * (uint32_t)slots[u] + 8u * (uint32_t)lane, lane ? below : v[u], v[u]);
                    }
                }
            };
            for (int t0 = producer; t0 < grp.total; t0 += kCoopUnroll * Cfg::kProducers) {
                if (t0 + Cfg::kProducers >= grp.total) run_tasks(std::integral_constant<int, 1>{}, t0);
                else run_tasks(std::integral_constant<int, kCoopUnroll>{}, t0);
            }
            if constexpr (C == 1) cta_sync();           // group g tabulated; the consumer has finished group g - 1
            else { cluster_arrive(); cluster_wait(); }  // phase g
        }
        if constexpr (C == 1) cta_sync();               // pairs with the consumer's barrier after the last group
    } else if (!is_consumer) {
        // ------------------------------------------------------------------ idle warps of the home CTA (C > 1)
        for (int g = 0; g < n_groups; ++g) { cluster_arrive(); cluster_wait(); }
    } else {
        // ------------------------------------------------------------------ consumer
        const uint32_t* const wbase = packed + wbeg;
        const uint64_t st0 = states[stream];
        uint32_t hi = (uint32_t)(st0 >> 32), lo = (uint32_t)st0;
        // The stream's next 64 words wait in registers, one (wbuf) and one (wnxt) per lane: lane L of
        // wbuf holds word wtop - 1 - L, of wnxt word wtop - 33 - L (words are consumed from the last
        // emitted to the first; words before the stream's first read as 0).  k words of wbuf are
        // consumed -- wtop - k are unread, negative once the stream has under-run --; the next one to
        // pull is w_next = wbuf of lane k, fetched by shuffle right after each pull, so it is ready
        // long before the following symbol can want it.  Before k can pass 31 the window slides (two
        // shuffles and a select) and wnxt is reloaded -- a load with some 200 symbols to arrive.
        int wtop = too_long ? 0 : (int)wcount, k = 0;
        auto ldw = [&](int idx) -> uint32_t { return idx >= 0 ? __ldg(wbase + idx) : 0u; };
        uint32_t wbuf = ldw(wtop - 1 - lane), wnxt = ldw(wtop - 33 - lane);
        uint32_t w_next = __shfl_sync(kFull, wbuf, 0);
        auto slide_words = [&]() {
            const int s = lane + k;
            const uint32_t a = __shfl_sync(kFull, wbuf, s), b = __shfl_sync(kFull, wnxt, s);   // source lane: s mod 32
            wbuf = s < 32 ? a : b;
            wtop -= k;
            k = 0;
            wnxt = ldw(wtop - 33 - lane);
        };
        int32_t flags = too_long ? ST_TOO_LONG : 0;
        ParamGuard guard = guard_init();
        float* const x_s = x_out + beg;
        // Shared memory is addressed in the shared space throughout: a generic pointer costs the
        // shared-window base at every use, which in a cluster kernel is a special-register read.
        uint32_t sm = (uint32_t)__cvta_generic_to_shared(s_raw);
        asm volatile("" : "+r"(sm));
        const uint32_t sm_code = sm + 512u, sm_lane = sm + 640u;
        auto sts32 = [&](uint32_t addr, int v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); };
        auto lds32 = [&](uint32_t addr) -> int {
            int v;
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
            return v;
        };
        auto lds_entry = [&](uint32_t addr) -> CoopEntry {
            CoopEntry e;
            asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(e.p), "=r"(e.v) : "r"(addr) : "memory");
            return e;
        };
        // the lane whose entry is the symbol's notes its index (a speculative note that is taken back
        // is simply overwritten)
        auto note_lane = [&](int j, bool mine) {
            if (mine) sts32(sm_lane + 4u * (uint32_t)j, lane);
        };
        // One tabulated entry against the state: this lane's candidate for the popped state and
        // whether the entry is the symbol's.  rans.pyx:87-90,108 with the pull folded in:
        // (h, l) is the state after the word pull, x = state >> 24, new = x freq + (mod - start).
        struct Cand { uint32_t hi, lo; bool mine; };
        auto candidate = [&](const CoopEntry e) -> Cand {
            const bool pulled = hi == 0u;
            const uint32_t h = pulled ? lo : hi, l = pulled ? w_next : lo;
            const uint32_t dm = (l & kProbMask) - (uint32_t)e.p;                           // mod - start
            const uint32_t fr = (uint32_t)(e.v - e.p);
            const uint32_t xl = (h << 8) | (l >> 24), xh = h >> 24;
            const uint64_t pr = (uint64_t)xl * fr + dm;
            Cand c;
            c.hi = (uint32_t)(pr >> 32) + xh * fr;
            c.lo = (uint32_t)pr;
            c.mine = dm < fr;                               // p <= mod < v, in unsigned arithmetic
            return c;
        };
        // After a symbol is decided: the word accounting of its pull (off the chain).
        auto account_pull = [&](bool pulled) {
            k += pulled ? 1 : 0;
            w_next = __shfl_sync(kFull, wbuf, k);
        };

        COOP_STAT(long long n_quad = 0, n_quad_fail = 0, n_one = 0, n_multi = 0, n_scalar = 0, n_slide = 0, t_wait = 0, t_setup = 0, t_quad = 0, t_one = 0, t_multi = 0, t_scalar = 0, t_store = 0;
                  const long long t_begin = clock64();)
        float mu, sc;
        bool valid = load_group(0, mu, sc);
        if constexpr (C > 1) {
            if (n_groups > 0) { cluster_arrive(); cluster_wait(); }      // phase 0
        }
        for (int g = 0; g < n_groups; ++g) {
            COOP_STAT(const long long t_s0 = clock64();)
            CoopSymbol d;
            const CoopGroup grp = coop_describe<C, false>(mu, sc, valid, lane, d);
            const int64_t base = len - 32 * (int64_t)(g + 1);
            const int j_lo = base < 0 ? (int)(-base) : 0;   // first group-slot that is a symbol
            valid = load_group(g + 1, mu, sc);
            sts32(sm_code + 4u * (uint32_t)lane, 0);
            __syncwarp();
            COOP_STAT(const long long t_s1 = clock64(); t_setup += t_s1 - t_s0;)
            if constexpr (C == 1) cta_sync();                            // group g tabulated
            else if (g + 1 < n_groups) cluster_arrive_relaxed();         // phase g + 1: group g - 1 is done
            COOP_STAT(t_wait += clock64() - t_s1;)
            // A warp issues in order: whatever sits in front of an instruction in the stream delays
            // it, needed or not.  So the common case -- single-chunk symbols found in their windows
            // -- runs up to four at a time in one straight line with nothing else in it:
            // "single-chunk" is a bit of `ones`, the head entries sit at fixed addresses (slot j), the
            // decoded value is worked out after the group from the lane index each step leaves
            // behind, and anything else is decoded by a step of its own.
            const uint32_t sm_slot = sm + (uint32_t)kCoopSlotBase + (uint32_t)(g % Cfg::kBuffers) * (uint32_t)(kSlots * 256) + 8u * (uint32_t)lane;
            // R single-chunk symbols j, j - 1, ..: decoded optimistically (no branch, so nothing between
            // one symbol's broadcast and the next one's search), and taken back as a whole if any of
            // them fell outside its window.
            auto run_singles = [&](auto rc, int j) -> bool {
                constexpr int R = decltype(rc)::value;
                const uint32_t a0 = sm_slot + 256u * (uint32_t)j;
                CoopEntry es[R];
#pragma unroll
                for (int i = 0; i < R; ++i) es[i] = lds_entry(a0 - 256u * (uint32_t)i);
                const uint32_t s_hi = hi, s_lo = lo, s_next = w_next;
                const int s_k = k;
                uint32_t all_found = 1u;
#pragma unroll
                for (int i = 0; i < R; ++i) {
                    const bool pulled = hi == 0u;
                    const Cand c = candidate(es[i]);
                    const uint32_t rhi = __reduce_or_sync(kFull, c.mine ? c.hi : 0u);
                    const uint32_t rlo = __reduce_or_sync(kFull, c.mine ? c.lo : 0u);
                    hi = rhi;
                    lo = rlo;
                    // a popped state is never 0 (state >> 24 >= 2^8): both zero means nobody had the symbol
                    all_found = (rhi | rlo) == 0u ? 0u : all_found;
                    note_lane(j - i, c.mine);
                    account_pull(pulled);
                }
                if (all_found) return true;
                hi = s_hi; lo = s_lo; w_next = s_next; k = s_k;
                return false;
            };
            // The lane kernel's step by all lanes at once: a symbol without a window, or outside it.
            // Mean, scale and window origin come from the owning lane; every lane builds the model.
            auto scalar_step = [&](int j) {
                const float mu_j = __shfl_sync(kFull, d.mean, j), sc_j = __shfl_sync(kFull, d.scale, j);
                const int w0_j = __shfl_sync(kFull, d.w0, j);
                const SymbolModel mj = make_model(mu_j, sc_j);
                const bool pulled = hi == 0u;
                const uint32_t h = pulled ? lo : hi, l = pulled ? w_next : lo;
                hi = h;
                lo = l;
                account_pull(pulled);
                const int sym = decode_symbol_model(hi, lo, mu_j, sc_j, mj, tab, guard, flags);
                sts32(sm_code + 4u * (uint32_t)j, sym - (w0_j - 1));
                sts32(sm_lane + 4u * (uint32_t)j, 0);
            };
            // A symbol with a wide window: no search.  The continuous model is solved for the symbol
            // (guess_symbol(), the lane kernel's first step: float arithmetic on mod), and the exact
            // CDF values around the guess are READ from the rows instead of evaluated -- four of them,
            // so a guess that is one bin off still decides.  Every lane does the same: nothing crosses
            // lanes, and the chain is guess -> address -> load -> compare -> pop.
            const uint32_t sm_rows = sm + (uint32_t)kCoopSlotBase + (uint32_t)(g % Cfg::kBuffers) * (uint32_t)(kSlots * 256) + 8192u;
            // Its per-symbol constants sit in the owning lane's registers; they are fetched (six
            // shuffles) at the end of the PREVIOUS wide step, or at the start of the group for the first
            // one, so their latency is not in front of the chain.
            struct WidePar { float m05, c2; int w0, bins; uint32_t row; int j; };
            auto fetch_wide = [&](int jn) {
                WidePar q;
                const int src = jn < 0 ? 0 : jn;
                q.m05 = __shfl_sync(kFull, d.mean * 256.0f - 0.5f, src);
                q.c2 = __shfl_sync(kFull, d.scale * (256.0f * 0.693147181f), src);
                q.w0 = __shfl_sync(kFull, d.w0, src);
                q.bins = 32 * __shfl_sync(kFull, d.n, src);
                q.row = sm_rows + 128u * (uint32_t)__shfl_sync(kFull, d.t_off, src);
                q.j = jn;
                return q;
            };
            auto next_wide = [&](int below) -> int {        // highest wide symbol under slot `below`; -1: none
                const unsigned m = below > 0 ? grp.multis & ((1u << below) - 1u) : 0u;
                return 31 - __clz((int)m);
            };
            WidePar wp = fetch_wide(31 - __clz((int)grp.multis));       // the group's first wide symbol (-1: none)
            auto wide_step = [&](int j) -> bool {
                const WidePar q = wp.j == j ? wp : fetch_wide(j);
                const bool pulled = hi == 0u;
                const uint32_t h = pulled ? lo : hi, l = pulled ? w_next : lo;
                const uint32_t mod = l & kProbMask;
                // The symbol from the sigmoid term alone: u0 = logit((mod - 1024) / A), s = ceil(m - 1/2 + c u0)
                // (guess_symbol() without its Newton step: the linear term it corrects for moves the
                // answer by less than a bin at these scales, and a guess one bin off still decides).
                // Tails (mod - 1024 or A + 1024 - mod not positive) give a NaN or a wild index: not inside.
                float lp, lq;
                asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lp) : "f"((float)(mod - 1024u)));
                asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lq) : "f"((float)(16776192u - mod)));
                const float sr = __fmaf_rn(q.c2, lp - lq, q.m05);
                const int r = (int)(__float_as_uint(__fadd_ru(sr, 12582912.0f)) - 0x4b400000u) - q.w0;
                const bool inside = (uint32_t)(r - 2) < (uint32_t)(q.bins - 3);       // r - 2 .. r + 1 are rows' values
                const uint32_t a = q.row + 4u * (uint32_t)(inside ? r : 2);
                const int v0 = lds32(a - 8u), v1 = lds32(a - 4u), v2 = lds32(a), v3 = lds32(a + 4u);
                wp = fetch_wide(next_wide(j));              // for the next wide symbol of the group
                // the symbol is the bin whose interval [CDF(b - 1), CDF(b)) holds mod: r, r - 1 or r + 1
                const int mi = (int)mod;
                const bool at = v1 <= mi && mi < v2, left = v0 <= mi && mi < v1, right = v2 <= mi && mi < v3;
                if (!(inside && (at || left || right))) return false;
                const int p = at ? v1 : (left ? v0 : v2), v = at ? v2 : (left ? v1 : v3);
                hi = h;
                lo = l;
                rans_pop32(hi, lo, (uint32_t)p, (uint32_t)(v - p));
                account_pull(pulled);
                if (lane == 0) {
                    sts32(sm_code + 4u * (uint32_t)j, r + (at ? 0 : (left ? -1 : 1)) + 1);   // x = w0 - 1 + code + lane note
                    sts32(sm_lane + 4u * (uint32_t)j, 0);
                }
                return true;
            };
            int j = 31;
            while (j >= j_lo) {
                COOP_STAT(const long long t_i0 = clock64();)
                if (k > 27) { slide_words(); COOP_STAT(++n_slide;) }
                // single-chunk symbols in a row from j down, at most 4: a nibble table indexed by the top
                // four bits of `ones` shifted to j (no find-leading-one, which is a slow-pipe operation).
                // Plain compares and branches from here: an indirect branch (switch) costs the lone
                // in-order warp 70 cycles more per step.
                const unsigned top = (grp.ones << (31 - j)) >> 28;
                const int run = (int)((0x4322111100000000ull >> (4u * top)) & 7ull);
                if (run == 0) {
                    if ((grp.multis >> j) & 1u) {
                        if (wide_step(j)) { --j; COOP_STAT(++n_multi; t_multi += clock64() - t_i0;) continue; }
                    }
                    scalar_step(j);
                    COOP_STAT(++n_scalar; t_scalar += clock64() - t_i0;)
                    --j;
                    continue;
                }
                if (run >= 4) {
                    if (run_singles(std::integral_constant<int, 4>{}, j)) { j -= 4; COOP_STAT(n_quad += 4; t_quad += clock64() - t_i0;) continue; }
                    COOP_STAT(++n_quad_fail;)
                } else if (run == 3) {
                    if (run_singles(std::integral_constant<int, 3>{}, j)) { j -= 3; COOP_STAT(n_one += 3; t_one += clock64() - t_i0;) continue; }
                } else if (run == 2) {
                    if (run_singles(std::integral_constant<int, 2>{}, j)) { j -= 2; COOP_STAT(n_one += 2; t_one += clock64() - t_i0;) continue; }
                }
                // a run that failed as a whole is retried symbol by symbol; a symbol that is not in its
                // window takes the lane kernel's step
                if (run_singles(std::integral_constant<int, 1>{}, j)) { --j; COOP_STAT(++n_one; t_one += clock64() - t_i0;) continue; }
                scalar_step(j);
                COOP_STAT(++n_scalar; t_scalar += clock64() - t_i0;)
                --j;
            }
            COOP_STAT(const long long t_x0 = clock64();)
            __syncwarp();
            // s = (w0 - 1) + code + index of the matching lane
            if (lane >= j_lo)
                x_s[base + lane] = (float)(d.w0 - 1 + lds32(sm_code + 4u * (uint32_t)lane) + lds32(sm_lane + 4u * (uint32_t)lane)) * 0.00390625f;   // s / 256., exact
            __syncwarp();
            COOP_STAT(t_store += clock64() - t_x0;)
            if constexpr (C > 1) {
                COOP_STAT(const long long t_w0 = clock64();)
                if (g + 1 < n_groups) cluster_wait();                    // group g + 1 tabulated
                COOP_STAT(t_wait += clock64() - t_w0;)
            }
        }
        if constexpr (C == 1) cta_sync();
        COOP_STAT(if (lane == 0 && stream == 0) printf("coop stats C=%d: symbols %lld cycles %lld (%.1f/symbol) quad %lld (%.0f cyc) quad_fail %lld run<4 %lld (%.0f) multi %lld (%.0f) scalar %lld (%.0f) slides %lld; per group: wait %.0f setup %.0f store %.0f cycles\n",
                         C, (long long)len, clock64() - t_begin, (double)(clock64() - t_begin) / (double)(len > 0 ? len : 1), n_quad, (double)t_quad / (double)(n_quad ? n_quad : 1), n_quad_fail,
                         n_one, (double)t_one / (double)(n_one ? n_one : 1), n_multi, (double)t_multi / (double)(n_multi ? n_multi : 1), n_scalar, (double)t_scalar / (double)(n_scalar ? n_scalar : 1), n_slide,
                         (double)t_wait / n_groups, (double)t_setup / n_groups, (double)t_store / n_groups);)
        if (lane == 0) {
            const int wrem = wtop - k;                      // unread words; < 0: under-run
            flags |= guard_flags(guard);
            if (wrem < 0) flags |= ST_UNDERRUN;
            const uint64_t state = ((uint64_t)hi << 32) | lo;
            if (check_end && !too_long && (state != kRansL || wrem != 0)) flags |= ST_BAD_END_STATE;
            end_states[stream] = state;
            status[stream] = flags;
            if (left.out) left.out[stream] = wrem > 0 ? (int64_t)wrem : 0;
        }
    }
}

template <int C>
static cudaError_t launch_coop(const uint32_t* packed, const int64_t* word_offsets, const uint64_t* states,
                               const float* mean, const float* scale, const int64_t* offsets, int64_t n_streams,
                               float* x_out, uint64_t* end_states, int32_t* status, int check_end, WordsLeft left,
                               cudaStream_t stream) {
    using Cfg = CoopCfg<C>;
    cudaError_t e = cudaFuncSetAttribute(rans_decode_coop_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)Cfg::kSmemBytes);
    if (e != cudaSuccess) return e;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(n_streams * C));
    cfg.blockDim = dim3(Cfg::kWarps * 32);
    cfg.dynamicSmemBytes = Cfg::kSmemBytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = C;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = C > 1 ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, rans_decode_coop_kernel<C>, packed, word_offsets, states, mean, scale, offsets,
                              n_streams, x_out, end_states, status, check_end, left);
}

// Clusters of `cluster` CTAs that can be resident at once on the current device (0: not supported),
// cached per device and cluster size.
int64_t coop_cluster_capacity(int cluster) {
    constexpr int kMaxDevices = 64;
    static std::atomic<int> cache[kMaxDevices][4] = {};
    const int slot = cluster == 2 ? 1 : cluster == 4 ? 2 : cluster == 8 ? 3 : 0;
    if (slot == 0) return 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) dev = 0;
    int n = cache[dev][slot].load(std::memory_order_relaxed);
    if (n == 0) {
        cudaLaunchConfig_t cfg = {};
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = cluster;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        cfg.gridDim = dim3(cluster);
        cudaError_t e = cudaSuccess;
        int k = 0;
        auto query = [&](auto kernel, size_t smem, int threads) {
            e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            cfg.blockDim = dim3(threads);
            cfg.dynamicSmemBytes = smem;
            if (e == cudaSuccess) e = cudaOccupancyMaxActiveClusters(&k, kernel, &cfg);
        };
        if (cluster == 2) query(rans_decode_coop_kernel<2>, CoopCfg<2>::kSmemBytes, CoopCfg<2>::kWarps * 32);
        else if (cluster == 4) query(rans_decode_coop_kernel<4>, CoopCfg<4>::kSmemBytes, CoopCfg<4>::kWarps * 32);
        else query(rans_decode_coop_kernel<8>, CoopCfg<8>::kSmemBytes, CoopCfg<8>::kWarps * 32);
        if (e != cudaSuccess) { (void)cudaGetLastError(); k = 0; }
        n = k > 0 ? k : -1;
        cache[dev][slot].store(n, std::memory_order_relaxed);
    }
    return n > 0 ? n : 0;
}

// cluster: CTAs per stream (1, 2, 4 or 8)
cudaError_t launch_rans_decode_coop(const uint32_t* packed, const int64_t* word_offsets,
                                    const uint64_t* states, const float* mean, const float* scale,
                                    const int64_t* offsets, int64_t n_streams, float* x_out,
                                    uint64_t* end_states, int32_t* status, int check_end,
                                    WordsLeft left, int cluster, cudaStream_t stream) {
    if (n_streams <= 0) return cudaSuccess;
#define FLIC_COOP_ARGS packed, word_offsets, states, mean, scale, offsets, n_streams, x_out, end_states, status, check_end, left, stream
    switch (cluster) {
        case 8: return launch_coop<8>(FLIC_COOP_ARGS);
        case 4: return launch_coop<4>(FLIC_COOP_ARGS);
        case 2: return launch_coop<2>(FLIC_COOP_ARGS);
        default: return launch_coop<1>(FLIC_COOP_ARGS);
    }
#undef FLIC_COOP_ARGS
}

}  // namespace flic
