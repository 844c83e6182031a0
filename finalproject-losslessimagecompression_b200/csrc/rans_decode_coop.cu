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
// where the symbol is likely to be: a window of 2 x (6 scale units + 3) bins around the mode
// (99.5 % of a logistic's mass), at most 512 bins.  The CTA is split by role:
//   producers (7 warps)  for the next group of 32 symbols: per-symbol model, window size and
//       placement; then one exact CDF evaluation per lane per 32-bin chunk of every window
//       (cdf_at(), the same function every other kernel uses), written to shared memory.
//       Chunks of all 32 symbols form one task list that the warps share round-robin.
//   consumer (1 warp)    the chain.  All lanes hold the same state; per symbol: pull a word if
//       needed, mod = state & 0xffffff, each lane compares one tabulated CDF value with mod,
//       a ballot finds the first one above it, (start, end) are read back from the row and the
//       state is popped -- about a hundred cycles instead of a thousand.
// A symbol whose window was not tabulated (wide distribution, group over capacity, bad
// parameters) or that falls outside it is decoded on the chain by decode_symbol_lean() exactly as
// in the lane kernel, by all lanes at once.  Groups are double-buffered with one CTA barrier per
// group.  Results are bit-identical by construction: the window holds exact CDF values and the
// search returns the smallest in-window s with CDF(s) > mod (CDF is non-decreasing).
//
// K3d -- the same with a thread-block CLUSTER per stream (a handful of streams: the reference's
// own partition is one stream per latent level, trainer.py:308-318, and one stream in
// rans/test.py).  With one CTA the producers keep up only with narrow distributions, so symbols
// wider than 12 bins per scale unit (a quarter of rans/test.py's mixture) are decoded on the chain
// at ~1000 cycles each.  A cluster of 2, 4 or 8 CTAs puts 16 producer warps on each of the other
// SMs of the cluster; they write their chunks straight into the home CTA's shared memory
// (st.shared::cluster through a mapa-translated address, distributed shared memory), groups are
// handed over with the cluster barrier, and the window grows to 32 chunks (1024 bins, 5.1 scale
// units either side up to a scale of 100 bins), which the chain searches in two ballots (chunk
// index, then entry).  The home CTA keeps the chain's scheduler free: its own producers, if any,
// are the warps of the other three sub-partitions.
#include "flic_device.cuh"
#include "flic_kernels.cuh"

#include <type_traits>

namespace flic {

#ifndef FLIC_COOP_UNROLL
#define FLIC_COOP_UNROLL 4
#endif
#ifndef FLIC_CLUSTER_HOME_PRODUCERS
#define FLIC_CLUSTER_HOME_PRODUCERS 1
#endif
constexpr int kCoopUnroll = FLIC_COOP_UNROLL;   // evaluations a producer warp interleaves
constexpr unsigned kFull = 0xffffffffu;
constexpr int kCoopSlotBase = 2560;     // bytes of shared memory in front of the slots
constexpr int kNever = 0x7fffffff;    // p = v = kNever: an empty interval no mod falls into

// Geometry by cluster size C (CTAs per stream).
template <int C>
struct CoopCfg {
    // C == 1: warps 0..6 produce, warp 7 is the consumer (more producers take issue slots from the
    // chain: 15 cost it 17 %); two CTAs fit an SM.  C > 1: 16 warps per CTA, the last one of the home
    // CTA is the consumer, alone on its sub-partition.
    static constexpr int kWarps = C == 1 ? 8 : 16;
    static constexpr int kConsumerWarp = kWarps - 1;
    static constexpr int kHomeProducers = C == 1 ? 7 : (FLIC_CLUSTER_HOME_PRODUCERS ? 12 : 0);
    static constexpr int kProducers = kHomeProducers + (C - 1) * kWarps;   // per stream
    static constexpr int kSlots = C == 1 ? 160 : 448;     // 32-entry slots per group buffer (two buffers)
    static constexpr int kMaxChunks = C == 1 ? 4 : 32;    // 32-bin chunks per symbol window
    static constexpr size_t kSmemBytes = kCoopSlotBase + 8u * 2 * kSlots * 32;
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
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
template <int C>
__device__ __forceinline__ void group_sync() {
    if constexpr (C == 1) cta_sync();
    else cluster_sync_all();
}
// one tabulated value into entry `e` of the home CTA's slot memory (byte address `a` of the entry)
template <int C>
__device__ __forceinline__ void st_slot(uint32_t a, int v) {
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
    int n;       // chunks tabulated; 0: none (window around the guess, on the chain); -1: bad parameters
                 // (scalar path, which flags them)
    int t_off;   // n > 1: first chunk slot; entry e of chunk c stands for bin w0 - 1 + 32 c + e
    int w0;
};
struct CoopGroup {
    unsigned ones;   // symbols with n == 1
    unsigned multis; // symbols with n > 1
    int n_ones;      // their number: tasks [0, n_ones) evaluate them, tasks from n_ones on the chunks >= 32
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
// chunks cover 5.1 scale units up to a scale of 100 bins, and 3.4 at the 148 bins that end
// rans/test.py's range).
template <int C>
__device__ __forceinline__ CoopGroup coop_describe(float mean, float scale, bool valid, int lane, CoopSymbol& d) {
    using Cfg = CoopCfg<C>;
    d.mean = mean;
    d.scale = scale;
    d.m = make_model(mean, scale);
    const float cb = scale * 256.0f;                // logistic scale in bins
    const bool ok = valid && params_ok(mean, scale);
    int n = 0;
    if (ok && cb <= 3.0f) n = 1;
    else if (ok && (C > 1 || cb <= 12.0f)) {
        n = (2 * ((int)(fminf(cb, 128.0f) * 5.1f) + 2) + 31) >> 5;  // C == 1: <= (2 * 63 + 31) / 32 = 4
        n = n < Cfg::kMaxChunks ? n : Cfg::kMaxChunks;
    }
    const int many = n > 1 ? n : 0;
    int incl = many;
#pragma unroll
    for (int s = 1; s < 32; s <<= 1) {
        const int o = __shfl_up_sync(kFull, incl, s);
        if (lane >= s) incl += o;
    }
    if (many && 32 + incl > Cfg::kSlots) n = 0;     // over capacity: decoded on the chain
    d.t_off = 32 + incl - many;
    d.n = ok ? n : -1;
    // centred on the mode (lower + 1024 = round(256 mean)): with at most 512 bins either side the
    // window lies strictly inside the coder's 2048-bin support, so it needs no edge cases
    d.w0 = d.m.lower + 1024 - 16 * n + 1;
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
    int* const s_code = reinterpret_cast<int*>(s_raw + 512);                                 // 32 x 4 B (consumer only)
    unsigned* const s_who = reinterpret_cast<unsigned*>(s_raw + 640);                        // 32 x 4 B (consumer only)
    int* const s_toff = reinterpret_cast<int*>(s_raw + 768);                                 // 32 x 4 B (consumer only)
    int4* const s_par = reinterpret_cast<int4*>(s_raw + 1024);                               // 32 x 48 B (consumer only)
    // slots: [2][kSlots][32] CoopEntry from s_raw + kCoopSlotBase, in the HOME CTA (cluster rank 0)
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
    if constexpr (C > 1) cluster_sync_all();

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

    if (producer >= 0) {
        // ------------------------------------------------------------------ producers
        uint32_t slot_base = (uint32_t)__cvta_generic_to_shared(s_raw) + (uint32_t)kCoopSlotBase;
        if constexpr (C > 1) slot_base = cluster_map(slot_base, 0u);
        float mu, sc;
        bool valid = load_group(0, mu, sc);
        for (int g = 0; g < n_groups; ++g) {
            CoopSymbol d;
            const CoopGroup grp = coop_describe<C>(mu, sc, valid, lane, d);
            valid = load_group(g + 1, mu, sc);          // in flight during the evaluations
            const uint32_t buf = slot_base + (uint32_t)(g & 1) * (uint32_t)(kSlots * 256);
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
                    } else {                            // a chunk of a multi-chunk symbol
                        sl = 32 + (t - grp.n_ones);
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
                    bins[u] = __shfl_sync(kFull, d.w0, j) - 1 + 32 * c + lane;
                }
                int v[U];
#pragma unroll
                for (int u = 0; u < U; ++u) v[u] = cdf_at(bins[u], mj[u], tab);
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    if (!on[u]) continue;
                    const int c = cs[u], n = ns[u];
                    // entry e of slot sl: byte address buf + 256 sl + 8 e; .p at +0, .v at +4
                    const uint32_t row = buf + 256u * (uint32_t)slots[u] + 8u * (uint32_t)lane;
                    // v is this entry's value and the next entry's p; the window's very first entry
                    // has no left neighbour and can never be the symbol (empty interval p == v; the
                    // chain tests (mod - p) < (v - p) in unsigned arithmetic)
                    st_slot<C>(row + 4u, v[u]);
                    if (lane < 31 || c + 1 < n) st_slot<C>(row + 8u, v[u]);   // lane 31: entry 0 of the next chunk
                    if (lane == 0 && c == 0) st_slot<C>(row, v[u]);            // p == v: an empty interval
                    if (n > 1) {
                        // index (head slot): entry c = (last value of chunk c - 1, last value of chunk c), so that
                        // p <= mod < v picks the chunk that holds the symbol; entries >= n never match
                        const uint32_t index = buf + 256u * (uint32_t)js[u];
                        if (lane == 31) {
                            st_slot<C>(index + 8u * (uint32_t)c + 4u, v[u]);
                            if (c + 1 < n) st_slot<C>(index + 8u * (uint32_t)(c + 1), v[u]);
                        }
                        if (c == 0) {
                            if (lane == 0) st_slot<C>(index, -1);
                            if (lane >= n) { st_slot<C>(index + 8u * (uint32_t)lane, kNever); st_slot<C>(index + 8u * (uint32_t)lane + 4u, kNever); }
                        }
                    }
                }
            };
            for (int t0 = producer; t0 < grp.total; t0 += kCoopUnroll * Cfg::kProducers) {
                if (t0 + Cfg::kProducers >= grp.total) run_tasks(std::integral_constant<int, 1>{}, t0);
                else run_tasks(std::integral_constant<int, kCoopUnroll>{}, t0);
            }
            group_sync<C>();   // group g tabulated; the consumer has finished group g - 1
        }
        group_sync<C>();       // pairs with the consumer's barrier after the last group
    } else if (!is_consumer) {
        // ------------------------------------------------------------------ idle warps of the home CTA
        for (int g = 0; g <= n_groups; ++g) group_sync<C>();
    } else {
        // ------------------------------------------------------------------ consumer
        const uint32_t* const wbase = packed + wbeg;
        int wrem = too_long ? 0 : (int)wcount;          // unread words (w_next and w_after included); < 0: under-run
        const uint64_t st0 = states[stream];
        uint32_t hi = (uint32_t)(st0 >> 32), lo = (uint32_t)st0;
        // the next two words of the stream wait in registers: a pull takes w_next, and the load that
        // replaces w_after has a whole symbol or more to arrive before it can be wanted
        uint32_t w_next = wrem > 0 ? __ldg(wbase + (wrem - 1)) : 0u;
        uint32_t w_after = wrem > 1 ? __ldg(wbase + (wrem - 2)) : 0u;
        int32_t flags = too_long ? ST_TOO_LONG : 0;
        ParamGuard guard = guard_init();
        float* const x_s = x_out + beg;
        // shared-space addresses (a generic pointer costs the shared-window base at every use)
        uint32_t sm = (uint32_t)__cvta_generic_to_shared(s_raw);
        asm volatile("" : "+r"(sm));
        const uint32_t sm_who = sm + 640u;
        auto sts_who = [&](int j, unsigned v) {
            asm volatile("st.shared.u32 [%0], %1;" ::"r"(sm_who + 4u * (uint32_t)j), "r"(v) : "memory");
        };
        auto lds_entry = [&](uint32_t addr) -> CoopEntry {
            CoopEntry e;
            asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(e.p), "=r"(e.v) : "r"(addr) : "memory");
            return e;
        };

        // rans.pyx:87-89, warp-uniform, in two halves: the selects sit on the chain, the count and
        // the conditional load of the word after next do not
        auto pull_state = [&]() -> uint32_t {
            const uint32_t hi0 = hi;
            hi = hi0 == 0u ? lo : hi0;
            lo = hi0 == 0u ? w_next : lo;
            return hi0;
        };
        auto pull_refill = [&](uint32_t hi0) {
            wrem -= hi0 == 0u ? 1 : 0;
            w_next = hi0 == 0u ? w_after : w_next;
            // if (hi0 == 0 && wrem >= 2) w_after = wbase[wrem - 2], as a predicated load
            asm volatile("{\n\t.reg .pred p;\n\t.reg .u64 a;\n\t"
                         "setp.eq.u32 p, %1, 0;\n\t"
                         "setp.gt.and.s32 p, %2, 1, p;\n\t"
                         "mad.wide.s32 a, %2, 4, %3;\n\t"
                         "@p ld.global.nc.u32 %0, [a+-8];\n\t}"
                         : "+r"(w_after) : "r"(hi0), "r"(wrem), "l"(wbase));
        };
        // The entry with p <= mod < v (and `live`), if this warp holds it, pops the state; everybody
        // takes the result.  Returns the ballot of the entry's lane (0: nobody holds it, state
        // unchanged).  A popped state is never 0 (state >> 24 >= 2^8), so the OR-reduction of the one
        // lane's candidate is the broadcast.
        auto take = [&](const CoopEntry e, int mod, bool live) -> unsigned {
            const bool mine = live && e.p <= mod && mod < e.v;
            const unsigned who = __ballot_sync(kFull, mine);
            uint32_t chi = hi, clo = lo;
            rans_pop32(chi, clo, (uint32_t)e.p, (uint32_t)(e.v - e.p));
            const uint32_t rhi = __reduce_or_sync(kFull, mine ? chi : 0u);
            const uint32_t rlo = __reduce_or_sync(kFull, mine ? clo : 0u);
            if (who) { hi = rhi; lo = rlo; }
            return who;
        };

        float mu, sc;
        bool valid = load_group(0, mu, sc);
        for (int g = 0; g < n_groups; ++g) {
            CoopSymbol d;
            const CoopGroup grp = coop_describe<C>(mu, sc, valid, lane, d);
            const int64_t base = len - 32 * (int64_t)(g + 1);
            const int j_lo = base < 0 ? (int)(-base) : 0;   // first group-slot that is a symbol
            valid = load_group(g + 1, mu, sc);
            // for the symbols that leave the fast path: model, parameters, n, first chunk slot, window origin
            s_par[3 * lane + 0] = make_int4(__double2loint(d.m.mean_d), __double2hiint(d.m.mean_d),
                                            __double2loint(d.m.scale_d), __double2hiint(d.m.scale_d));
            s_par[3 * lane + 1] = make_int4(__double2loint(d.m.rscale), __double2hiint(d.m.rscale), d.m.lower, d.w0);
            s_par[3 * lane + 2] = make_int4(__float_as_int(d.mean), __float_as_int(d.scale), d.n, d.t_off);
            s_code[lane] = 0;
            s_toff[lane] = d.t_off;
            __syncwarp();
            group_sync<C>();   // group g tabulated
            // A warp issues in order: whatever sits in front of an instruction in the stream delays
            // it, needed or not.  So the common case -- a single-chunk symbol found in its window --
            // gets a loop of its own with nothing else in it: the head entry is read one symbol ahead
            // (fixed address: slot j), "single-chunk" is a bit of `ones`, the decoded value is worked
            // out after the loop from the ballot each step leaves behind, and anything else LEAVES
            // the loop, is decoded out of line, and the loop is entered again.
            const uint32_t sm_slot = sm + (uint32_t)kCoopSlotBase + (uint32_t)(g & 1) * (uint32_t)(kSlots * 256) + 8u * (uint32_t)lane;
            // One single-chunk symbol: this lane's candidate for the popped state and whether the
            // lane's entry is the symbol's.  (h, l) is the state after the word pull.
            struct Cand { uint32_t hi, lo; bool mine; };
            auto candidate = [&](const CoopEntry e, uint32_t h, uint32_t l) -> Cand {
                const uint32_t dm = (l & kProbMask) - (uint32_t)e.p;                           // mod - start
                const uint32_t fr = (uint32_t)(e.v - e.p);
                // pop as if this lane's entry were the one: x = state >> 24, new = x freq + (mod - start)
                const uint32_t xl = (h << 8) | (l >> 24), xh = h >> 24;
                const uint64_t pr = (uint64_t)xl * fr + dm;
                Cand c;
                c.hi = (uint32_t)(pr >> 32) + xh * fr;
                c.lo = (uint32_t)pr;
                c.mine = dm < fr;
                return c;
            };
            // A symbol without a window, or outside it: the lane kernel's step (guess, two exact
            // evaluations, bracket search if the guess is off), by all lanes at once, with the model
            // the group description already holds.
            auto scalar_step = [&](int j) {
                const int4 q0 = s_par[3 * j], q1 = s_par[3 * j + 1], q2 = s_par[3 * j + 2];
                const uint32_t hi0 = pull_state();
                pull_refill(hi0);
                SymbolModel mj;
                mj.mean_d = __hiloint2double(q0.y, q0.x);
                mj.scale_d = __hiloint2double(q0.w, q0.z);
                mj.rscale = __hiloint2double(q1.y, q1.x);
                mj.lower = q1.z;
                const int sym = decode_symbol_model(hi, lo, __int_as_float(q2.x), __int_as_float(q2.y), mj, tab, guard, flags);
                s_code[j] = sym - (q1.w - 1);
                sts_who(j, 1u);
            };
            // A multi-chunk symbol: its head slot indexes the chunks (entry c = last values of chunks
            // c - 1 and c), so one ballot picks the chunk and a second one the entry.
            auto multi_step = [&](int j) -> bool {
                const CoopEntry index = lds_entry(sm_slot + 256u * (uint32_t)j);
                const int t_off = s_toff[j];
                const uint32_t hi0 = hi, h = hi0 == 0u ? lo : hi0, l = hi0 == 0u ? w_next : lo;   // pulled state
                const int mod = (int)(l & kProbMask);
                const unsigned which = __ballot_sync(kFull, index.p <= mod && mod < index.v);
                const int c = __popc(which - 1u) & 31;       // one bit set: its index
                const Cand cd = candidate(lds_entry(sm_slot + 256u * (uint32_t)(t_off + c)), h, l);
                const unsigned who = __ballot_sync(kFull, cd.mine);
                const uint32_t rhi = __reduce_or_sync(kFull, cd.mine ? cd.hi : 0u);
                const uint32_t rlo = __reduce_or_sync(kFull, cd.mine ? cd.lo : 0u);
                if (which == 0u || who == 0u) return false;
                pull_refill(hi0);
                sts_who(j, who);
                s_code[j] = 32 * c;
                hi = rhi;
                lo = rlo;
                return true;
            };
            int j = 31;
            if (grp.total == 0) {
                // nothing tabulated in this group (wide distributions): the lane kernel's step for every
                // symbol, its model and parameters fetched from the owning lane one symbol ahead
                struct Par { SymbolModel m; float mean, scale; };
                auto par_of = [&](int k) {
                    const int4 q0 = s_par[3 * k], q1 = s_par[3 * k + 1], q2 = s_par[3 * k + 2];
                    Par q;
                    q.m.mean_d = __hiloint2double(q0.y, q0.x);
                    q.m.scale_d = __hiloint2double(q0.w, q0.z);
                    q.m.rscale = __hiloint2double(q1.y, q1.x);
                    q.m.lower = q1.z;
                    q.mean = __int_as_float(q2.x);
                    q.scale = __int_as_float(q2.y);
                    return q;
                };
                Par pa = par_of(31), pb;
#pragma unroll 1
                for (; j >= j_lo; j -= 2) {
                    pb = par_of(j > 0 ? j - 1 : 0);
                    uint32_t hi0 = pull_state();
                    pull_refill(hi0);
                    const int sa = decode_symbol_model(hi, lo, pa.mean, pa.scale, pa.m, tab, guard, flags);
                    if (lane == j) s_code[j] = sa - (d.w0 - 1);
                    if (j - 1 < j_lo) { --j; break; }
                    pa = par_of(j > 1 ? j - 2 : 0);
                    hi0 = pull_state();
                    pull_refill(hi0);
                    const int sb = decode_symbol_model(hi, lo, pb.mean, pb.scale, pb.m, tab, guard, flags);
                    if (lane == j - 1) s_code[j - 1] = sb - (d.w0 - 1);
                }
                // every symbol of such a group is "entry 0 of its code"
                sts_who(lane, 1u);
                j = j_lo - 1;
            }
            while (j >= j_lo) {
                const unsigned w = grp.ones << (31 - j);     // symbol j - k at bit 31 - k
                if ((w >> 28) == 0xfu) {
                    // Four single-chunk symbols in a row: decoded optimistically in one straight line
                    // (no branch, so nothing between one symbol's broadcast and the next one's
                    // search), and taken back if any of them fell outside its window.
                    const uint32_t a0 = sm_slot + 256u * (uint32_t)j;
                    const CoopEntry e0 = lds_entry(a0), e1 = lds_entry(a0 - 256u), e2 = lds_entry(a0 - 512u),
                                    e3 = lds_entry(a0 - 768u);
                    const uint32_t s_hi = hi, s_lo = lo, s_next = w_next, s_after = w_after;
                    const int s_wrem = wrem;
                    unsigned who[4];
                    const CoopEntry es[4] = {e0, e1, e2, e3};
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint32_t hi0 = pull_state();
                        const Cand c = candidate(es[k], hi, lo);
                        who[k] = __ballot_sync(kFull, c.mine);
                        const uint32_t rhi = __reduce_or_sync(kFull, c.mine ? c.hi : 0u);
                        const uint32_t rlo = __reduce_or_sync(kFull, c.mine ? c.lo : 0u);
                        pull_refill(hi0);
                        hi = rhi;
                        lo = rlo;
                    }
                    if (who[0] != 0u && who[1] != 0u && who[2] != 0u && who[3] != 0u) {
#pragma unroll
                        for (int k = 0; k < 4; ++k) sts_who(j - k, who[k]);
                        j -= 4;
                        continue;
                    }
                    hi = s_hi; lo = s_lo; w_next = s_next; w_after = s_after; wrem = s_wrem;
                }
                if (w & 0x80000000u) {
                    // one single-chunk symbol
                    const CoopEntry e = lds_entry(sm_slot + 256u * (uint32_t)j);
                    const uint32_t hi0 = hi, h = hi0 == 0u ? lo : hi0, l = hi0 == 0u ? w_next : lo;
                    const Cand c = candidate(e, h, l);
                    const unsigned who = __ballot_sync(kFull, c.mine);
                    const uint32_t rhi = __reduce_or_sync(kFull, c.mine ? c.hi : 0u);
                    const uint32_t rlo = __reduce_or_sync(kFull, c.mine ? c.lo : 0u);
                    if (who != 0u) {
                        pull_refill(hi0);
                        sts_who(j, who);
                        hi = rhi;
                        lo = rlo;
                        --j;
                        continue;
                    }
                } else if ((grp.multis >> j) & 1u) {
                    if (multi_step(j)) { --j; continue; }
                }
                scalar_step(j);
                --j;
            }
            __syncwarp();
            // s = (w0 - 1) + code + index of the matching lane
            if (lane >= j_lo)
                x_s[base + lane] = (float)(d.w0 - 1 + s_code[lane] + __ffs((int)s_who[lane]) - 1) * 0.00390625f;   // s / 256., exact
            __syncwarp();
        }
        group_sync<C>();
        if (lane == 0) {
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
