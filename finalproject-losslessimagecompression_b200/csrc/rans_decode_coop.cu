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
#include "flic_device.cuh"
#include "flic_kernels.cuh"

namespace flic {

constexpr int kCoopProducers = 7;     // producer warps per CTA; warp kCoopProducers is the consumer (more
                                      // producers take issue slots from the chain: 15 cost it 17 %)
constexpr int kCoopSlots = 160;       // 32-entry slots per group buffer (40 KB; two buffers)
#ifndef FLIC_COOP_UNROLL
#define FLIC_COOP_UNROLL 4
#endif
constexpr int kCoopUnroll = FLIC_COOP_UNROLL;   // evaluations a producer warp interleaves
constexpr int kCoopMaxChunks = 8;     // 32-bin chunks per symbol window (at most 256 bins)
constexpr unsigned kFull = 0xffffffffu;
constexpr int kCoopSlotBase = 2560;     // bytes of shared memory in front of the slots
constexpr int kNever = 0x7fffffff;    // p = v = kNever: an empty interval no mod falls into

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

// Every warp of the CTA computes the same description from the same inputs, so no metadata has
// to cross warps.
//
// Window: one chunk (15 bins either side of the mode) up to a logistic scale of 3 bins (1.3 % of
// a logistic's mass lies outside, and a symbol there costs a few hundred cycles on the chain);
// up to 12 bins, 5.1 scale units either side in 2 to 4 chunks, which the chain searches in two
// steps; wider distributions are not tabulated: at ~45 cycles per chunk (7 producer warps) the
// tabulation would take as long as the lane kernel's step, which is what those symbols get.
__device__ __forceinline__ CoopGroup coop_describe(float mean, float scale, bool valid, int lane, CoopSymbol& d) {
    d.mean = mean;
    d.scale = scale;
    d.m = make_model(mean, scale);
    const float cb = scale * 256.0f;                // logistic scale in bins
    const bool ok = valid && params_ok(mean, scale);
    int n = 0;
    if (ok && cb <= 3.0f) n = 1;
    else if (ok && cb <= 12.0f) n = (2 * ((int)(cb * 5.1f) + 2) + 31) >> 5;  // <= (2 * 63 + 31) / 32 = 4
    const int many = n > 1 ? n : 0;
    int incl = many;
#pragma unroll
    for (int s = 1; s < 32; s <<= 1) {
        const int o = __shfl_up_sync(kFull, incl, s);
        if (lane >= s) incl += o;
    }
    if (32 + incl > kCoopSlots) n = 0;              // over capacity: decoded on the chain
    d.t_off = 32 + incl - many;
    d.n = ok ? n : -1;
    // centred on the mode (lower + 1024 = round(256 mean)): with at most 256 bins either side the
    // window lies strictly inside the coder's 2048-bin support, so it needs no edge cases
    d.w0 = d.m.lower + 1024 - 16 * n + 1;
    CoopGroup grp;
    grp.ones = __ballot_sync(kFull, n == 1);
    grp.multis = __ballot_sync(kFull, n > 1);
    grp.n_ones = __popc(grp.ones);
    grp.total = grp.n_ones + __shfl_sync(kFull, incl, 31);
    return grp;
}

__global__ void __launch_bounds__((kCoopProducers + 1) * 32)
rans_decode_coop_kernel(const uint32_t* __restrict__ packed, const int64_t* __restrict__ word_offsets,
                        const uint64_t* __restrict__ states, const float* __restrict__ mean,
                        const float* __restrict__ scale, const int64_t* __restrict__ offsets,
                        int64_t n_streams, float* __restrict__ x_out, uint64_t* __restrict__ end_states,
                        int32_t* __restrict__ status, int check_end, WordsLeft left) {
    extern __shared__ __align__(256) unsigned char s_raw[];
    uint64_t* const s_tab = reinterpret_cast<uint64_t*>(s_raw);                              // 256 B
    int* const s_code = reinterpret_cast<int*>(s_raw + 512);                                 // 32 x 4 B (consumer only)
    unsigned* const s_who = reinterpret_cast<unsigned*>(s_raw + 640);                        // 32 x 4 B (consumer only)
    int* const s_toff = reinterpret_cast<int*>(s_raw + 768);                                 // 32 x 4 B (consumer only)
    int4* const s_par = reinterpret_cast<int4*>(s_raw + 1024);                               // 32 x 48 B (consumer only)
    CoopEntry (*const s_slot)[kCoopSlots][32] =
        reinterpret_cast<CoopEntry (*)[kCoopSlots][32]>(s_raw + kCoopSlotBase);              // [2][slots][32]
    const ExpTab tab = stage_exp_table(s_tab);

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int64_t stream = blockIdx.x;
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

    // group g holds the symbols [len - 32 (g + 1), len - 32 g) that exist; lane j <-> the j-th of them
    auto load_group = [&](int g, float& mu, float& sc) -> bool {
        const int64_t i = len - 32 * (int64_t)(g + 1) + lane;
        const bool valid = g < n_groups && i >= 0;
        mu = valid ? __ldg(mean_s + i) : 0.0f;
        sc = valid ? __ldg(scale_s + i) : 1.0f;
        return valid;
    };

    if (warp < kCoopProducers) {
        // ------------------------------------------------------------------ producers
        float mu, sc;
        bool valid = load_group(0, mu, sc);
        for (int g = 0; g < n_groups; ++g) {
            CoopSymbol d;
            const CoopGroup grp = coop_describe(mu, sc, valid, lane, d);
            valid = load_group(g + 1, mu, sc);          // in flight during the evaluations
            CoopEntry (*const slot)[32] = s_slot[g & 1];
            // kCoopUnroll tasks per pass: their evaluations are independent and interleave (a lone
            // warp is latency-bound on the ~35-deep dependency chain of one evaluation)
            for (int t0 = warp; t0 < grp.total; t0 += kCoopUnroll * kCoopProducers) {
                int bins[kCoopUnroll], cs[kCoopUnroll], ns[kCoopUnroll], js[kCoopUnroll], slots[kCoopUnroll];
                bool on[kCoopUnroll];
                SymbolModel mj[kCoopUnroll];
#pragma unroll
                for (int u = 0; u < kCoopUnroll; ++u) {
                    const int t = t0 + u * kCoopProducers;
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
                int v[kCoopUnroll];
#pragma unroll
                for (int u = 0; u < kCoopUnroll; ++u) v[u] = cdf_at(bins[u], mj[u], tab);
#pragma unroll
                for (int u = 0; u < kCoopUnroll; ++u) {
                    if (!on[u]) continue;
                    const int c = cs[u], n = ns[u];
                    CoopEntry* const row = slot[slots[u]];
                    // v is this entry's value and the next entry's p; the window's very first entry
                    // has no left neighbour and can never be the symbol (empty interval p == v; the
                    // chain tests (mod - p) < (v - p) in unsigned arithmetic)
                    row[lane].v = v[u];
                    if (lane < 31) row[lane + 1].p = v[u];
                    else if (c + 1 < n) row[32].p = v[u];          // entry 0 of the next chunk
                    if (lane == 0 && c == 0) row[0].p = v[u];       // p == v: an empty interval
                    if (n > 1) {
                        // index (head slot): entry c = (last value of chunk c - 1, last value of chunk c), so that
                        // p <= mod < v picks the chunk that holds the symbol; entries >= n never match
                        CoopEntry* const index = slot[js[u]];
                        if (lane == 31) {
                            index[c].v = v[u];
                            if (c + 1 < n) index[c + 1].p = v[u];
                        }
                        if (c == 0) {
                            if (lane == 0) index[0].p = -1;
                            if (lane >= n) { index[lane].p = kNever; index[lane].v = kNever; }
                        }
                    }
                }
            }
            cta_sync();   // group g tabulated; the consumer has finished group g - 1
        }
        cta_sync();       // pairs with the consumer's barrier after the last group
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
            const CoopGroup grp = coop_describe(mu, sc, valid, lane, d);
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
            cta_sync();   // group g tabulated
            const CoopEntry (*const slot)[32] = s_slot[g & 1];
            // A warp issues in order: whatever sits in front of an instruction in the stream delays
            // it, needed or not.  So the common case -- a single-chunk symbol found in its window --
            // gets a loop of its own with nothing else in it: the head entry is read one symbol ahead
            // (fixed address: slot j), "single-chunk" is a bit of `ones`, the decoded value is worked
            // out after the loop from the ballot each step leaves behind, and anything else LEAVES
            // the loop, is decoded out of line, and the loop is entered again.
            const uint32_t sm_slot = sm + (uint32_t)kCoopSlotBase + (uint32_t)(g & 1) * (uint32_t)(kCoopSlots * 256) + 8u * (uint32_t)lane;
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
                const int c = __popc(which - 1u) & (kCoopMaxChunks - 1);       // one bit set: its index
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
        cta_sync();
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

constexpr size_t kCoopSmemBytes = kCoopSlotBase + sizeof(CoopEntry) * 2 * kCoopSlots * 32;

cudaError_t launch_rans_decode_coop(const uint32_t* packed, const int64_t* word_offsets,
                                    const uint64_t* states, const float* mean, const float* scale,
                                    const int64_t* offsets, int64_t n_streams, float* x_out,
                                    uint64_t* end_states, int32_t* status, int check_end,
                                    WordsLeft left, cudaStream_t stream) {
    if (n_streams <= 0) return cudaSuccess;
    cudaError_t e = cudaFuncSetAttribute(rans_decode_coop_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)kCoopSmemBytes);
    if (e != cudaSuccess) return e;
    rans_decode_coop_kernel<<<(unsigned)n_streams, (kCoopProducers + 1) * 32, kCoopSmemBytes, stream>>>(
        packed, word_offsets, states, mean, scale, offsets, n_streams, x_out, end_states, status, check_end, left);
    return cudaGetLastError();
}

}  // namespace flic
