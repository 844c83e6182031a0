// tests/host_harness.cpp -- compiles the device arithmetic header (csrc/flic_core.cuh) with plain
// g++ so that the search / rANS / division logic can be compared with the oracle on the CPU.
// TEST INFRASTRUCTURE ONLY: the product library never runs these functions on the host and this
// file is not part of it.  Built by tests/conftest.py into tests/_build/libflic_host.so with
// -ffp-contract=off (the header's dadd/dmul/... wrappers are plain operators on the host).
#include "flic_core.cuh"

#include <stdint.h>
#include <stdlib.h>

using namespace flic;

static const uint64_t kTab[32] = {FLIC_EXP2F_TABLE};

extern "C" {

float hh_expf(float x) { return expf_glibc(x, kTab); }

void hh_expf_array(uint32_t lo_bits, int64_t n, float* out) {
    for (int64_t i = 0; i < n; ++i) {
        uint32_t u = lo_bits + (uint32_t)i;
        float x;
        memcpy(&x, &u, 4);
        out[i] = expf_glibc(x, kTab);
    }
}

int hh_lower(float mean) { return lower_of(mean); }

int hh_part1(float arg) { return part1_from_arg(arg_from_quotient((double)arg), kTab); }

int hh_cdf(int s, float mean, float scale) {
    SymbolModel m = make_model(mean, scale);
    return cdf_at(s, m, kTab);
}

int hh_guess(uint32_t mod, float mean, float scale) { return guess_symbol(mod, mean, scale, lower_of(mean)); }

int hh_tables(const float* x, const float* mean, const float* scale, int64_t n, uint32_t* start,
              uint32_t* freq) {
    int32_t flags = 0;
    for (int64_t i = 0; i < n; ++i) {
        SymbolTable t = make_table(x[i], mean[i], scale[i], kTab, flags);
        start[i] = t.start;
        freq[i] = t.freq;
    }
    return flags;
}

// One stream, forward order, words in emission order.
int hh_encode(const float* x, const float* mean, const float* scale, int64_t n, uint32_t* words,
              int64_t* n_words, uint64_t* state_out) {
    uint64_t state = kRansL;
    int64_t nw = 0;
    int32_t flags = 0;
    for (int64_t i = 0; i < n; ++i) {
        SymbolTable t = make_table(x[i], mean[i], scale[i], kTab, flags);
        uint32_t w;
        if (rans_push(state, t.start, t.freq, w)) words[nw++] = w;
    }
    *n_words = nw;
    *state_out = state;
    return flags;
}

// One stream; mean/scale in FORWARD order, words in emission order; decodes back to front.
int hh_decode(const uint32_t* words, int64_t n_words, uint64_t state, const float* mean,
              const float* scale, int64_t n, float* out, uint64_t* end_state, int64_t* evals) {
    int64_t pos = n_words;
    int32_t flags = 0;
    int64_t ne = 0;
    for (int64_t i = n - 1; i >= 0; --i) {
        if (state < kRansL) {
            if (pos <= 0) { flags |= ST_UNDERRUN; break; }
            state = (state << 32) | words[--pos];
        }
        const uint32_t mod = (uint32_t)state & kProbMask;
        SymbolModel m = make_model(mean[i], scale[i]);
        flags |= param_flags(mean[i], scale[i]);
        SearchState st = search_begin(mod, mean[i], scale[i], m);
        while (!st.done) {
            const int c = cdf_at(st.probe, m, kTab);
            ++ne;
            search_feed(st, c, mod);
        }
        if (st.hi > m.lower + (kWindow - 1)) flags |= ST_NO_SYMBOL;
        out[i] = (float)st.hi * 0.00390625f;
        rans_pop(state, (uint32_t)st.c_lo, (uint32_t)(st.c_hi - st.c_lo));
    }
    *end_state = state;
    *evals = ne;
    if (state != kRansL) flags |= ST_BAD_END_STATE;
    return flags;
}

// Same as hh_decode but through decode_symbol(), the exact function the CUDA kernel calls.
int hh_decode_fast(const uint32_t* words, int64_t n_words, uint64_t state, const float* mean,
                   const float* scale, int64_t n, float* out, uint64_t* end_state) {
    int64_t pos = n_words;
    int32_t flags = 0;
    for (int64_t i = n - 1; i >= 0; --i) {
        if (state < kRansL) {
            if (pos <= 0) { flags |= ST_UNDERRUN; break; }
            state = (state << 32) | words[--pos];
        }
        const int s = decode_symbol(state, mean[i], scale[i], kTab, flags);
        out[i] = (float)s * 0.00390625f;
    }
    *end_state = state;
    if (state != kRansL || pos != 0) flags |= ST_BAD_END_STATE;
    return flags;
}

// Markstein division by a float-valued divisor vs IEEE division.  Operands are drawn the way
// the coder produces them: a = (k/256 + 1/512) - mean, b = (double)(float scale).
static inline uint64_t splitmix(uint64_t& s) {
    uint64_t z = (s += 0x9e3779b97f4a7c15ull);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}

int64_t hh_div_check(int64_t n, uint64_t seed, int mode) {
    int64_t bad = 0;
    uint64_t s = seed;
    for (int64_t i = 0; i < n; ++i) {
        uint64_t r1 = splitmix(s), r2 = splitmix(s), r3 = splitmix(s);
        float mean, scale;
        int k;
        if (mode == 0) {  // realistic magnitudes
            mean = (float)((double)(int64_t)(r1 % 2000001) / 1000000.0 - 1.0) * 4.0f;
            uint32_t eb = 90 + (uint32_t)(r2 % 60);  // scale in [2^-37, 2^22]
            uint32_t bits = (eb << 23) | (uint32_t)(r2 >> 41);
            memcpy(&scale, &bits, 4);
            k = (int)(r3 % 4097) - 2048 + (int)(mean * 256.0f);
        } else {  // any finite positive float scale, any float mean of moderate size
            uint32_t mb = (uint32_t)r1;
            memcpy(&mean, &mb, 4);
            if (!(fabsf(mean) <= 16384.0f)) mean = 0.25f;
            uint32_t bits = (uint32_t)(r2 >> 33);
            if (bits == 0 || bits >= 0x7f800000u) bits = 0x3f800001u;
            memcpy(&scale, &bits, 4);
            k = (int)(r3 % 8388608) - 4194304;
        }
        SymbolModel m = make_model(mean, scale);
        const double a = ((double)k * 0.00390625 + 0.001953125) - (double)mean;
        const double q = div_by_scale(a, m);
        const double qq = a / (double)scale;
        if (f64_bits(q) != f64_bits(qq)) ++bad;
    }
    return bad;
}

// Integer division by reciprocal as rans_push() does it, over random (state, freq) pairs that
// satisfy the post-renormalisation precondition state < freq << 40.
int64_t hh_push_check(int64_t n, uint64_t seed) {
    int64_t bad = 0;
    uint64_t s = seed;
    for (int64_t i = 0; i < n; ++i) {
        uint64_t r1 = splitmix(s), r2 = splitmix(s), r3 = splitmix(s);
        uint32_t freq = (uint32_t)(r1 % 16777215) + 1;
        if ((r3 & 7) == 0) freq = (uint32_t)(r3 >> 8) % 64 + 1;
        uint64_t hi = (uint64_t)freq << 40;
        uint64_t state = kRansL + r2 % (hi - kRansL);
        if ((r3 & 0x30) == 0) state = hi - 1 - (r2 & 0xffff);
        uint32_t start = (uint32_t)(r3 >> 40) & 0xffffff;
        uint64_t st = state;
        uint32_t w;
        bool emit = rans_push(st, start, freq, w);
        uint64_t want = ((state / freq) << 24) + (state % freq) + start;
        if (emit || st != want) ++bad;
    }
    return bad;
}

}  // extern "C"
