// flic_core.cuh -- bit-exact arithmetic of the reference coder, written once for the device.
//
// Everything the kernels need to reproduce rans/rans.pyx on sm_100a lives here as
// __host__ __device__ inline functions:
//   * expf_glibc()    glibc >= 2.27 expf algorithm (sysdeps/ieee754/flt-32/e_expf.c), which is
//                     what the reference links (expf@GLIBC_2.27, rans/rans.cpp:1301)
//   * lower_of()      window origin                       rans/rans.pyx:51,92  (rans.cpp:1683,2145)
//   * SymbolModel     per-symbol constants shared by every CDF evaluation of that symbol
//   * cdf_at()        CDF(s/256) = part1 + part2          rans/rans.pyx:31-35  (rans.cpp:1418-1449)
//   * rans_push()     encoder renorm + state update       rans/rans.pyx:61-66  (rans.cpp:1765-1845)
//   * rans_pop_*()    decoder renorm / state update       rans/rans.pyx:87-90,108
//   * search_symbol() smallest s in the window with CDF(s) > mod   rans/rans.pyx:92-104
// The float/double promotion order follows the reference's generated C++.
// Every IEEE operation whose rounding matters goes through dadd/dsub/dmul/dfma below so that
// neither nvcc nor the host compiler can contract or reassociate it.
//
// The same header compiles with plain g++ (tests/host_harness.cpp) so that the search / rANS
// logic is checked against the oracle on the CPU before any GPU time is spent.  That harness is
// test infrastructure; the product never runs these functions on the host.
#pragma once
#include <stdint.h>
#include <math.h>
#include <string.h>

#if defined(__CUDACC__)
#define FLIC_HD __host__ __device__ __forceinline__
#else
#define FLIC_HD static inline
#endif

namespace flic {

// ---- status bits (per stream), a superset of the reference's exceptions ----------------------
enum : int32_t {
    ST_OK = 0,
    ST_ZERO_SCALE = 1,     // reference: ZeroDivisionError "float division" (rans.cpp:1435-1437)
    ST_OUT_OF_WINDOW = 2,  // reference: silent corruption; flagged here
    ST_UNDERRUN = 4,       // reference: unchecked buffer[pos] (rans.cpp:2109); flagged here
    ST_NONFINITE = 8,      // NaN/inf scale, or |mean| > 16384 (window arithmetic no longer exact)
    ST_BAD_END_STATE = 16, // decoder did not return to 1<<32 (rans/test.py:26 prints it)
    ST_NO_SYMBOL = 32,     // decoder: mod >= CDF(upper); reference would emit lower+2048
    ST_TOO_LONG = 64,      // decoder: a single stream of 2^32 words or more is not supported
};

constexpr uint64_t kRansL = 0x100000000ull;  // rans.pyx:13
constexpr uint32_t kProbMask = 0xffffffu;    // rans.pyx:90
constexpr int kWindow = 2048;                // rans.pyx:22
constexpr double kPart1Scale = 16775168.0;   // M - 2048, rans.pyx:34

// ---- IEEE primitives that must not be contracted ----------------------------------------------
FLIC_HD double dadd(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dadd_rn(a, b);
#else
    return a + b;
#endif
}
FLIC_HD double dsub(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dsub_rn(a, b);
#else
    return a - b;
#endif
}
FLIC_HD double dmul(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dmul_rn(a, b);
#else
    return a * b;
#endif
}
FLIC_HD double dfma(double a, double b, double c) {
#if defined(__CUDA_ARCH__)
    return __fma_rn(a, b, c);
#else
    return fma(a, b, c);
#endif
}
// Reciprocals.  The hardware seed (MUFU.RCP64H) looks at the high word of its operand only, so its
// relative error e0 is up to ~2^-20; every operand this file feeds it is a normal double with an
// exponent far from the extremes -- a float widened to double (2^-149 .. 2^128), 1 + e with
// e in [0, 2^185], or an integer frequency in [1, 2^24] -- so no range test is needed.
//
// rcp_cubic: seed, then y (1 + e + e^2) with e = 1 - a y: relative error ~ e0^3 + 2^-53 < 2^-52.
// Enough wherever the quotient built from it is corrected with an exact remainder (div_by_scale,
// rans_push).
FLIC_HD double rcp_cubic(double a) {
#if defined(__CUDA_ARCH__)
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
    double e = __fma_rn(-a, y, 1.0);
    e = __fma_rn(e, e, e);
    return __fma_rn(y, e, y);
#else
    return 1.0 / a;
#endif
}
// drcp: stands in for the reference's correctly rounded `1.0 / a` in part1_from_arg().  It is the
// same cubic step, whose result is RN of a value within ~2^-58 (relative) of 1/a -- the correctly
// rounded reciprocal except when 1/a lies that close to a rounding boundary, and then one ulp
// off.  One ulp of p moves p * 16775168 by at most one ulp, which changes part1 only if that
// product sits on a float rounding boundary to within 2^-52.  part1_from_arg() is a function of
// one 32-bit float, and the exhaustive sweep of all 2^32 of them against the reference
// arithmetic (tests/test_gpu_flow.py::test_device_part1_exhaustive_over_every_float_argument)
// shows that this never happens: the shorter sequence is exact for every input that exists.
#ifndef FLIC_DRCP_STEPS
#define FLIC_DRCP_STEPS 3
#endif
FLIC_HD double drcp(double a) {
#if defined(__CUDA_ARCH__)
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
    double e = __fma_rn(-a, y, 1.0);
#if FLIC_DRCP_STEPS == 3
    e = __fma_rn(e, e, e);
    return __fma_rn(y, e, y);
#else
    y = __fma_rn(y, e, y);
    e = __fma_rn(-a, y, 1.0);
    return __fma_rn(y, e, y);
#endif
#else
    return 1.0 / a;
#endif
}
FLIC_HD float ffma(float a, float b, float c) {
#if defined(__CUDA_ARCH__)
    return __fmaf_rn(a, b, c);
#else
    return fmaf(a, b, c);
#endif
}
// (int) roundf(f) for f >= 0: floor(f + 0.5) computed with a round-toward-zero add, which can
// never carry across an integer the way a round-to-nearest add can (0.49999997f + 0.5f).
FLIC_HD int round_half_away_nonneg(float f) {
#if defined(__CUDA_ARCH__)
    return __float2int_rz(__fadd_rz(f, 0.5f));
#else
    return (int)floor((double)f + 0.5);
#endif
}
FLIC_HD float d2f(double a) {
#if defined(__CUDA_ARCH__)
    return __double2float_rn(a);
#else
    return (float)a;
#endif
}
FLIC_HD uint64_t f64_bits(double d) {
#if defined(__CUDA_ARCH__)
    return (uint64_t)__double_as_longlong(d);
#else
    uint64_t u; memcpy(&u, &d, 8); return u;
#endif
}
FLIC_HD double bits_f64(uint64_t u) {
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)u);
#else
    double d; memcpy(&d, &u, 8); return d;
#endif
}
// high / low 32-bit words of a double, and back: register renaming on the device, so integer
// work on the exponent field costs 32-bit instructions without carry chains
FLIC_HD uint32_t f64_hi(double d) {
#if defined(__CUDA_ARCH__)
    return (uint32_t)__double2hiint(d);
#else
    return (uint32_t)(f64_bits(d) >> 32);
#endif
}
FLIC_HD uint32_t f64_lo(double d) {
#if defined(__CUDA_ARCH__)
    return (uint32_t)__double2loint(d);
#else
    return (uint32_t)f64_bits(d);
#endif
}
FLIC_HD double f64_from_words(uint32_t hi, uint32_t lo) {
#if defined(__CUDA_ARCH__)
    return __hiloint2double((int)hi, (int)lo);
#else
    return bits_f64(((uint64_t)hi << 32) | lo);
#endif
}
FLIC_HD uint32_t f32_bits(float f) {
#if defined(__CUDA_ARCH__)
    return __float_as_uint(f);
#else
    uint32_t u; memcpy(&u, &f, 4); return u;
#endif
}

// float -> int, truncating, saturating, NaN -> 0 (the device conversion's semantics)
FLIC_HD int f2i_rz(float v) {
#if defined(__CUDA_ARCH__)
    return __float2int_rz(v);
#else
    if (v != v) return 0;
    if (v >= 2147483648.0f) return 2147483647;
    if (v <= -2147483648.0f) return (-2147483647 - 1);
    return (int)v;
#endif
}
FLIC_HD int d2i_rz(double v) {
#if defined(__CUDA_ARCH__)
    return __double2int_rz(v);
#else
    if (v != v) return 0;
    if (v >= 2147483648.0) return 2147483647;
    if (v <= -2147483648.0) return (-2147483647 - 1);
    return (int)v;
#endif
}

// ---- conversions done on the FP64 / integer pipes -------------------------------------------------
// The conversion instructions (F2F, F2I, I2F) issue to the quarter-rate XU pipe on sm_100 -- one
// costs as much pipe time as four FP64 operations -- and the coder needs about twenty of them
// per symbol when written naively.  The helpers below get the same values from exact FP64
// additions with "magic" constants and from integer operations on the bit patterns.  They are
// portable C++ (the host harness runs the same text), except the truncating add.

// v rounded to a 24-bit significand, ties to even: (double)(float)v for every v whose float image
// is normal.  M = 1.5 * 2^(e+29), e = exponent of v, puts the sum in the binade whose ulp is a
// float's ulp in v's binade (2^(e-23)); the FP64 adder does the rounding and the subtraction is
// exact.  M's significand is even in those ulps, so ties break exactly as the conversion's do.
// Differences from the conversion pair: no overflow to inf and no subnormal range.  Callers show
// that both only occur where part1 is saturated anyway (see part1_at).
// M's low word only has to be even in units of its last place for the ties to break as the
// conversion's do, and small enough to leave v + M in M's binade; taking the exponent field itself
// (bits 20..30 only: even, and below 2^-21 relative) saves the move that would zero a register.
FLIC_HD double round24(double v) {
    const uint32_t ex = f64_hi(v) & 0x7ff00000u;
    const double M = f64_from_words(ex + 0x01d80000u, ex);
    return dsub(dadd(v, M), M);
}

// floor(w) for 0 <= w < 2^32 as an integer: w + 2^52 rounded toward zero leaves floor(w) in the
// low word of the sum.
FLIC_HD uint32_t floor_nonneg_u32(double w) {
#if defined(__CUDA_ARCH__)
    return (uint32_t)__double2loint(__dadd_rz(w, 4503599627370496.0));
#else
    return (uint32_t)(uint64_t)floor(w);
#endif
}

// (double)n for |n| < 2^31, exact: the bits of 2^52 + 2^51 + n are 0x4338000000000000 + n.
FLIC_HD double i32_to_f64(int n) {
    return dsub(bits_f64(0x4338000000000000ull + (uint64_t)(int64_t)n), 6755399441055744.0);
}
// (double)n for n < 2^32, exact.
FLIC_HD double u32_to_f64(uint32_t n) {
    return dsub(bits_f64(0x4330000000000000ull | (uint64_t)n), 4503599627370496.0);
}
// (double)(hi * 2^32 + lo), correctly rounded (one rounding, in the last add):
// bits 0x45300000:hi = 2^84 + hi 2^32, bits 0x43300000:lo = 2^52 + lo.
FLIC_HD double u64_to_f64(uint32_t hi, uint32_t lo) {
    const double dh = dsub(bits_f64(0x4530000000000000ull | (uint64_t)hi), 19342813118337666422669312.0);  // 2^84 + 2^52
    return dadd(dh, bits_f64(0x4330000000000000ull | (uint64_t)lo));
}

// q with its magnitude limited to [0, 128 (1 + 2^-20)): beyond |q| = 128 the caller's result is
// saturated and only the sign matters.  Integer min on the high word; NaN becomes finite too.
FLIC_HD double clamp_mag128(double q) {
    const uint32_t hi = f64_hi(q);
    uint32_t ah = hi & 0x7fffffffu;
    ah = ah < 0x40600000u ? ah : 0x40600000u;
    return f64_from_words(ah | (hi & 0x80000000u), f64_lo(q));
}

// ---- glibc expf ---------------------------------------------------------------------------------
// T[i] = bits(2^(i/32)) - (i << 47).  Published table of glibc's __exp2f_data (N = 32).
#define FLIC_EXP2F_TABLE                                                                          \
    0x3ff0000000000000ull, 0x3fefd9b0d3158574ull, 0x3fefb5586cf9890full, 0x3fef9301d0125b51ull,   \
    0x3fef72b83c7d517bull, 0x3fef54873168b9aaull, 0x3fef387a6e756238ull, 0x3fef1e9df51fdee1ull,   \
    0x3fef06fe0a31b715ull, 0x3feef1a7373aa9cbull, 0x3feedea64c123422ull, 0x3feece086061892dull,   \
    0x3feebfdad5362a27ull, 0x3feeb42b569d4f82ull, 0x3feeab07dd485429ull, 0x3feea47eb03a5585ull,   \
    0x3feea09e667f3bcdull, 0x3fee9f75e8ec5f74ull, 0x3feea11473eb0187ull, 0x3feea589994cce13ull,   \
    0x3feeace5422aa0dbull, 0x3feeb737b0cdc5e5ull, 0x3feec49182a3f090ull, 0x3feed503b23e255dull,   \
    0x3feee89f995ad3adull, 0x3feeff76f2fb5e47ull, 0x3fef199bdd85529cull, 0x3fef3720dcef9069ull,   \
    0x3fef5818dcfba487ull, 0x3fef7c97337b9b5full, 0x3fefa4afa2a490daull, 0x3fefd0765b6e4540ull

// `tab` names the 32-entry table: on the device the shared-space byte address of a 256-byte
// aligned copy (each lane indexes its own entry, so constant memory would serialise; a 32-bit
// shared address costs one LOP3 per lookup where a generic pointer costs the shared-window base
// -- S2R + LEA -- per use), a static array on the host.
#if defined(__CUDACC__)
typedef uint32_t ExpTab;
#else
typedef const uint64_t* ExpTab;
#endif
FLIC_HD uint64_t exp_tab_entry(ExpTab tab, uint32_t ki) {
#if defined(__CUDA_ARCH__)
    uint64_t v;
    asm("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(tab | ((ki << 3) & 0xf8u)));
    return v;
#elif defined(__CUDACC__)
    (void)tab; (void)ki;
    return 0;   // nvcc's host pass only; the product never evaluates the coder on the host
#else
    return tab[ki & 31];
#endif
}
//
// exp_core(xd) is the double-precision body of glibc's expf for xd = (double)x: it returns the
// product y * s whose conversion to float is the function's result.  Valid (no exponent
// overflow in s) for |xd| < 700.
// The polynomial is evaluated with fused multiply-adds in Horner form (see below); the GPU suite
// sweeps every float |x| <= 104 against the host libm: equal everywhere except two inputs deep
// inside part1's saturated range, where the host itself departs from the published algorithm.
// `neg` evaluates exp(-xd): (-InvLn2N) * xd is bit-identical to InvLn2N * (-xd).
#if defined(__CUDACC__)
// InvLn2N, C0, C1, C2 of exp_core(): their low words are not zero, so they cannot be immediates;
// from constant memory two of them arrive per uniform load instead of costing two moves each
// every time register pressure makes the compiler rebuild them.
static __constant__ double c_expk[4] = {0x1.71547652b82fep+0 * 32.0, 0x1.c6af84b912394p-5 / 32.0 / 32.0 / 32.0,
                                        0x1.ebfce50fac4f3p-3 / 32.0 / 32.0, 0x1.62e42ff0c52d6p-1 / 32.0};
#endif
FLIC_HD double exp_core(double xd, ExpTab tab, bool neg) {
#if defined(__CUDA_ARCH__)
    const double InvLn2N = c_expk[0], C0 = c_expk[1], C1 = c_expk[2], C2 = c_expk[3];
#else
    const double InvLn2N = 0x1.71547652b82fep+0 * 32.0;
    const double C0 = 0x1.c6af84b912394p-5 / 32.0 / 32.0 / 32.0;
    const double C1 = 0x1.ebfce50fac4f3p-3 / 32.0 / 32.0;
    const double C2 = 0x1.62e42ff0c52d6p-1 / 32.0;
#endif
    const double Shift = 0x1.8p+52;
    const double z = dmul(neg ? -InvLn2N : InvLn2N, xd);
    double kd = dadd(z, Shift);
    const uint32_t ki = f64_lo(kd);          // k mod 2^32 sits in the low mantissa bits
    kd = dsub(kd, Shift);
    const double r = dsub(z, kd);
    // T[k % 32] + (k << 47): the shift only reaches the high word (47 - 32 = 15)
    const uint64_t tv = exp_tab_entry(tab, ki);
    const double s = f64_from_words((uint32_t)(tv >> 32) + (ki << 15), (uint32_t)tv);
    // glibc evaluates (C0 r + C1) r^2 + (C2 r + 1); Horner's form has one operation less and
    // differs from it by an ulp or two of the double -- which could change the float result only
    // if y s sat within ~2^-51 of a float rounding boundary.  Both exhaustive sweeps (every float
    // |x| <= 104 against the host libm, and part1 over all 2^32 arguments) pass with this form, so
    // it is exact for every input that exists.
    double y = dfma(C0, r, C1);
    y = dfma(y, r, C2);
    y = dfma(y, r, 1.0);
    return dmul(y, s);
}

// glibc's expf, branch-free; equal to it for every non-NaN float (tests sweep all of them).
// glibc special-cases |x| >= 88: x > 0x1.62e42ep6 -> +inf, x < -0x1.9fe368p6 -> 0, inf/nan
// passthrough.  Clamping x to [-104, 89] and running the main path gives the same floats:
// e^89 overflows the final double->float conversion to +inf, e^-104 rounds to 0 (it is below
// half the smallest subnormal, the very definition of glibc's underflow threshold), and the
// results in between are the main path's anyway.  NaN is mapped to the lower clamp.
// The coder itself uses exp_core() directly (part1_at); this wrapper exists for the sweeps.
FLIC_HD float expf_glibc(float x, ExpTab tab) {  // x by value: clamped below
    x = fminf(fmaxf(x, -104.0f), 89.0f);
    return d2f(exp_core((double)x, tab, false));
}

// ---- per-symbol model -----------------------------------------------------------------------------
FLIC_HD int lower_of_d(double mean_d) {
    // (int) round(mean_d * 256.0 - 1024.0), C round(): half away from zero.
    // v is exact (24-bit significand times 2^8, minus 2^10), so one fma computes it; |v| + 0.5
    // is exact too, and half away from zero is sign(v) * floor(|v| + 0.5).
    const double v = dfma(mean_d, 256.0, -1024.0);
    const int n = (int)floor_nonneg_u32(dadd(fabs(v), 0.5));
    return (int64_t)f64_bits(v) < 0 ? -n : n;
}
// The same integer on the float pipe.  m = 256 mean is exact in float (a power-of-two scaling;
// |mean| <= 16384 keeps it below 2^22), and with v = m - 1024:
//   v >= 0:  round(v) = floor(m + 0.5) - 1024
//   v <  0:  round(v) = ceil(m - 0.5) - 1024 = -floor(-m + 0.5) - 1024        (half away from zero)
// floor(w + 0.5) is taken from a round-DOWN add: RD(w + 0.5) lies between floor(w + 0.5) -- an
// integer below 2^23, hence a float, and not above the exact sum -- and the exact sum, so its floor
// is the exact sum's.  Three FP64 operations and two constant moves less per symbol than lower_of_d.
FLIC_HD int lower_of(float mean) {
    const float m = mean * 256.0f;
    const bool neg = m < 1024.0f;
    const float w = neg ? -m : m;
#if defined(__CUDA_ARCH__)
    const int t = __float2int_rd(__fadd_rd(w, 0.5f));
#else
    const double td = floor((double)w + 0.5);
    const int t = td >= 2147483648.0 ? 2147483647 : (td <= -2147483648.0 ? (-2147483647 - 1) : (int)td);   // NaN / inf: any value, the stream is flagged
#endif
    return (neg ? -t : t) - 1024;
}

struct SymbolModel {
    double mean_d;   // (double)mean
    double scale_d;  // (double)scale
    double rscale;   // 1 / scale_d to within 2^-52 relative
    int lower;       // window origin in 1/256 units
};

// Parameters the arithmetic below is exact for: finite non-zero scale, and |mean| limited so that
// every window index stays below 2^23 and the float arithmetic of rans.pyx:33 (x - lower) is
// exact (8-bit image latents are within a few units of zero).  NaN fails every comparison.
FLIC_HD bool params_ok(float mean, float scale) {
    return (fabsf(mean) <= 16384.0f) && (scale < INFINITY) && (scale > 0.0f);
}
// Status bits for parameters that are not ok (slow path only).  A negative scale makes part1
// decrease in s (freq <= 0: the reference divides by zero or corrupts the stream); it is
// reported with the non-finite ones.
FLIC_HD int32_t param_flags(float mean, float scale) {
    int32_t f = 0;
    if (scale == 0.0f) f |= ST_ZERO_SCALE;
    if (!(fabsf(mean) <= 16384.0f) || !(scale < INFINITY) || scale < 0.0f) f |= ST_NONFINITE;
    return f;
}

// The same test for a whole stream at three integer instructions per symbol: running unsigned
// maxima of bits(scale) - 1 (zero wraps to 0xffffffff, negative values and NaN / inf are
// >= 0x7f7fffff) and of bits(|mean|) (16384.0f is 0x46800000; NaN is above it).  The hot loops
// note every symbol and turn the maxima into status bits once per stream.
struct ParamGuard {
    uint32_t smax, mmax;
};
FLIC_HD ParamGuard guard_init() { ParamGuard g; g.smax = 0u; g.mmax = 0u; return g; }
FLIC_HD void guard_note(ParamGuard& g, float mean, float scale) {
    const uint32_t sb = f32_bits(scale) - 1u, mb = f32_bits(mean) & 0x7fffffffu;
    g.smax = sb > g.smax ? sb : g.smax;
    g.mmax = mb > g.mmax ? mb : g.mmax;
}
FLIC_HD int32_t guard_flags(const ParamGuard& g) {
    int32_t f = 0;
    if (g.smax >= 0x7f7fffffu) f |= (g.smax == 0xffffffffu || g.smax == 0x7fffffffu) ? ST_ZERO_SCALE : ST_NONFINITE;
    if (g.mmax > 0x46800000u) f |= ST_NONFINITE;
    return f;
}

// The model is computed unconditionally; for parameters that are not ok its contents are
// meaningless but harmless (no trap, no unbounded loop) and the stream is flagged by the caller.
FLIC_HD SymbolModel make_model(float mean, float scale) {
    SymbolModel m;
    m.mean_d = (double)mean;
    m.scale_d = (double)scale;
    m.rscale = rcp_cubic(m.scale_d);
    m.lower = lower_of(mean);
    return m;
}

// Correctly rounded a / b from r ~ 1/b (relative error eps < 2^-52): q0 = RN(a r);
// rem = a - b q0 (exact, fma); q = RN(q0 + rem r).  The value rounded in the last step is
// Q + (Q - q0) eps' with Q = a / b, |Q - q0| <= 2^-51 |Q| and |eps'| < 2^-52, i.e. within 2^-103
// (relative) of Q.  b is a float widened to double (24-bit significand), and a quotient of a
// 53-bit by a 24-bit significand is either exactly representable (b a power of two) or at least
// 2^-78 (relative) away from every double and every midpoint between doubles (|a 2^k - m b| >= 1
// for integers), so no rounding boundary can fall in between: q = RN(a / b).  No intermediate can
// over/underflow for the operand ranges of this file.  tests/ compares it with IEEE division on
// the host (millions of cases) and every table / bitstream parity test exercises it on the device.
FLIC_HD double div_by_scale(double a, const SymbolModel& m) {
    const double q0 = dmul(a, m.rscale);
    const double e = dfma(-q0, m.scale_d, a);
    return dfma(e, m.rscale, q0);
}

// part1 of CDF at the point whose (xq + 1/512) is `a`:
//   (int) roundf( (float)( 1/(1+expf(-arg)) * 16775168 ) ),  arg = (float)( (a - mean) / scale ).
// The three float roundings are round24() on doubles, never leaving the FP64 pipe.  Where that
// differs from a real conversion the result is the same:
//   * arg: |q| >= 2^128 would convert to inf and glibc's expf special-cases |x| >= 88; here q is
//     limited to |q| < 128.001 first.  For arg in (88.7, 128] the reference has e = inf, p = 0,
//     part1 = 0, and here e <= e^128.001 is a finite double, p < 2^-184, part1 = 0; for arg in
//     [-128, -103.9) the reference has e = 0 and here e < 2^-149, 1 + e == 1 both ways.
//     |q| < 2^-126 (float subnormal) keeps more bits here, but expf(-arg) == 1.0f for all
//     |arg| < 2^-25.
//   * e: a float-subnormal e (< 2^-126) keeps more bits here, but 1 + e == 1 below 2^-53.
//   * the product is at most 16775168 and at least 2^-184 * A > 0: no range issue; below 2^-126
//     both round to 0.
// Which of the three float roundings use the conversion instructions (XU pipe) and which the
// FP64-pipe round24().  All eight combinations give identical tables (tools/variant_bench.cu);
// measured on B200, K1 runs 152 / 161 / 169 / 163 G symbols/s for 000 / 001 / 101 / 111: the
// XU pipe (8 cycles per warp-instruction) and the FP64 pipe (2 cycles) overlap, so the work is
// split between them.
#ifndef FLIC_ARG_XU
#define FLIC_ARG_XU 1
#endif
#ifndef FLIC_E_XU
#define FLIC_E_XU 0
#endif
#ifndef FLIC_V_XU
#define FLIC_V_XU 1
#endif
// Everything after the argument: arg is (double)(float)arg_f with |arg| <= 128 (or <= 128.001 from
// the FP64-pipe rounding).  A pure function of a 32-bit float, so it is swept exhaustively
// against the reference arithmetic on the GPU (tests/test_gpu_flow.py, flic_debug_part1).
FLIC_HD int part1_from_arg(double arg, ExpTab tab) {
#if FLIC_E_XU
    const double e = (double)fminf(d2f(exp_core(arg, tab, true)), 3.402823466e+38f);
#else
    const double e = round24(exp_core(arg, tab, true));
#endif
    const double p = drcp(dadd(1.0, e));
#if FLIC_V_XU
    return round_half_away_nonneg(d2f(dmul(p, kPart1Scale)));
#else
    const double v = round24(dmul(p, kPart1Scale));
    return (int)floor_nonneg_u32(dadd(v, 0.5));  // roundf of a non-negative float: floor(v + 0.5), exact sum
#endif
}

// The quotient q = t4 / scale rounded to float and limited to [-128, 128].
FLIC_HD double arg_from_quotient(double q) {
#if FLIC_ARG_XU
    return (double)fminf(fmaxf(d2f(q), -128.0f), 128.0f);
#else
    return round24(clamp_mag128(q));
#endif
}

FLIC_HD int part1_at(double a, const SymbolModel& m, ExpTab tab) {
    const double t4 = dsub(a, m.mean_d);
    return part1_from_arg(arg_from_quotient(div_by_scale(t4, m)), tab);
}

// (double)xq + 1/512 = (2 s + 1) / 512 exactly, built without a conversion: the bits of
// 1.5 * 2^43 + (2^31 + n) / 512 are 0x42a80000:00000000 + 2^31 + n (ulp 2^-9 in that binade).  The
// offset 2^31 keeps the low word from borrowing for negative n (|n| < 2^25), so the high word is a
// constant and the low word one three-input add.
FLIC_HD double half_bin_point(int s) {
    const uint32_t lo = (uint32_t)s + (uint32_t)s + 0x80000001u;
    return dsub(f64_from_words(0x42a80000u, lo), 13194143727616.0);   // 1.5 * 2^43 + 2^22
}

// CDF(s/256) for integer symbol s: part1 + part2, part2 = round((xq - lower_f) * 256) + 1
// = s - lower + 1 exactly when |s| and |lower| are below 2^24 (both floats exact, difference
// exact).
FLIC_HD int cdf_at(int s, const SymbolModel& m, ExpTab tab) {
    return part1_at(half_bin_point(s), m, tab) + (s - m.lower + 1);
}

// CDF(s-1) and CDF(s) together: the (start, end) pair of symbol s (rans.pyx:52-53, :106-107).
// The two evaluations are independent and interleave in the instruction stream.
FLIC_HD void cdf_pair(int s, const SymbolModel& m, ExpTab tab, int& c_lo, int& c_hi) {
    const double a_hi = half_bin_point(s);
    const double a_lo = dsub(a_hi, 0.00390625);  // exact
    c_hi = part1_at(a_hi, m, tab) + (s - m.lower + 1);
    c_lo = part1_at(a_lo, m, tab) + (s - m.lower);
}

// The same pair for the decoder's first try, without the limit of the argument to [-128, 128]:
// returns whether both float arguments lie in [-680, 680].  Up to there the unlimited evaluation
// gives what the limited one gives: inside [-128, 128] it is the same computation, and beyond it
// both are saturated -- arg > 128: e <= e^-128, 1 + e == 1, part1 = A either way; arg < -128:
// e >= e^128 (at most e^680 < 2^982: every double on the way is normal, neither the table's
// exponent add nor round24()'s constant 2^(ex + 29) overflows), p A < 2^-160 rounds to the float
// 0, part1 = 0 either way.  This keeps
// distributions much narrower than a bin (|a - mean| / scale beyond 128 at the bin's own edges,
// scale < 1.5e-5) on the fast path.  Outside [-680, 680] (scale below 5.7e-6, an overflowed
// quotient) the values are meaningless but harmless -- nothing traps, the exp table index is
// masked -- and the caller's search, which evaluates with the limit, starts from the guess without
// them.  Two min/max per evaluation become one max and one compare per pair.
#ifndef FLIC_TRY_UNLIMITED
#define FLIC_TRY_UNLIMITED 1
#endif
FLIC_HD bool cdf_pair_try(int s, const SymbolModel& m, ExpTab tab, int& c_lo, int& c_hi) {
#if FLIC_ARG_XU && FLIC_TRY_UNLIMITED
    const double a_hi = half_bin_point(s);
    const double a_lo = dsub(a_hi, 0.00390625);  // exact
    const float f_hi = d2f(div_by_scale(dsub(a_hi, m.mean_d), m));
    const float f_lo = d2f(div_by_scale(dsub(a_lo, m.mean_d), m));
    c_hi = part1_from_arg((double)f_hi, tab) + (s - m.lower + 1);
    c_lo = part1_from_arg((double)f_lo, tab) + (s - m.lower);
    return fmaxf(fabsf(f_hi), fabsf(f_lo)) <= 680.0f;
#else
    cdf_pair(s, m, tab, c_lo, c_hi);
    return true;
#endif
}

struct SymbolTable {
    uint32_t start;  // CDF(x - 1/256)
    uint32_t freq;   // CDF(x) - start  (>= 1)
};

// encode pass 1 for one symbol (rans.pyx:50-56).  flags accumulates status bits.
// A symbol is codable when its parameters are ok, x is an exact multiple of 1/256 and its grid
// index lies in [lower, lower + 2047]; then every float operation of rans.pyx:33 is exact and
// part2 is plain integer arithmetic.  Anything else makes the reference silently emit an
// undecodable stream; here the stream is flagged and a harmless entry keeps
// the coder alive.
FLIC_HD SymbolTable make_table(float x, float mean, float scale, ExpTab tab, int32_t& flags) {
    const SymbolModel m = make_model(mean, scale);
    const float xs = x * 256.0f;
    const int s = f2i_rz(xs);
    const bool in_window = ((float)s == xs) && ((uint32_t)(s - m.lower) < (uint32_t)kWindow);
    SymbolTable t;
    if (!(in_window && params_ok(mean, scale))) {
        flags |= param_flags(mean, scale) | (in_window ? 0 : ST_OUT_OF_WINDOW);
        t.start = 0; t.freq = 1;
        return t;
    }
    int c0, c1;
    cdf_pair(s, m, tab, c0, c1);
    t.start = (uint32_t)c0;
    t.freq = (uint32_t)(c1 - c0);
    return t;
}

// make_table() for the hot loops: parameters are noted in `guard` instead of being tested per
// symbol, and the entry is computed unconditionally -- for a symbol or parameters that are not
// codable it is meaningless but harmless (no trap, no unbounded loop, at most one word per
// symbol either way) and the stream is flagged.
FLIC_HD SymbolTable make_table_lean(float x, float mean, float scale, ExpTab tab, ParamGuard& guard, int32_t& flags) {
    const SymbolModel m = make_model(mean, scale);
    guard_note(guard, mean, scale);
    const float xs = x * 256.0f;
    const int s = f2i_rz(xs);
    if (!(((float)s == xs) && ((uint32_t)(s - m.lower) < (uint32_t)kWindow))) flags |= ST_OUT_OF_WINDOW;
    int c0, c1;
    cdf_pair(s, m, tab, c0, c1);
    SymbolTable t;
    t.start = (uint32_t)c0;
    t.freq = (uint32_t)(c1 - c0);
    return t;
}

// ---- rANS state machine ---------------------------------------------------------------------------
// 1/a biased low: (1/a)(1 - 2^-49) to within 2^-52 relative (hardware seed with relative error
// e0 ~ 2^-20, one cubic Newton step whose residual carries the bias; the neglected terms are
// e0^3 and 2^-49 e0).  Feeds the quotient estimate below, which is corrected in integers.
FLIC_HD double rcp_biased_low(double a) {
#if defined(__CUDA_ARCH__)
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
    double e = __fma_rn(-a, y, 0x1.ffffffffffff8p-1);  // (1 - 2^-49) - a y
    e = __fma_rn(e, e, e);
    return __fma_rn(y, e, y);
#else
    return (1.0 / a) * 0x1.ffffffffffff8p-1;
#endif
}

// Encoder step (rans.pyx:62-65).  Returns true and sets `word` when a 32-bit word is emitted.
// state / freq (64 by 24 bits) through FP64: after renormalisation state < freq << 40, so
// q < 2^40.  qd = state_d * rf with rf biased low by 2^-49 is never above the real quotient and
// less than 2^-8 below it (relative errors: bias 2^-49, state_d 2^-53, rf 2^-52, product 2^-53;
// times q < 2^40), so floor(qd) is q or q - 1 and one integer correction makes it exact.
// The remainder fits 32 bits, so it is computed modulo 2^32.
FLIC_HD double push_reciprocal(uint32_t freq) { return rcp_biased_low(u32_to_f64(freq)); }

// rf = push_reciprocal(freq): it does not depend on the state, so a caller whose table entries are
// produced ahead of the serial recurrence (the producer / consumer encoder) has the producers
// compute it, which leaves convert -> multiply -> fix-up on the chain.
FLIC_HD bool rans_push_rf(uint64_t& state, uint32_t start, uint32_t freq, double rf, uint32_t& word) {
    uint32_t hi = (uint32_t)(state >> 32), lo = (uint32_t)state;
    // state >= freq << 40: the threshold's low 32 bits are zero, so only the high words compare
    const bool emit = hi >= (freq << 8);
    if (emit) { word = lo; lo = hi; hi = 0; }
    const double qd = dmul(u64_to_f64(hi, lo), rf);
#if defined(__CUDA_ARCH__)
    const uint64_t qb = (uint64_t)__double_as_longlong(__dadd_rz(qd, 4503599627370496.0));  // 2^52 + floor(qd)
#else
    const uint64_t qb = (uint64_t)floor(qd);
#endif
    const uint32_t q_lo = (uint32_t)qb;
    uint32_t rem = lo - q_lo * freq;              // state - q freq, in [0, 2 freq)
    const bool up = rem >= freq;                  // q was one short
    // new state = ((q + up) << 24) + (rem - up freq) + start; the exponent bits of qb leave through
    // the top of the shift (q < 2^40)
    const uint64_t add = (uint64_t)(rem + start) + (up ? (uint64_t)(0x1000000u - freq) : 0ull);
    state = (qb << 24) + add;
    return emit;
}
FLIC_HD bool rans_push(uint64_t& state, uint32_t start, uint32_t freq, uint32_t& word) {
    return rans_push_rf(state, start, freq, push_reciprocal(freq), word);
}

// Decoder: state update after the symbol is known (rans.pyx:108).
FLIC_HD void rans_pop(uint64_t& state, uint32_t start, uint32_t freq) {
    state = (state >> 24) * (uint64_t)freq + (state & kProbMask) - (uint64_t)start;
}

// The same on the state's two words: x = state >> 24 is 40 bits (xh:xl), mod - start is in
// [0, freq), and the new state is x freq + (mod - start) < 2^64.
FLIC_HD void rans_pop32(uint32_t& hi, uint32_t& lo, uint32_t start, uint32_t freq) {
    const uint32_t xl = (hi << 8) | (lo >> 24), xh = hi >> 24;
    const uint32_t d = (lo & kProbMask) - start;
    const uint64_t p = (uint64_t)xl * freq + d;
    hi = (uint32_t)(p >> 32) + xh * freq;
    lo = (uint32_t)p;
}

// ---- decoder symbol search ----------------------------------------------------------------------
// First guess for "smallest s with CDF(s) > mod" from the continuous model
//   g(s) = A sigmoid((s + 0.5 - 256 mean) / (256 scale)) + (s - lower + 1),  A = 16775168
// solved for g = mod + 0.5: start at u0 = logit(p0), p0 = (mod - 1024) / A (the sigmoid term alone,
// window centre), then one Newton step on h(u) = A sig(u) + c u + (m - lower - mod),
// u = (s + 0.5 - m) / c, m = 256 mean, c = 256 scale.  At u0 the sigmoid is p0 by construction, so
// the step needs no exponential: A sig(u0) = P and A (1 - sig(u0)) = Q with P = mod - 1024,
// Q = A - P, h(u0) = c u0 + (m - lower - 1024) and h'(u0) = P Q / A + c.  The step is taken in the
// form  s + 0.5 = m + c u1 = m + c u0 (P Q) / (P Q + c A):  the offset m - lower - 1024 -- the
// distance of the mean from the bin grid, at most half a count of the 2^24 -- is left out of h.
// Half a count moves the answer by 0.5 / freq bins, so over a whole window this costs a wrong guess
// (= a bracket search) with probability sum_bins (freq / 2^24) (0.5 / freq) = 6e-5 per symbol
// whatever the distribution, and saves the conversion of `lower` and two additions on all others.
// Three MUFU operations (two lg2, one rcp).  Only speed depends on the guess.
FLIC_HD int guess_symbol(uint32_t mod, float mean, float scale, int lower) {
    (void)lower;
    // P = mod - 1024, Q = A - P.  Neither is kept positive in the tails (mod < 1025 or
    // mod > A + 1023: 1.2e-4 of all slots) and the guess is not limited to the window: whatever comes
    // out there -- a wrapped integer, a NaN turned into an arbitrary index -- fails the caller's test
    // (in-window and CDF(g - 1) <= mod < CDF(g)) and the bracket search takes over.
    const float Pf = (float)(mod - 1024u);
    const float Qf = 16775168.0f - Pf;                 // exact: integers below 2^24 (mod >= 1024)
    const float cl = scale * 177.445678223f;           // 256 ln 2 scale: c u0 = cl (lg2 P - lg2 Q)
    const float cA = scale * 4294443008.0f;            // 256 A scale (A 2^8 is a float)
    const float mh = mean * 256.0f - 0.5f;             // the product is exact and shared with lower_of()
    const float PQ = Pf * Qf;
#if defined(__CUDA_ARCH__)
    float lp, lq, rd;   // flush-to-zero forms: no subnormal fix-ups (P, Q are integers)
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lp) : "f"(Pf));
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lq) : "f"(Qf));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rd) : "f"(PQ + cA));
    const float sr = ffma((lp - lq) * cl, PQ * rd, mh);
    // ceil through a round-up add of 1.5 * 2^23 (exact integer in the low mantissa bits when
    // |sr| < 2^22; anything else lands outside the window)
    return (int)(__float_as_uint(__fadd_ru(sr, 12582912.0f)) - 0x4b400000u);
#else
    const float sr = ceilf(ffma((log2f(Pf) - log2f(Qf)) * cl, PQ * (1.0f / (PQ + cA)), mh));
    return fabsf(sr) < 4.0e6f ? (int)sr : lower - 1;   // NaN too: outside the window
#endif
}

// A guess is usable when it lies in the window; the callers test that together with the CDF pair.
FLIC_HD bool guess_in_window(int g, int lower) { return (uint32_t)(g - lower) < (uint32_t)kWindow; }
FLIC_HD int clamp_to_window(int g, int lower) {
    g = g < lower ? lower : g;
    return g > lower + (kWindow - 1) ? lower + (kWindow - 1) : g;
}

// Bracketing search.  Invariant: CDF(lo) <= mod < CDF(hi), where lo = lower-1 and
// hi = lower+2048 start as *virtual* ends exactly as in the reference's binary search
// (rans.pyx:96-104 never evaluates outside the window, and treats the window's left edge as
// "not greater").  Finishes when hi == lo + 1 with both CDF values evaluated, which are then
// the (start, end) the reference recomputes at rans.pyx:106-107.  Any probe order returns the
// reference's answer because CDF is non-decreasing in s.
struct SearchState {
    int lo, hi;        // bracket
    int c_lo, c_hi;    // CDF values, -1 = not evaluated yet
    int probe;         // next point to evaluate
    int step;          // gallop step
    bool done;
};

FLIC_HD SearchState search_begin(uint32_t mod, float mean, float scale, const SymbolModel& m) {
    SearchState st;
    st.lo = m.lower - 1;
    st.hi = m.lower + kWindow;
    st.c_lo = -1;
    st.c_hi = -1;
    st.step = 1;
    st.done = false;
    st.probe = clamp_to_window(guess_symbol(mod, mean, scale, m.lower), m.lower);
    return st;
}

// Feed CDF(st.probe) = c; choose the next probe or finish.
FLIC_HD void search_feed(SearchState& st, int c, uint32_t mod) {
    const bool greater = c > (int)mod;  // mod < 2^24; c < 0 (impossible in-window) reads as "not greater"
    if (st.probe == st.lo) { st.c_lo = c; }
    else if (st.probe == st.hi) { st.c_hi = c; }
    else if (greater) { st.hi = st.probe; st.c_hi = c; }
    else { st.lo = st.probe; st.c_lo = c; }

    if (st.hi - st.lo == 1) {
        if (st.c_hi < 0 && st.c_lo < 0) { st.probe = st.hi; return; }
        if (st.c_lo < 0) { st.probe = st.lo; return; }
        if (st.c_hi < 0) { st.probe = st.hi; return; }
        st.done = true;
        return;
    }
    if (st.c_lo >= 0 && st.c_hi >= 0) {  // both ends real: bisect
        st.probe = st.lo + ((st.hi - st.lo) >> 1);
        return;
    }
    // gallop away from the end that is known
    if (st.c_hi >= 0) {  // answer is at or left of hi
        int p = st.hi - st.step;
        st.probe = p > st.lo ? p : st.lo + 1;
    } else {             // answer is right of lo
        int p = st.lo + st.step;
        st.probe = p < st.hi ? p : st.hi - 1;
    }
    st.step <<= 2;
}

// One symbol.  Returns the integer grid index; updates state.
FLIC_HD int decode_symbol(uint64_t& state, float mean, float scale,
                                             ExpTab s_tab, int32_t& flags) {
    const uint32_t mod = (uint32_t)state & kProbMask;
    const SymbolModel m = make_model(mean, scale);
    if (!params_ok(mean, scale)) flags |= param_flags(mean, scale);
    const int g = clamp_to_window(guess_symbol(mod, mean, scale, m.lower), m.lower);
    int c_lo, c_hi;
    cdf_pair(g, m, s_tab, c_lo, c_hi);
    int s = g;
    const bool left_ok = (g == m.lower) || (c_lo <= (int)mod);  // window's left edge is virtual
    if (!(left_ok && c_hi > (int)mod)) {
        // slow path: bracket from what is already known, then gallop / bisect
        SearchState st;
        st.lo = m.lower - 1; st.hi = m.lower + kWindow;
        st.c_lo = -1; st.c_hi = -1; st.step = 2; st.done = false;
        if (c_hi <= (int)mod) {  // answer is right of g
            st.lo = g; st.c_lo = c_hi;
            st.probe = g + 1 < st.hi ? g + 1 : st.hi;
        } else {                 // c_lo > mod and g - 1 >= lower: answer is at or left of g - 1
            st.hi = g - 1; st.c_hi = c_lo;
            st.probe = g - 2 > st.lo ? g - 2 : st.lo;
        }
        while (!st.done) {
            const int c = cdf_at(st.probe, m, s_tab);
            search_feed(st, c, mod);
        }
        s = st.hi; c_hi = st.c_hi; c_lo = st.c_lo;
        if (s > m.lower + (kWindow - 1)) flags |= ST_NO_SYMBOL;
    }
    rans_pop(state, (uint32_t)c_lo, (uint32_t)(c_hi - c_lo));
    return s;
}

// ---- the decoder's hot path --------------------------------------------------------------------
// decode_symbol() split where the probabilities split: the guess is right for all but a few
// symbols per million, so the bracket search is a function of its own (not inlined on the
// device: its registers and code stay out of the unrolled loop), parameters are noted in a
// ParamGuard instead of being tested, and the state is the (hi, lo) word pair the word pull
// works on.
struct SymbolHit {
    int s, c_lo, c_hi;
};

#ifndef FLIC_SEARCH_NOINLINE
#define FLIC_SEARCH_NOINLINE 1
#endif
#if defined(__CUDA_ARCH__) && FLIC_SEARCH_NOINLINE
__device__ __noinline__
#elif defined(__CUDA_ARCH__)
__device__ __forceinline__
#else
static inline
#endif
SymbolHit decode_symbol_search(uint32_t mod, float mean, float scale, ExpTab tab, int g, int c_lo, int c_hi, bool evidence) {
    const SymbolModel m = make_model(mean, scale);
    SearchState st;
    st.lo = m.lower - 1; st.hi = m.lower + kWindow;
    st.c_lo = -1; st.c_hi = -1; st.step = 2; st.done = false;
    if (!(guess_in_window(g, m.lower) && evidence)) {   // nothing is known yet: start at the guess / the nearest window end
        st.probe = clamp_to_window(g, m.lower);
    } else if (g == m.lower && c_hi > (int)mod) {   // the window's left edge is virtual: "not greater"
        SymbolHit h;
        h.s = g; h.c_lo = c_lo; h.c_hi = c_hi;
        return h;
    } else if (c_hi <= (int)mod) {  // answer is right of g
        st.lo = g; st.c_lo = c_hi;
        st.probe = g + 1 < st.hi ? g + 1 : st.hi;
    } else {                 // c_lo > mod and g - 1 >= lower: answer is at or left of g - 1
        st.hi = g - 1; st.c_hi = c_lo;
        st.probe = g - 2 > st.lo ? g - 2 : st.lo;
    }
    while (!st.done) {
        const int c = cdf_at(st.probe, m, tab);
        search_feed(st, c, mod);
    }
    SymbolHit h;
    h.s = st.hi; h.c_lo = st.c_lo; h.c_hi = st.c_hi;
    return h;
}

// The same with the model given (a caller that has it ahead of the chain keeps its ~100 cycles
// of conversions and reciprocal out of the serial dependency).
FLIC_HD int decode_symbol_model(uint32_t& hi, uint32_t& lo, float mean, float scale, const SymbolModel& m,
                                ExpTab tab, ParamGuard& guard, int32_t& flags) {
    const uint32_t mod = lo & kProbMask;
    guard_note(guard, mean, scale);
    const int g = guess_symbol(mod, mean, scale, m.lower);
    int c_lo, c_hi;
    const bool args_ok = cdf_pair_try(g, m, tab, c_lo, c_hi);
    int s = g;
    // (a guess on the window's left edge whose CDF(g - 1) exceeds mod -- only a corrupt stream has
    // that -- is sorted out by the search function: the edge is virtual)
    if (!(c_lo <= (int)mod && c_hi > (int)mod && guess_in_window(g, m.lower) && args_ok)) {
        const SymbolHit h = decode_symbol_search(mod, mean, scale, tab, g, c_lo, c_hi, args_ok);
        s = h.s; c_lo = h.c_lo; c_hi = h.c_hi;
        if (s > m.lower + (kWindow - 1)) flags |= ST_NO_SYMBOL;
    }
    rans_pop32(hi, lo, (uint32_t)c_lo, (uint32_t)(c_hi - c_lo));
    return s;
}

FLIC_HD int decode_symbol_lean(uint32_t& hi, uint32_t& lo, float mean, float scale, ExpTab tab,
                               ParamGuard& guard, int32_t& flags) {
    const uint32_t mod = lo & kProbMask;
    const SymbolModel m = make_model(mean, scale);
    guard_note(guard, mean, scale);
    const int g = guess_symbol(mod, mean, scale, m.lower);
    int c_lo, c_hi;
    const bool args_ok = cdf_pair_try(g, m, tab, c_lo, c_hi);
    int s = g;
    // (a guess on the window's left edge whose CDF(g - 1) exceeds mod -- only a corrupt stream has
    // that -- is sorted out by the search function: the edge is virtual)
    if (!(c_lo <= (int)mod && c_hi > (int)mod && guess_in_window(g, m.lower) && args_ok)) {
        const SymbolHit h = decode_symbol_search(mod, mean, scale, tab, g, c_lo, c_hi, args_ok);
        s = h.s; c_lo = h.c_lo; c_hi = h.c_hi;
        if (s > m.lower + (kWindow - 1)) flags |= ST_NO_SYMBOL;
    }
    rans_pop32(hi, lo, (uint32_t)c_lo, (uint32_t)(c_hi - c_lo));
    return s;
}

}  // namespace flic
