// omc_format.cuh -- printf("%e ") / printf("%f ") of a double as fixed-width text, exactly as glibc prints it, usable on the device.
//
// outputResults() of the reference (ucodes/omc_dosxyz/omc_dosxyz.c:841-879) writes the dose block with "%e " and the
// uncertainty block with "%f ", one fprintf per voxel; at 1 mm voxels (8e7 of them) that text conversion is the wall-clock
// bottleneck SURVEY.md 8f-2 names.  Here the conversion runs where the numbers already are (format_kernel, omc_lockstep.cu) and
// the host only streams bytes to the file.
//
// "%e": 13 bytes "d.dddddde+XX " -- 7 significant digits, correctly rounded (round-half-even on the exact binary value, what glibc
// does in the default rounding mode).  v = M * 2^E is multiplied by a 128-bit approximation of 10^-(k-6) (relative error < 2^-126,
// table built exactly on the host by big-integer arithmetic, Pow10Table below), giving the integer digits and > 100 fraction
// bits; the rounding direction is certain unless the fraction is within 2^-80 of one half.  Those cases (exact ties included),
// negative numbers, NaN/Inf and three-digit exponents are NOT guessed: the function returns 1 and the host formats that one value
// with snprintf (their share is ~2^-79 of random inputs; exact ties need a dyadic rational with a short decimal expansion).
// "%f": 9 bytes "d.dddddd " for 0 <= v < 9.9999995: v * 10^6 is EXACT in 128-bit integer arithmetic, so round-half-even is decided
// exactly on the device; anything else returns 1.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define OMC_HD __host__ __device__ __forceinline__
#else
#define OMC_HD static inline
#endif

namespace omc {

typedef unsigned __int128 u128;

struct Pow10 {           // 10^-q ~= (hi * 2^64 + lo) * 2^pe, 2^127 <= mantissa < 2^128
    uint64_t hi, lo;
    int32_t pe, pad;
};
constexpr int kPow10Min = -330, kPow10Max = 330;           // q range: doubles span 4.9e-324 .. 1.8e308, q = k - 6
constexpr int kPow10N = kPow10Max - kPow10Min + 1;
constexpr int kFmtEWidth = 13, kFmtFWidth = 9;

OMC_HD uint64_t dbits(double v) {
#if defined(__CUDA_ARCH__)
    return (uint64_t)__double_as_longlong(v);
#else
    union { double d; uint64_t u; } c;
    c.d = v;
    return c.u;
#endif
}

OMC_HD int clz64(uint64_t x) {
#if defined(__CUDA_ARCH__)
    return __clzll((long long)x);
#else
    return __builtin_clzll(x);
#endif
}

OMC_HD void put7(char *out, uint32_t d) {           // "d.dddddd" from a 7-digit integer
    char t[7];
    for (int i = 6; i >= 0; i--) { t[i] = (char)('0' + d % 10u); d /= 10u; }
    out[0] = t[0]; out[1] = '.';
    for (int i = 1; i < 7; i++) out[1 + i] = t[i];
}

// returns 0 and fills out[13] when the text is certain, 1 when the host must format this value
OMC_HD int fmt_e(double v, const Pow10 *__restrict__ tab, char *out) {
    const uint64_t b = dbits(v);
    if (b >> 63) return 1;                                   // negative (or -0.0): other width
    const int be = (int)(b >> 52);
    uint64_t m = b & 0xFFFFFFFFFFFFFull;
    if (be == 0x7FF) return 1;                               // inf / nan
    if (be == 0 && m == 0) {
        const char z[14] = "0.000000e+00 ";
        for (int i = 0; i < 13; i++) out[i] = z[i];
        return 0;
    }
    int e2;                                                  // v = m * 2^e2 with m normalised to 64 bits
    if (be == 0) { const int s = clz64(m); m <<= s; e2 = -1074 - s; }
    else { m = (m | (1ull << 52)) << 11; e2 = be - 1075 - 11; }
    const int lg2 = e2 + 63;                                 // floor(log2 v)
    int k = (lg2 * 78913) >> 18;                             // floor(lg2 * log10(2)): k or k - 1
    uint32_t d = 0;
    bool up = false;
    int tries = 0;
    for (;; tries++) {
        if (tries == 3) return 1;
        const int q = k - 6;
        if (q < kPow10Min || q > kPow10Max) return 1;
        const Pow10 p = tab[q - kPow10Min];
        const u128 a = (u128)m * p.hi, bl = (u128)m * p.lo;
        const u128 t = a + (uint64_t)(bl >> 64);             // top 128 bits of m * mantissa, in [2^126, 2^128)
        const int s = -(64 + e2 + p.pe);                     // value = t * 2^-s
        if (s > 127) { k--; continue; }                      // (cannot happen for a correct k; defensive)
        if (s < 100) { k++; continue; }
        const u128 di = t >> s;
        if (di >= 10000000u) { k++; continue; }
        if (di < 1000000u) { k--; continue; }
        const u128 frac = t & ((((u128)1) << s) - 1), half = ((u128)1) << (s - 1);
        const u128 diff = frac > half ? frac - half : half - frac;
        if (diff < (((u128)1) << 24)) return 1;              // within 2^-(s-25) <= 2^-75 of a tie: let the host decide exactly
        up = frac > half;
        d = (uint32_t)di;
        break;
    }
    if (up && ++d == 10000000u) { d = 1000000u; k++; }
    const int ak = k < 0 ? -k : k;
    if (ak > 99) return 1;                                   // three-digit exponent: other width
    put7(out, d);
    out[8] = 'e'; out[9] = k < 0 ? '-' : '+';
    out[10] = (char)('0' + ak / 10); out[11] = (char)('0' + ak % 10); out[12] = ' ';
    return 0;
}

// "%f " with six decimals, exact; 1 = the host must format this value (negative, >= 9.9999995, inf, nan)
OMC_HD int fmt_f(double v, char *out) {
    const uint64_t b = dbits(v);
    if (b >> 63) return 1;
    const int be = (int)(b >> 52);
    uint64_t m = b & 0xFFFFFFFFFFFFFull;
    if (be == 0x7FF) return 1;
    int e2;
    if (be == 0) e2 = -1074; else { m |= 1ull << 52; e2 = be - 1075; }
    if (e2 > -49) return 1;                                  // v >= 16
    const u128 t = (u128)m * 1000000u;                       // exact, < 2^73
    const int s = -e2;                                       // v * 10^6 = t * 2^-s, s >= 49
    uint32_t d;
    if (s > 127) d = 0;                                      // < 2^-54: rounds to 0
    else {
        const u128 di = t >> s;
        if (di >= 10000000u) return 1;
        const u128 frac = t & ((((u128)1) << s) - 1), half = ((u128)1) << (s - 1);
        d = (uint32_t)di;
        if (frac > half || (frac == half && (d & 1u))) d++;
        if (d >= 10000000u) return 1;                        // "10.000000": other width
    }
    put7(out, d);
    out[8] = ' ';
    return 0;
}

// Host-side, exact: 10^|q| as a little-endian big integer; mantissa = its top 128 bits (q <= 0) or the top 128 bits of its
// reciprocal by binary long division (q > 0).  Truncated, so the relative error of every entry is below 2^-127.
static inline void build_pow10_table(Pow10 *tab) {
    enum { NL = 40 };                                        // 10^330 < 2^1097 < 2^(32*40)
    for (int q = kPow10Min; q <= kPow10Max; q++) {
        uint32_t big[NL] = {1};
        const int n = q < 0 ? -q : q;
        for (int i = 0; i < n; i++) {
            uint64_t c = 0;
            for (int j = 0; j < NL; j++) { c += (uint64_t)big[j] * 10u; big[j] = (uint32_t)c; c >>= 32; }
        }
        int top = NL - 1;
        while (top > 0 && big[top] == 0) top--;
        const int L = 32 * top + (32 - __builtin_clz(big[top]));          // bit length of 10^n
        auto bit = [&](int i) -> unsigned { return i < 0 ? 0u : (big[i >> 5] >> (i & 31)) & 1u; };
        Pow10 &p = tab[q - kPow10Min];
        p.pad = 0;
        if (q <= 0) {                                        // 10^-q = 10^n: top 128 bits
            u128 mant = 0;
            for (int i = 0; i < 128; i++) mant = (mant << 1) | bit(L - 1 - i);
            p.hi = (uint64_t)(mant >> 64); p.lo = (uint64_t)mant; p.pe = L - 128;
        } else {                                             // 10^-q = 1 / 10^n: floor(2^(L+127) / 10^n), 129 quotient bits, the first is 0
            uint32_t r[NL + 1] = {0};
            r[(L - 1) >> 5] = 1u << ((L - 1) & 31);          // R = 2^(L-1) <= D
            u128 quo = 0;
            for (int it = 0; it <= 128; it++) {
                int ge = 1;                                  // R >= D ?
                if (r[NL] == 0) {
                    for (int j = NL - 1; j >= 0; j--)
                        if (r[j] != big[j]) { ge = r[j] > big[j]; break; }
                }
                if (ge) {
                    int64_t bw = 0;
                    for (int j = 0; j < NL; j++) {
                        int64_t x = (int64_t)r[j] - big[j] + bw;
                        bw = x < 0 ? -1 : 0;
                        r[j] = (uint32_t)x;
                    }
                    r[NL] = (uint32_t)((int64_t)r[NL] + bw);
                }
                quo = (quo << 1) | (unsigned)ge;
                uint32_t c = 0;
                for (int j = 0; j <= NL; j++) { const uint32_t nc = r[j] >> 31; r[j] = (r[j] << 1) | c; c = nc; }
            }
            p.hi = (uint64_t)(quo >> 64); p.lo = (uint64_t)quo; p.pe = -(L + 127);
        }
    }
}

}  // namespace omc
