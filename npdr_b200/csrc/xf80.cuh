// xf80.cuh -- software emulation of x87 80-bit extended arithmetic (64-bit significand,
// round-to-nearest-even), usable on host and device.
//
// Why: the reference's neighbourhood radii come from base-R colSums()/sd()
// (R/nearestNeighbors.R:244-246), which accumulate in C `long double` = x87 extended on
// x86-64.  The multiSURF/SURF membership test `d < radius` (R/nearestNeighbors.R:630) must
// give bit-identical neighbour lists, so the device reproduces those sums bit for bit
// instead of approximating them.  Range: finite, normal-range values only (distances and
// their moments); no infinities, NaNs, denormals or exponent overflow handling.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define XF_HD __host__ __device__ __forceinline__
#else
#define XF_HD inline
#endif

struct xf80 {
    uint64_t m;   // significand, bit 63 set unless the value is zero
    int32_t e;    // value = (-1)^neg * m * 2^(e - 63)
    int32_t neg;
};

XF_HD int xf_clz64(uint64_t x) {
#if defined(__CUDA_ARCH__)
    return __clzll((long long)x);
#else
    return __builtin_clzll(x);
#endif
}

XF_HD uint64_t xf_mulhi(uint64_t a, uint64_t b) {
#if defined(__CUDA_ARCH__)
    return __umul64hi(a, b);
#else
    return (uint64_t)(((unsigned __int128)a * b) >> 64);
#endif
}

XF_HD uint64_t xf_d2bits(double d) {
#if defined(__CUDA_ARCH__)
    return (uint64_t)__double_as_longlong(d);
#else
    union { double d; uint64_t u; } c; c.d = d; return c.u;
#endif
}
XF_HD double xf_bits2d(uint64_t u) {
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)u);
#else
    union { double d; uint64_t u; } c; c.u = u; return c.d;
#endif
}

XF_HD xf80 xf_zero() { xf80 r; r.m = 0; r.e = 0; r.neg = 0; return r; }

XF_HD xf80 xf_from_double(double d) {
    uint64_t b = xf_d2bits(d);
    xf80 r;
    r.neg = (int32_t)(b >> 63);
    int32_t ex = (int32_t)((b >> 52) & 0x7FF);
    uint64_t fr = b & 0xFFFFFFFFFFFFFull;
    if (ex == 0) {
        if (fr == 0) { r.m = 0; r.e = 0; return r; }
        int lz = xf_clz64(fr);              // denormal: normalise
        r.m = fr << lz;
        r.e = -1022 - (lz - 11);
        return r;
    }
    r.m = (1ull << 63) | (fr << 11);
    r.e = ex - 1023;
    return r;
}

XF_HD xf80 xf_from_int(int64_t v) {
    xf80 r;
    if (v == 0) return xf_zero();
    r.neg = v < 0;
    uint64_t a = (uint64_t)(v < 0 ? -v : v);
    int lz = xf_clz64(a);
    r.m = a << lz;
    r.e = 63 - lz;
    return r;
}

// round a 128-bit significand (hi:lo, hi normalised) to 64 bits, nearest-even
XF_HD xf80 xf_round_pack(int32_t neg, int32_t e, uint64_t hi, uint64_t lo) {
    const uint64_t half = 1ull << 63;
    if ((lo > half) || (lo == half && (hi & 1ull))) {
        hi += 1;
        if (hi == 0) { hi = half; e += 1; }
    }
    xf80 r; r.m = hi; r.e = e; r.neg = neg;
    return r;
}

XF_HD xf80 xf_add(xf80 a, xf80 b) {
    if (a.m == 0) return b;
    if (b.m == 0) return a;
    // make |a| >= |b|
    if (b.e > a.e || (b.e == a.e && b.m > a.m)) { xf80 t = a; a = b; b = t; }
    int32_t sh = a.e - b.e;
    uint64_t bh, bl;
    if (sh == 0) { bh = b.m; bl = 0; }
    else if (sh < 64) { bh = b.m >> sh; bl = b.m << (64 - sh); }
    else if (sh == 64) { bh = 0; bl = b.m; }
    else if (sh < 128) { bh = 0; bl = (b.m >> (sh - 64)) | (((b.m << (128 - sh)) != 0) ? 1ull : 0ull); }
    else { bh = 0; bl = 1ull; }
    uint64_t hi, lo;
    int32_t e = a.e;
    if (a.neg == b.neg) {
        lo = bl;                              // a's low word is zero
        hi = a.m + bh;
        if (hi < a.m) {                       // carry out: shift right one, keep sticky
            lo = (lo >> 1) | (lo & 1ull) | (hi << 63);
            hi = (hi >> 1) | (1ull << 63);
            e += 1;
        }
    } else {
        lo = 0ull - bl;
        hi = a.m - bh - (bl != 0 ? 1ull : 0ull);
        if (hi == 0 && lo == 0) return xf_zero();
        if (hi == 0) { hi = lo; lo = 0; e -= 64; }
        int lz = xf_clz64(hi);
        if (lz) {
            hi = (hi << lz) | (lo >> (64 - lz));
            lo <<= lz;
            e -= lz;
        }
    }
    return xf_round_pack(a.neg, e, hi, lo);
}

XF_HD xf80 xf_neg(xf80 a) { a.neg ^= 1; return a; }
XF_HD xf80 xf_sub(xf80 a, xf80 b) { return xf_add(a, xf_neg(b)); }

XF_HD xf80 xf_mul(xf80 a, xf80 b) {
    if (a.m == 0 || b.m == 0) return xf_zero();
    uint64_t hi = xf_mulhi(a.m, b.m);
    uint64_t lo = a.m * b.m;
    int32_t e = a.e + b.e + 1;
    if (!(hi >> 63)) {
        hi = (hi << 1) | (lo >> 63);
        lo <<= 1;
        e -= 1;
    }
    return xf_round_pack(a.neg ^ b.neg, e, hi, lo);
}

// a / b, correctly rounded (restoring long division; used a handful of times per instance)
XF_HD xf80 xf_div(xf80 a, xf80 b) {
    if (a.m == 0) return xf_zero();
    uint64_t nh, nl;
    int32_t e = a.e - b.e;
    if (a.m >= b.m) { nh = a.m >> 1; nl = a.m << 63; }
    else { nh = a.m; nl = 0; e -= 1; }
    uint64_t q = 0, r = nh, d = b.m;        // invariant: r < d
    for (int i = 63; i >= 0; --i) {
        uint64_t carry = r >> 63;
        r = (r << 1) | ((nl >> i) & 1ull);
        if (carry || r >= d) { r -= d; q |= (1ull << i); }
    }
    // round: compare 2r with d without overflow
    uint64_t rem2_gt = (r > d - r);
    uint64_t rem2_eq = (r == d - r);
    if (rem2_gt || (rem2_eq && (q & 1ull))) {
        q += 1;
        if (q == 0) { q = 1ull << 63; e += 1; }
    }
    xf80 res; res.m = q; res.e = e; res.neg = a.neg ^ b.neg;
    return res;
}

// (double) of an extended value: round 64 -> 53 bits, nearest-even (normal range only)
XF_HD double xf_to_double(xf80 a) {
    if (a.m == 0) return a.neg ? -0.0 : 0.0;
    uint64_t keep = a.m >> 11;
    uint64_t rem = a.m & 0x7FFull;
    int32_t e = a.e;
    if (rem > 0x400ull || (rem == 0x400ull && (keep & 1ull))) {
        keep += 1;
        if (keep == (1ull << 53)) { keep >>= 1; e += 1; }
    }
    int32_t ex = e + 1023;
    if (ex <= 0) return a.neg ? -0.0 : 0.0;                 // below normal range: flushed (documented)
    if (ex >= 0x7FF) return xf_bits2d(((uint64_t)a.neg << 63) | 0x7FF0000000000000ull);
    return xf_bits2d(((uint64_t)a.neg << 63) | ((uint64_t)ex << 52) | (keep & 0xFFFFFFFFFFFFFull));
}
