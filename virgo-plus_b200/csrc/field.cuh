// F_{p^2} = F_p[i]/(i^2+1), p = 2^61-1  --  Virgo's field, device + host.
//
// Replaces: /root/reference/lib/virgo/src/fieldElement.cpp:34-96 (operator+,*,-),
//           :336-360 (myMod, mymult); storage layout fieldElement.hpp:96-97
//           ({u64 real; u64 img;}, 16 bytes, always canonical: both limbs < p).
//
// All operations here return canonical values (< p). Because the arithmetic is exact, any
// association order yields the same bits as the reference; only canonical output matters.
//
// Device multiply: the 61x61->122 bit product is built from 32-bit limbs (IMAD.WIDE.U32 chains),
// and reduced with the Mersenne identities 2^61 = 1, 2^64 = 8 (mod p). Sum-of-products code
// (round polynomial accumulation) uses the lazy `Acc` type that keeps unreduced 128-bit sums of
// the three Karatsuba products and folds once at the end.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define VP_HD __host__ __device__ __forceinline__
#define VP_D __device__ __forceinline__
#else
#define VP_HD inline
#define VP_D inline
#endif

namespace vp {

typedef unsigned long long u64;
typedef unsigned int u32;

static constexpr u64 P = 2305843009213693951ULL;  // 2^61 - 1

struct alignas(16) F {
    u64 re, im;
};

// ---------------------------------------------------------------- base field F_p
VP_HD u64 fp_red1(u64 x) {  // x < 2p  ->  [0,p)
    return x >= P ? x - P : x;
}
VP_HD u64 fp_fold(u64 x) {  // any u64 -> [0, p + 7]
    return (x & P) + (x >> 61);
}
VP_HD u64 fp_add(u64 a, u64 b) { return fp_red1(a + b); }
VP_HD u64 fp_sub(u64 a, u64 b) { return fp_red1(a + (P - b)); }
VP_HD u64 fp_neg(u64 a) { return a ? P - a : 0; }

// 128-bit unsigned value as two u64 (device has no __int128).
struct U128 {
    u64 lo, hi;
};

VP_HD U128 mul_wide(u64 a, u64 b) {
#if defined(__CUDA_ARCH__)
    U128 r;
    r.lo = a * b;
    r.hi = __umul64hi(a, b);
    return r;
#else
    unsigned __int128 t = (unsigned __int128)a * b;
    U128 r;
    r.lo = (u64)t;
    r.hi = (u64)(t >> 64);
    return r;
#endif
}

VP_HD U128 add128(U128 a, U128 b) {
    U128 r;
    r.lo = a.lo + b.lo;
    r.hi = a.hi + b.hi + (r.lo < a.lo ? 1ULL : 0ULL);
    return r;
}

// Reduce any 128-bit value to [0,p).  x = hi*2^64 + lo;  2^61 = 1, 2^64 = 8, 2^125 = 8 (mod p).
VP_HD u64 fp_reduce128(U128 x) {
    u64 h = (x.hi << 3) | (x.lo >> 61);                        // bits 61..124 of x
    u64 s = fp_fold(h) + (x.lo & P) + ((x.hi >> 61) << 3);     // < 2^62 + 64
    s = fp_fold(s);                                            // <= p + 2
    return fp_red1(s);
}

VP_HD u64 fp_mul(u64 a, u64 b) {  // a,b < 2^62
    return fp_reduce128(mul_wide(a, b));
}

// ---------------------------------------------------------------- extension field
VP_HD F f_zero() { return F{0, 0}; }
VP_HD F f_one() { return F{1, 0}; }
VP_HD F f_make(u64 re, u64 im) { return F{re, im}; }
VP_HD bool f_is_zero(const F& a) { return (a.re | a.im) == 0; }
VP_HD bool f_eq(const F& a, const F& b) { return a.re == b.re && a.im == b.im; }
VP_HD F f_add(const F& a, const F& b) { return F{fp_add(a.re, b.re), fp_add(a.im, b.im)}; }
VP_HD F f_sub(const F& a, const F& b) { return F{fp_sub(a.re, b.re), fp_sub(a.im, b.im)}; }
VP_HD F f_neg(const F& a) { return F{fp_neg(a.re), fp_neg(a.im)}; }
VP_HD F f_dbl(const F& a) { return f_add(a, a); }

// Lazy accumulator for sums of F-products. Holds unreduced 128-bit sums of the three Karatsuba
// base products; each product < 2^124, so up to 8 products may be added before a fold is needed
// (we fold on `acc_compress`).
struct Acc {
    U128 ac, bd, x;  // sum a*c, sum b*d, sum (a+b)*(c+d)
};
VP_HD Acc acc_zero() { return Acc{{0, 0}, {0, 0}, {0, 0}}; }
VP_HD void acc_mad(Acc& s, const F& a, const F& b) {
    s.ac = add128(s.ac, mul_wide(a.re, b.re));
    s.bd = add128(s.bd, mul_wide(a.im, b.im));
    s.x = add128(s.x, mul_wide(a.re + a.im, b.re + b.im));
}
VP_HD F acc_reduce(const Acc& s) {
    u64 ac = fp_reduce128(s.ac), bd = fp_reduce128(s.bd), x = fp_reduce128(s.x);
    F r;
    r.re = fp_sub(ac, bd);
    r.im = fp_sub(fp_sub(x, ac), bd);
    return r;
}

VP_HD F f_mul(const F& a, const F& b) {
    u64 ac = fp_mul(a.re, b.re), bd = fp_mul(a.im, b.im);
    u64 x = fp_mul(a.re + a.im, b.re + b.im);
    F r;
    r.re = fp_sub(ac, bd);
    r.im = fp_sub(fp_sub(x, ac), bd);
    return r;
}

// a * b where b is in the base field (b.im == 0)
VP_HD F f_mul_base(const F& a, u64 b) { return F{fp_mul(a.re, b), fp_mul(a.im, b)}; }

// v0 + r*(v1 - v0): linear_poly{v1-v0, v0}.eval(r)   (reference src/polynomial.cpp:128-131)
VP_HD F f_fold(const F& v0, const F& v1, const F& r) { return f_add(v0, f_mul(f_sub(v1, v0), r)); }

}  // namespace vp
