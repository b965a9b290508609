// F_{p^2} = F_p[i]/(i^2+1), p = 2^61-1  --  Virgo's field, device + host.
//
// Replaces: /root/reference/lib/virgo/src/fieldElement.cpp:34-96 (operator+,*,-),
//           :336-360 (myMod, mymult); storage layout fieldElement.hpp:96-97
//           ({u64 real; u64 img;}, 16 bytes, always canonical: both limbs < p).
//
// Every function here returns CANONICAL values (< p) unless its name says otherwise. The
// arithmetic is exact, so any association order gives the same bits as the reference.
//
// Multiplication (the hot operation) is built for the sm_100a integer pipes:
//   * operands are split into two 31-bit limbs  x = x1*2^31 + x0;
//   * a complex product is two chains of four IMAD.WIDE.U32 per component (schoolbook, the
//     subtraction in re = ac - bd is folded in by multiplying with p - b instead of b), accumulated
//     in 64-bit registers with NO carry handling: four 31x31-bit products cannot overflow 64 bits;
//   * weights: 2^62 = 2 (mod p) is absorbed by pre-doubling one hi limb, so a component is just
//     u + 2^31 * t  with two 64-bit partials, reduced ONCE with the Mersenne identities
//     2^61 = 1: (u & p) + (u >> 61) + (t >> 30) + ((t mod 2^30) << 31).
// This moves the work to the FMA pipe (IMAD.WIDE) and leaves ~16 ALU-pipe instructions per
// reduced component; see DESIGN.md "Field arithmetic" for the instruction budget.
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
VP_HD u64 fp_red1(u64 x) { return x >= P ? x - P : x; }          // x < 2p  ->  [0,p)
VP_HD u64 fp_fold(u64 x) { return (x & P) + (x >> 61); }         // any u64 -> [0, p + 7], same class
VP_HD u64 fp_canon(u64 x) { return fp_red1(fp_fold(x)); }        // any u64 < 2^64 - 2^61 -> [0,p)
VP_HD u64 fp_add(u64 a, u64 b) { return fp_red1(a + b); }
VP_HD u64 fp_sub(u64 a, u64 b) { return fp_red1(a + (P - b)); }
VP_HD u64 fp_neg(u64 a) { return a ? P - a : 0; }

// x = hi*2^31 + lo
struct Limbs {
    u32 lo, hi;
};
VP_HD Limbs split31(u64 x) {  // x < 2^63
    Limbs r;
    r.lo = (u32)x & 0x7FFFFFFFu;
    r.hi = (u32)(x >> 31);
    return r;
}
// 32x32 -> 64 multiply and multiply-accumulate. On the device these are spelled as PTX mul.wide / mad.wide:
// the C expression (u64)a * b + c makes nvcc 12.9 emit one junk "add 0 to the high word" (VIADD) per product.
VP_HD u64 mul32(u32 a, u32 b) {
#if defined(__CUDA_ARCH__)
    u64 d;
    asm("mul.wide.u32 %0, %1, %2;" : "=l"(d) : "r"(a), "r"(b));
    return d;
#else
    return (u64)a * (u64)b;
#endif
}
VP_HD u64 mad32(u32 a, u32 b, u64 c) {  // a*b + c  (IMAD.WIDE.U32 with a 64-bit addend)
#if defined(__CUDA_ARCH__)
    u64 d;
    asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(d) : "r"(a), "r"(b), "l"(c));
    return d;
#else
    return (u64)a * (u64)b + c;
#endif
}

// value = u + 2^31*t + extra (mod p), canonical.  u, t < 2^64; extra < 2^62.
VP_HD u64 fp_reduce_ut(u64 u, u64 t, u64 extra) {
    u64 s = (u & P) + (u >> 61) + (t >> 30) + ((t & 0x3FFFFFFFULL) << 31) + extra;  // < 2^63 + 2^35
    return fp_canon(s);
}
// Same value, folded once only: result <= p + 4 (same residue class, not canonical). extra < 2^62.
VP_HD u64 fp_reduce_ut_loose(u64 u, u64 t, u64 extra) {
    u64 s = (u & P) + (u >> 61) + (t >> 30) + ((t & 0x3FFFFFFFULL) << 31) + extra;
    return fp_fold(s);
}
// value = u + 2^31*t + 2^62*w + extra (mod p), canonical.  u, t, w < 2^64; extra < 2^62.
VP_HD u64 fp_reduce_utw(u64 u, u64 t, u64 w, u64 extra) {
    // 2^62 * w = 2w: (w << 1) may drop bit 63 of w, worth 2^64 = 8 (mod p)
    u64 s = (u & P) + (u >> 61) + (t >> 30) + ((t & 0x3FFFFFFFULL) << 31) + extra;   // < 2^63 + 2^35
    u64 w2 = ((w << 1) & P) + ((w << 1) >> 61) + ((w >> 63) << 3);                   // < 2^61 + 16
    return fp_canon(s + w2);
}

VP_HD u64 fp_mul(u64 a, u64 b) {  // a, b < 2^62
    const Limbs x = split31(a), y = split31(b);
    const u64 u = mul32(x.lo, y.lo), t = mad32(x.hi, y.lo, mul32(x.lo, y.hi)), w = mul32(x.hi, y.hi);
    return fp_reduce_utw(u, t, w, 0);
}

// ---------------------------------------------------------------- extension field: basic ops
VP_HD F f_zero() { return F{0, 0}; }
VP_HD F f_one() { return F{1, 0}; }
VP_HD F f_make(u64 re, u64 im) { return F{re, im}; }
VP_HD bool f_is_zero(const F& a) { return (a.re | a.im) == 0; }
VP_HD bool f_eq(const F& a, const F& b) { return a.re == b.re && a.im == b.im; }
VP_HD F f_add(const F& a, const F& b) { return F{fp_add(a.re, b.re), fp_add(a.im, b.im)}; }
VP_HD F f_sub(const F& a, const F& b) { return F{fp_sub(a.re, b.re), fp_sub(a.im, b.im)}; }
VP_HD F f_neg(const F& a) { return F{fp_neg(a.re), fp_neg(a.im)}; }
VP_HD F f_dbl(const F& a) { return f_add(a, a); }

// ---------------------------------------------------------------- complex product by limb chains
// Left operand m (re, im < 2^62): limbs of re, im and of the negated im (2p - im).
struct LOp {
    u32 re0, re1, im0, im1, nim0, nim1;
};
// Right operand v (re, im < 2^62): limbs of re, im.
struct ROp {
    u32 re0, re1, im0, im1;
};
VP_HD LOp make_lop(u64 re, u64 im) {  // re, im <= 2p
    const Limbs a = split31(re), b = split31(im), c = split31(2 * P - im);
    return LOp{a.lo, a.hi, b.lo, b.hi, c.lo, c.hi};
}
VP_HD ROp make_rop(u64 re, u64 im) {
    const Limbs a = split31(re), b = split31(im);
    return ROp{a.lo, a.hi, b.lo, b.hi};
}

// Partials of m*v: component = u + 2^31 t + 2^62 w.
struct CPart {
    u64 u_re, t_re, w_re, u_im, t_im, w_im;
};
// Valid for m, v components <= 2p (hi limbs < 2^31): every chain is a sum of <= 4 products < 2^62.
VP_HD CPart cprod_parts(const LOp& m, const ROp& v) {
    CPart c;
    c.u_re = mad32(m.nim0, v.im0, mul32(m.re0, v.re0));
    c.t_re = mad32(m.nim1, v.im0, mad32(m.nim0, v.im1, mad32(m.re1, v.re0, mul32(m.re0, v.re1))));
    c.w_re = mad32(m.nim1, v.im1, mul32(m.re1, v.re1));
    c.u_im = mad32(m.im0, v.re0, mul32(m.re0, v.im0));
    c.t_im = mad32(m.im1, v.re0, mad32(m.im0, v.re1, mad32(m.re1, v.im0, mul32(m.re0, v.im1))));
    c.w_im = mad32(m.im1, v.re1, mul32(m.re1, v.im1));
    return c;
}
// m*v + acc for operands with components <= 2p; acc < 2^62. Canonical result.
VP_HD F f_mul_add_loose(const LOp& m, const ROp& v, const F& acc) {
    const CPart c = cprod_parts(m, v);
    return F{fp_reduce_utw(c.u_re, c.t_re, c.w_re, acc.re), fp_reduce_utw(c.u_im, c.t_im, c.w_im, acc.im)};
}
// Same with a result that is only folded once (<= p + 5, not canonical).
VP_HD u64 fp_reduce_utw_loose(u64 u, u64 t, u64 w, u64 extra) {
    u64 s = (u & P) + (u >> 61) + (t >> 30) + ((t & 0x3FFFFFFFULL) << 31) + extra;
    u64 w2 = ((w << 1) & P) + ((w << 1) >> 61) + ((w >> 63) << 3);
    return fp_fold(s + w2);
}
VP_HD F f_mul_add_loose2(const LOp& m, const ROp& v, const F& acc) {
    const CPart c = cprod_parts(m, v);
    return F{fp_reduce_utw_loose(c.u_re, c.t_re, c.w_re, acc.re), fp_reduce_utw_loose(c.u_im, c.t_im, c.w_im, acc.im)};
}

// Canonical operands (< 2^61: hi limbs < 2^30) let the 2^62-weight products ride in the u chain by
// pre-doubling the right operand's hi limbs: two partials per component instead of three.
struct ROpD {
    u32 re0, re1, re1d, im0, im1, im1d;
};
VP_HD ROpD make_ropd(const F& v) {  // v canonical
    const Limbs a = split31(v.re), b = split31(v.im);
    return ROpD{a.lo, a.hi, a.hi << 1, b.lo, b.hi, b.hi << 1};
}
// m components <= 2p, v canonical: u chains hold 2 x (< 2^62) + 2 x (2^31 * 2^31) < 2^64.
VP_HD F f_mul_add_k(const LOp& m, const ROpD& v, const F& acc) {
    const u64 u_re = mad32(m.nim1, v.im1d, mad32(m.re1, v.re1d, mad32(m.nim0, v.im0, mul32(m.re0, v.re0))));
    const u64 t_re = mad32(m.nim1, v.im0, mad32(m.nim0, v.im1, mad32(m.re1, v.re0, mul32(m.re0, v.re1))));
    const u64 u_im = mad32(m.im1, v.re1d, mad32(m.re1, v.im1d, mad32(m.im0, v.re0, mul32(m.re0, v.im0))));
    const u64 t_im = mad32(m.im1, v.re0, mad32(m.im0, v.re1, mad32(m.re1, v.im0, mul32(m.re0, v.im1))));
    return F{fp_reduce_ut(u_re, t_re, acc.re), fp_reduce_ut(u_im, t_im, acc.im)};
}
// Same, but the result is only folded once (<= p + 4, NOT canonical): for running sums that are fed back as `acc`.
VP_HD F f_mul_add_k_loose(const LOp& m, const ROpD& v, const F& acc) {
    const u64 u_re = mad32(m.nim1, v.im1d, mad32(m.re1, v.re1d, mad32(m.nim0, v.im0, mul32(m.re0, v.re0))));
    const u64 t_re = mad32(m.nim1, v.im0, mad32(m.nim0, v.im1, mad32(m.re1, v.re0, mul32(m.re0, v.re1))));
    const u64 u_im = mad32(m.im1, v.re1d, mad32(m.re1, v.im1d, mad32(m.im0, v.re0, mul32(m.re0, v.im0))));
    const u64 t_im = mad32(m.im1, v.re0, mad32(m.im0, v.re1, mad32(m.re1, v.im0, mul32(m.re0, v.im1))));
    return F{fp_reduce_ut_loose(u_re, t_re, acc.re), fp_reduce_ut_loose(u_im, t_im, acc.im)};
}

VP_HD F f_mul(const F& a, const F& b) {  // canonical operands
    return f_mul_add_k(make_lop(a.re, a.im), make_ropd(b), f_zero());
}
VP_HD F f_mul_add(const F& a, const F& b, const F& acc) { return f_mul_add_k(make_lop(a.re, a.im), make_ropd(b), acc); }
// a * b where b is in the base field (b.im == 0)
VP_HD F f_mul_base(const F& a, u64 b) { return F{fp_mul(a.re, b), fp_mul(a.im, b)}; }

// v0 + r*(v1 - v0): linear_poly{v1-v0, v0}.eval(r)   (reference src/polynomial.cpp:128-131)
// FoldK = the challenge r pre-split (built once per kernel).
typedef ROpD FoldK;
VP_HD FoldK make_foldk(const F& r) { return make_ropd(r); }
VP_HD F f_fold_k(const F& v0, const F& v1, const FoldK& k) {
    // d = v1 - v0 + p  in (0, 2p)
    return f_mul_add_k(make_lop(v1.re + P - v0.re, v1.im + P - v0.im), k, v0);
}
VP_HD F f_fold(const F& v0, const F& v1, const F& r) { return f_fold_k(v0, v1, make_foldk(r)); }

// ---------------------------------------------------------------- lazy sums of products
// Acc: unreduced 128-bit sums of the Karatsuba base products; kept for the dot-product kernels.
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
    u64 h = (x.hi << 3) | (x.lo >> 61);                     // bits 61..124 of x
    u64 s = fp_fold(h) + (x.lo & P) + ((x.hi >> 61) << 3);  // < 2^62 + 64
    return fp_canon(s);
}
struct Acc {
    U128 ac, bd, x;  // sum a*c, sum b*d, sum (a+b)*(c+d); up to 8 products between reductions
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

}  // namespace vp
