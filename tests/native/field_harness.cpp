// Host build of virgo-plus_b200/csrc/field.cuh (the same inline code nvcc compiles for the device)
// exposed to pytest through ctypes; checked against Python big-integer arithmetic.
#include "../../virgo-plus_b200/csrc/field.cuh"
using namespace vp;
extern "C" {
void h_mul(const F* a, const F* b, F* out, int n) { for (int i = 0; i < n; ++i) out[i] = f_mul(a[i], b[i]); }
void h_mul_add(const F* a, const F* b, const F* c, F* out, int n) { for (int i = 0; i < n; ++i) out[i] = f_mul_add(a[i], b[i], c[i]); }
void h_fold(const F* v0, const F* v1, const F* r, F* out, int n) { for (int i = 0; i < n; ++i) out[i] = f_fold(v0[i], v1[i], r[i]); }
void h_add(const F* a, const F* b, F* out, int n) { for (int i = 0; i < n; ++i) out[i] = f_add(a[i], b[i]); }
void h_sub(const F* a, const F* b, F* out, int n) { for (int i = 0; i < n; ++i) out[i] = f_sub(a[i], b[i]); }
void h_fp_mul(const u64* a, const u64* b, u64* out, int n) { for (int i = 0; i < n; ++i) out[i] = fp_mul(a[i], b[i]); }
// loose product: a, b components up to 2p (not canonical), acc canonical
void h_mul_add_loose(const F* a, const F* b, const F* c, F* out, int n) {
    for (int i = 0; i < n; ++i) out[i] = f_mul_add_loose(make_lop(a[i].re, a[i].im), make_rop(b[i].re, b[i].im), c[i]);
}
void h_reduce_ut(const u64* u, const u64* t, const u64* e, u64* out, int n) { for (int i = 0; i < n; ++i) out[i] = fp_reduce_ut(u[i], t[i], e[i]); }
void h_reduce_utw(const u64* u, const u64* t, const u64* w, const u64* e, u64* out, int n) {
    for (int i = 0; i < n; ++i) out[i] = fp_reduce_utw(u[i], t[i], w[i], e[i]);
}
void h_acc_dot(const F* a, const F* b, F* out, int n) {  // lazy Acc with a reduction every 8 products
    Acc acc = acc_zero(); F run = f_zero(); int pend = 0;
    for (int i = 0; i < n; ++i) { acc_mad(acc, a[i], b[i]); if (++pend == 8) { run = f_add(run, acc_reduce(acc)); acc = acc_zero(); pend = 0; } }
    *out = f_add(run, acc_reduce(acc));
}
}
extern "C" {
// loose running-sum variants: result must be in the right residue class and <= p + 5
void h_mul_add_k_loose(const F* a, const F* b, const F* c, F* out, int n) {
    for (int i = 0; i < n; ++i) out[i] = f_mul_add_k_loose(make_lop(a[i].re, a[i].im), make_ropd(b[i]), c[i]);
}
void h_mul_add_loose2(const F* a, const F* b, const F* c, F* out, int n) {
    for (int i = 0; i < n; ++i) out[i] = f_mul_add_loose2(make_lop(a[i].re, a[i].im), make_rop(b[i].re, b[i].im), c[i]);
}
}
extern "C" {
// weakly canonical primitives of the pass kernel: operands in [0,p] (p = alias of 0)
void h_fold_w(const F* v0, const F* v1, const F* r, F* out, int n) {
    for (int i = 0; i < n; ++i) out[i] = f_fold_w(v0[i], f_diff2p(v0[i], v1[i]), make_constk(r[i]));
}
void h_fold_w_dw(const F* v0, const F* v1, const F* r, F* out, int n) {   // with the difference folded into [0,p]
    for (int i = 0; i < n; ++i) out[i] = f_fold_w(v0[i], f_diff_w(v0[i], v1[i]), make_constk(r[i]));
}
void h_fold_w_real(const u64* v0, const u64* v1, const F* r, F* out, int n) {
    for (int i = 0; i < n; ++i) out[i] = f_fold_w_real(v0[i], fp_weak(v1[i] + P - v0[i]), make_constk(r[i]));
}
void h_reduce_ut_weak(const u64* u, const u64* t, const u64* e, u64* out, int n) { for (int i = 0; i < n; ++i) out[i] = fp_reduce_ut_weak(u[i], t[i], e[i]); }
void h_fp_weak(const u64* x, u64* out, int n) { for (int i = 0; i < n; ++i) out[i] = fp_weak(x[i]); }
// lazy dot products: sum m1*v + the pair form the kernel uses, sum (m1 - m0) * (v1 - v0)
void h_cacc_dot(const F* m, const F* v, F* out, int n) {
    CAcc s = cacc_zero();
    for (int i = 0; i < n; ++i) cacc_mad(s, make_lop(m[i].re, m[i].im), make_ropd(v[i]));
    *out = cacc_reduce(s);
}
void h_cacc_dot_diff(const F* m0, const F* m1, const F* v0, const F* v1, F* out, int n) {
    CAcc s = cacc_zero();
    for (int i = 0; i < n; ++i) {
        const F dm = f_diff2p(m0[i], m1[i]), dv = f_diff_w(v0[i], v1[i]);
        cacc_mad(s, make_lop(dm.re, dm.im), make_ropd(dv));
    }
    *out = cacc_reduce(s);
}
void h_cacc_dot_real(const F* m0, const F* m1, const u64* v, F* out, int n) {   // sum (m1 - m0 in [0,2p]) * v
    CAcc s = cacc_zero();
    for (int i = 0; i < n; ++i) cacc_mad_real(s, f_diff2p(m0[i], m1[i]), v[i]);
    *out = cacc_reduce(s);
}
}
