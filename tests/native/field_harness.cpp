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
