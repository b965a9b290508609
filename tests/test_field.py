"""CPU tests of the product's field arithmetic (virgo-plus_b200/csrc/field.cuh compiled for the host:
the very same inline integer code nvcc compiles for sm_100a) against Python big integers."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = (1 << 61) - 1
FD = np.dtype([("re", "<u8"), ("im", "<u8")])


@pytest.fixture(scope="module")
def H(tmp_path_factory):
    so = tmp_path_factory.mktemp("fh") / "field_harness.so"
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-x", "c++",
                           os.path.join(ROOT, "tests", "native", "field_harness.cpp"), "-o", str(so)])
    return C.CDLL(str(so))


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


EDGE = [0, 1, 2, 3, P - 1, P - 2, P - 3, (1 << 60), (1 << 60) - 1, (1 << 60) + 1, (1 << 31) - 1, 1 << 31, (1 << 31) + 1,
        (1 << 32) - 1, 1 << 32, (1 << 30), (1 << 30) - 1, 0x7FFFFFFF7FFFFFFF % P, (1 << 61) - (1 << 31), (1 << 61) - (1 << 31) - 1]


def _operands(rng, n_rand, hi=P):
    vals = list(EDGE) + [int(x) for x in rng.integers(0, hi, n_rand, dtype=np.uint64)]
    return [v % hi if hi == P else v for v in vals]


def _fe_arrays(rng, n):
    base = _operands(rng, 40)
    re = rng.choice(base, n)
    im = rng.choice(base, n)
    a = np.zeros(n, FD)
    a["re"] = re.astype(np.uint64)
    a["im"] = im.astype(np.uint64)
    return a


def _cmul(a, b):
    return ((a[0] * b[0] - a[1] * b[1]) % P, (a[0] * b[1] + a[1] * b[0]) % P)


def _t(x):
    return (int(x["re"]), int(x["im"]))


def test_mul_add_sub_fold(H):
    rng = np.random.default_rng(7)
    n = 20000
    a, b, c = _fe_arrays(rng, n), _fe_arrays(rng, n), _fe_arrays(rng, n)
    out = np.zeros(n, FD)
    H.h_mul(_p(a), _p(b), _p(out), n)
    for i in range(n):
        assert _t(out[i]) == _cmul(_t(a[i]), _t(b[i])), (i, a[i], b[i])
    H.h_mul_add(_p(a), _p(b), _p(c), _p(out), n)
    for i in range(0, n, 3):
        m = _cmul(_t(a[i]), _t(b[i]))
        assert _t(out[i]) == ((m[0] + int(c[i]["re"])) % P, (m[1] + int(c[i]["im"])) % P)
    H.h_fold(_p(a), _p(b), _p(c), _p(out), n)  # a + c*(b - a)
    for i in range(0, n, 3):
        d = ((int(b[i]["re"]) - int(a[i]["re"])) % P, (int(b[i]["im"]) - int(a[i]["im"])) % P)
        m = _cmul(d, _t(c[i]))
        assert _t(out[i]) == ((m[0] + int(a[i]["re"])) % P, (m[1] + int(a[i]["im"])) % P), i
    H.h_add(_p(a), _p(b), _p(out), n)
    assert all(_t(out[i]) == ((int(a[i]["re"]) + int(b[i]["re"])) % P, (int(a[i]["im"]) + int(b[i]["im"])) % P) for i in range(0, n, 5))
    H.h_sub(_p(a), _p(b), _p(out), n)
    assert all(_t(out[i]) == ((int(a[i]["re"]) - int(b[i]["re"])) % P, (int(a[i]["im"]) - int(b[i]["im"])) % P) for i in range(0, n, 5))
    assert (out["re"] < P).all() and (out["im"] < P).all()


def test_all_edge_pairs_exhaustive(H):
    """every (re, im) x (re, im) combination of the edge values"""
    e = np.array(EDGE, dtype=np.uint64)
    g = np.array(np.meshgrid(e, e, e, e, indexing="ij")).reshape(4, -1)
    n = g.shape[1]
    a, b = np.zeros(n, FD), np.zeros(n, FD)
    a["re"], a["im"], b["re"], b["im"] = g
    out = np.zeros(n, FD)
    H.h_mul(_p(a), _p(b), _p(out), n)
    ar, ai, br, bi = (x.astype(object) for x in g)
    want_re = (ar * br - ai * bi) % P
    want_im = (ar * bi + ai * br) % P
    assert (out["re"].astype(object) == want_re).all() and (out["im"].astype(object) == want_im).all()


def test_loose_operands_up_to_2p(H):
    rng = np.random.default_rng(8)
    n = 20000
    loose_edge = [0, 1, P - 1, P, P + 1, 2 * P - 1, 2 * P, (1 << 61), (1 << 62) - 3, (1 << 31) * ((1 << 31) - 1)]
    def arr():
        vals = loose_edge + [int(x) for x in rng.integers(0, 2 * P + 1, 50, dtype=np.uint64)]
        a = np.zeros(n, FD)
        a["re"] = rng.choice(vals, n).astype(np.uint64)
        a["im"] = rng.choice(vals, n).astype(np.uint64)
        return a
    a, b, c = arr(), arr(), _fe_arrays(rng, n)
    out = np.zeros(n, FD)
    H.h_mul_add_loose(_p(a), _p(b), _p(c), _p(out), n)
    for i in range(n):
        m = _cmul(_t(a[i]), _t(b[i]))
        assert _t(out[i]) == ((m[0] + int(c[i]["re"])) % P, (m[1] + int(c[i]["im"])) % P), (i, a[i], b[i])


def test_fp_mul_and_reductions(H):
    rng = np.random.default_rng(9)
    n = 20000
    big = [0, 1, (1 << 64) - 1, (1 << 63), (1 << 63) - 1, (1 << 61), (1 << 61) - 1, (1 << 62), 0xFFFFFFFF00000000, 0x3FFFFFFF, 0x40000000]
    def u64s(hi):
        v = [x for x in big if x < hi] + [int(x) for x in rng.integers(0, min(hi, (1 << 64) - 1), 60, dtype=np.uint64)]
        return rng.choice(np.array(v, dtype=np.uint64), n)
    a, b = u64s(1 << 62), u64s(1 << 62)
    out = np.zeros(n, np.uint64)
    H.h_fp_mul(_p(a), _p(b), _p(out), n)
    assert all(int(out[i]) == int(a[i]) * int(b[i]) % P for i in range(n))
    u, t, w, e = u64s(1 << 64), u64s(1 << 64), u64s(1 << 64), u64s(1 << 62)
    H.h_reduce_ut(_p(u), _p(t), _p(e), _p(out), n)
    assert all(int(out[i]) == (int(u[i]) + (int(t[i]) << 31) + int(e[i])) % P for i in range(n))
    H.h_reduce_utw(_p(u), _p(t), _p(w), _p(e), _p(out), n)
    assert all(int(out[i]) == (int(u[i]) + (int(t[i]) << 31) + (int(w[i]) << 62) + int(e[i])) % P for i in range(n))


def test_lazy_accumulator_dot(H):
    rng = np.random.default_rng(10)
    for n in [1, 7, 8, 9, 100]:
        a, b = _fe_arrays(rng, n), _fe_arrays(rng, n)
        a["re"][0] = a["im"][0] = b["re"][0] = b["im"][0] = P - 1
        out = np.zeros(1, FD)
        H.h_acc_dot(_p(a), _p(b), _p(out), n)
        acc = (0, 0)
        for i in range(n):
            m = _cmul(_t(a[i]), _t(b[i]))
            acc = ((acc[0] + m[0]) % P, (acc[1] + m[1]) % P)
        assert _t(out[0]) == acc


def test_loose_running_sums(H):
    """f_mul_add_k_loose / f_mul_add_loose2: right residue class, bounded by p + 5, and stable when fed back"""
    rng = np.random.default_rng(11)
    n = 5000
    a, b = _fe_arrays(rng, n), _fe_arrays(rng, n)
    acc = np.zeros(n, FD)
    want = [(0, 0)] * n
    for it in range(6):   # feed the loose result back as the addend, like the round kernel does
        out = np.zeros(n, FD)
        fn = H.h_mul_add_k_loose if it % 2 == 0 else H.h_mul_add_loose2
        fn(_p(a), _p(b), _p(acc), _p(out), n)
        for i in range(0, n, 7):
            m = _cmul(_t(a[i]), _t(b[i]))
            want[i] = ((want[i][0] + m[0]) % P, (want[i][1] + m[1]) % P)
            assert int(out[i]["re"]) % P == want[i][0] and int(out[i]["im"]) % P == want[i][1]
        assert int(out["re"].max()) <= P + 5 and int(out["im"].max()) <= P + 5
        acc = out
        a = np.roll(a, 1)


# ---------------------------------------------------------------- weakly canonical / lazy primitives of k_phase_dfs
WEAK_EDGE = EDGE + [P]   # p itself is a legal (weak) representation of 0


def _weak_arrays(rng, n):
    base = WEAK_EDGE + [int(x) for x in rng.integers(0, P, 40, dtype=np.uint64)]
    a = np.zeros(n, FD)
    a["re"] = rng.choice(base, n).astype(np.uint64)
    a["im"] = rng.choice(base, n).astype(np.uint64)
    return a


def test_weak_fold(H):
    rng = np.random.default_rng(21)
    n = 30000
    v0, v1, r = _weak_arrays(rng, n), _weak_arrays(rng, n), _fe_arrays(rng, n)
    for fn in (H.h_fold_w, H.h_fold_w_dw):
        out = np.zeros(n, FD)
        fn(_p(v0), _p(v1), _p(r), _p(out), n)
        assert int(out["re"].max()) <= P and int(out["im"].max()) <= P
        for i in range(0, n, 3):
            d = ((int(v1[i]["re"]) - int(v0[i]["re"])) % P, (int(v1[i]["im"]) - int(v0[i]["im"])) % P)
            m = _cmul(d, _t(r[i]))
            want = ((int(v0[i]["re"]) + m[0]) % P, (int(v0[i]["im"]) + m[1]) % P)
            assert (int(out[i]["re"]) % P, int(out[i]["im"]) % P) == want
    # base-field data
    a0 = v0["re"].copy(); a1 = v1["re"].copy()
    out = np.zeros(n, FD)
    H.h_fold_w_real(_p(a0), _p(a1), _p(r), _p(out), n)
    assert int(out["re"].max()) <= P and int(out["im"].max()) <= P
    for i in range(0, n, 3):
        d = (int(a1[i]) - int(a0[i])) % P
        want = ((int(a0[i]) + d * int(r[i]["re"])) % P, (d * int(r[i]["im"])) % P)
        assert (int(out[i]["re"]) % P, int(out[i]["im"]) % P) == want


def test_weak_reductions(H):
    rng = np.random.default_rng(22)
    big = [0, 1, P, P + 1, 2 * P, 2 * P + 1, (1 << 64) - 1, (1 << 64) - 2, (1 << 63), (1 << 62) - 1, (1 << 61), 3 * P, 7 * P]
    vals = big + [int(x) for x in rng.integers(0, 1 << 64, 200, dtype=np.uint64)]
    u = np.array([a for a in vals for _ in vals][:20000], dtype=np.uint64)
    t = np.array([b for _ in vals for b in vals][:20000], dtype=np.uint64)
    e = np.array([vals[k] for k in rng.integers(0, len(vals), len(u))], dtype=np.uint64)
    out = np.zeros(len(u), np.uint64)
    H.h_reduce_ut_weak(_p(u), _p(t), _p(e), _p(out), len(u))
    assert int(out.max()) <= P
    for i in range(0, len(u), 5):
        assert int(out[i]) % P == (int(u[i]) + (int(t[i]) << 31) + int(e[i])) % P
    x = np.array(vals, dtype=np.uint64)
    out = np.zeros(len(x), np.uint64)
    H.h_fp_weak(_p(x), _p(out), len(x))
    assert int(out.max()) <= P
    assert [int(o) % P for o in out] == [v % P for v in vals]


def test_lazy_complex_accumulators(H):
    rng = np.random.default_rng(23)
    for n in [1, 2, 9, 300, 5000]:
        m0, m1, v0, v1 = (_weak_arrays(rng, n) for _ in range(4))
        for k in range(min(n, 4)):   # worst-case magnitudes first
            for a in (m0, v0):
                a["re"][k] = a["im"][k] = 0
            for a in (m1, v1):
                a["re"][k] = a["im"][k] = P
        out = np.zeros(1, FD)
        H.h_cacc_dot(_p(m1), _p(v1), _p(out), n)
        acc = [0, 0]
        for i in range(n):
            m = _cmul(_t(m1[i]), _t(v1[i]))
            acc = [(acc[0] + m[0]) % P, (acc[1] + m[1]) % P]
        assert _t(out[0]) == tuple(acc)
        H.h_cacc_dot_diff(_p(m0), _p(m1), _p(v0), _p(v1), _p(out), n)
        acc = [0, 0]
        for i in range(n):
            dm = ((int(m1[i]["re"]) - int(m0[i]["re"])) % P, (int(m1[i]["im"]) - int(m0[i]["im"])) % P)
            dv = ((int(v1[i]["re"]) - int(v0[i]["re"])) % P, (int(v1[i]["im"]) - int(v0[i]["im"])) % P)
            m = _cmul(dm, dv)
            acc = [(acc[0] + m[0]) % P, (acc[1] + m[1]) % P]
        assert _t(out[0]) == tuple(acc)
        vr = v1["re"].copy()
        H.h_cacc_dot_real(_p(m0), _p(m1), _p(vr), _p(out), n)
        acc = [0, 0]
        for i in range(n):
            dm = ((int(m1[i]["re"]) - int(m0[i]["re"])) % P, (int(m1[i]["im"]) - int(m0[i]["im"])) % P)
            acc = [(acc[0] + dm[0] * int(vr[i])) % P, (acc[1] + dm[1] * int(vr[i])) % P]
        assert _t(out[0]) == tuple(acc)
