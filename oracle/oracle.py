"""TEST INFRASTRUCTURE ONLY -- ctypes wrapper of oracle/libgkr_oracle.so (the CPU restatement in
gkr_oracle.c). Imported only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs; never by the product package."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libgkr_oracle.so")
F_DTYPE = np.dtype([("re", "<u8"), ("im", "<u8")])


class _Circuit(C.Structure):
    _fields_ = [("n_layers", C.c_int32)] + [(k, C.c_void_p) for k in (
        "layer_size", "gate_off", "ty", "l", "u", "v", "lv", "c", "is_assert", "dad_size", "dad_off", "dad_id", "inputs")]


_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", HERE, "oracle"])


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        L = C.CDLL(LIB_PATH)
        L.ogkr_transcript_len.restype = C.c_size_t
        L.ogkr_transcript_len.argtypes = [C.c_void_p]
        L.ogkr_challenge_count.restype = C.c_size_t
        L.ogkr_challenge_count.argtypes = [C.c_void_p]
        L.ogkr_prove.argtypes = [C.c_void_p, C.c_uint, C.c_void_p, C.c_void_p, C.POINTER(C.c_double)]
        L.ogkr_verify.argtypes = [C.c_void_p, C.c_uint, C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.ogkr_evaluate.argtypes = [C.c_void_p, C.c_void_p]
        L.ogkr_evaluate.restype = None
        L.ogkr_sumcheck_tables.argtypes = [C.c_void_p] * 3 + [C.c_int, C.c_void_p, C.c_void_p]
        L.ogkr_sumcheck_tables.restype = None
        L.ogkr_beta_table.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_uint64, C.c_uint64]
        L.ogkr_beta_table.restype = None
        for fn in (L.ofe_add, L.ofe_sub, L.ofe_mul):
            fn.argtypes = [C.c_uint64] * 4
            fn.restype = _Fe
        L.ogkr_seed.argtypes = [C.c_uint]
        L.ogkr_seed.restype = None
        L.ogkr_random_field.restype = _Fe
        _lib = L
    return _lib


class _Fe(C.Structure):
    _fields_ = [("re", C.c_uint64), ("im", C.c_uint64)]


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class OracleCircuit:
    """Wraps the dict produced by binding.Circuit.flat() as an ogkr_circuit."""

    def __init__(self, flat):
        self.flat = {k: (np.ascontiguousarray(v) if isinstance(v, np.ndarray) else v) for k, v in flat.items()}
        f = self.flat
        any_c = bool(np.any(f["c"]["re"] | f["c"]["im"]))
        any_a = bool(np.any(f["is_assert"]))
        self.s = _Circuit(
            f["n_layers"], _p(f["layer_size"]), _p(f["gate_off"]), _p(f["ty"]), _p(f["l"]), _p(f["u"]), _p(f["v"]),
            _p(f["lv"]), _p(f["c"]) if any_c else None, _p(f["is_assert"]) if any_a else None, _p(f["dad_size"]),
            _p(f["dad_off"]), _p(f["dad_id"]), _p(f["inputs"]))

    @property
    def ref(self):
        return C.byref(self.s)

    @property
    def transcript_len(self):
        return lib().ogkr_transcript_len(self.ref)

    @property
    def challenge_count(self):
        return lib().ogkr_challenge_count(self.ref)

    def prove(self, seed=3396):
        """-> (transcript, challenges, prove_seconds); raises if an assert gate is violated."""
        tr = np.zeros(self.transcript_len, F_DTYPE)
        ch = np.zeros(self.challenge_count, F_DTYPE)
        sec = C.c_double()
        rc = lib().ogkr_prove(self.ref, seed, _p(tr), _p(ch), C.byref(sec))
        if rc != 0:
            raise RuntimeError("oracle: assert gate violated")
        return tr, ch, sec.value

    def verify(self, transcript, seed=3396):
        """-> (accept: bool, fail_code, fail_layer)"""
        tr = np.ascontiguousarray(transcript, dtype=F_DTYPE)
        assert len(tr) == self.transcript_len
        code, layer = C.c_int(), C.c_int()
        ok = lib().ogkr_verify(self.ref, seed, _p(tr), C.byref(code), C.byref(layer))
        return bool(ok), code.value, layer.value

    # ---- Fiat-Shamir mode (transcriptCache restated; order of virgo-plus_b200/host/fiat_shamir.h)
    def prove_fs(self, seed32):
        """-> (transcript, challenges in the usual layout)"""
        L = lib()
        L.ogkr_prove_fs.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_void_p]
        tr = np.zeros(self.transcript_len, F_DTYPE)
        ch = np.zeros(self.challenge_count, F_DTYPE)
        rc = L.ogkr_prove_fs(self.ref, bytes(seed32), _p(tr), _p(ch))
        if rc != 0:
            raise RuntimeError("oracle: assert gate violated")
        return tr, ch

    def fs_challenges(self, seed32, transcript):
        L = lib()
        L.ogkr_fs_challenges.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_void_p]
        L.ogkr_fs_challenges.restype = None
        tr = np.ascontiguousarray(transcript, dtype=F_DTYPE)
        ch = np.zeros(self.challenge_count, F_DTYPE)
        L.ogkr_fs_challenges(self.ref, bytes(seed32), _p(tr), _p(ch))
        return ch

    def verify_fs(self, seed32, transcript):
        L = lib()
        L.ogkr_verify_fs.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        tr = np.ascontiguousarray(transcript, dtype=F_DTYPE)
        code, layer = C.c_int(), C.c_int()
        ok = L.ogkr_verify_fs(self.ref, bytes(seed32), _p(tr), C.byref(code), C.byref(layer))
        return bool(ok), code.value, layer.value

    def evaluate(self):
        n = int(self.flat["layer_size"].sum())
        out = np.zeros(n, F_DTYPE)
        lib().ogkr_evaluate(self.ref, _p(out))
        return out


def sumcheck_tables(V, add, mult, r):
    V, add, mult, r = [np.ascontiguousarray(x, dtype=F_DTYPE) for x in (V, add, mult, r)]
    log_n = len(r)
    assert len(V) == 1 << log_n
    out = np.zeros(3 * log_n + 3, F_DTYPE)
    lib().ogkr_sumcheck_tables(_p(V), _p(add), _p(mult), log_n, _p(r), _p(out))
    return out


def beta_table(r, init=(1, 0)):
    r = np.ascontiguousarray(r, dtype=F_DTYPE)
    out = np.zeros(1 << len(r), F_DTYPE)
    lib().ogkr_beta_table(_p(out), len(r), _p(r), init[0], init[1])
    return out


def _fe(x):
    return (int(x["re"]), int(x["im"])) if not isinstance(x, tuple) else x


def f_add(a, b):
    a, b = _fe(a), _fe(b)
    r = lib().ofe_add(a[0], a[1], b[0], b[1])
    return (r.re, r.im)


def f_sub(a, b):
    a, b = _fe(a), _fe(b)
    r = lib().ofe_sub(a[0], a[1], b[0], b[1])
    return (r.re, r.im)


def f_mul(a, b):
    a, b = _fe(a), _fe(b)
    r = lib().ofe_mul(a[0], a[1], b[0], b[1])
    return (r.re, r.im)


def draw_challenges(n, seed=3396):
    """n field elements from srandom(seed) + fieldElement::random() (global glibc state)."""
    L = lib()
    L.ogkr_seed(seed)
    out = np.zeros(n, F_DTYPE)
    for i in range(n):
        r = L.ogkr_random_field()
        out[i] = (r.re, r.im)
    return out


# ---------------------------------------------------------------- the compiled reference (oracle/_ref)
REF_DIR = os.path.join(HERE, "_ref")
REF_LIB = os.path.join(REF_DIR, "libref_gkr.so")
_ref = None


def ref_available():
    return os.path.exists(REF_LIB)


def ref_prove(flat):
    """Run the UNMODIFIED reference prover (libref_gkr.so) on a flat circuit dict.
    -> (transcript, prove_seconds (the reference's `Prove Time`), evaluate_seconds)"""
    global _ref
    if _ref is None:
        _ref = C.CDLL(REF_LIB)
        _ref.ref_gkr_prove.argtypes = [C.c_int] + [C.c_void_p] * 6 + [C.c_uint, C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    f = {k: (np.ascontiguousarray(v) if isinstance(v, np.ndarray) else v) for k, v in flat.items()}
    oc = OracleCircuit(flat)
    tr = np.zeros(oc.transcript_len, F_DTYPE)
    ps, es = C.c_double(), C.c_double()
    if np.any(f["c"]["re"] | f["c"]["im"]) or np.any(f["is_assert"]):   # gate constants / assert flags: the extended entry point
        _ref.ref_gkr_prove2.argtypes = [C.c_int] + [C.c_void_p] * 8 + [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        n = _ref.ref_gkr_prove2(f["n_layers"], _p(f["layer_size"]), _p(f["ty"]), _p(f["l"]), _p(f["u"]), _p(f["v"]),
                                _p(f["inputs"]), _p(f["c"]), _p(f["is_assert"]), _p(tr), C.byref(ps), C.byref(es))
    else:
        n = _ref.ref_gkr_prove(f["n_layers"], _p(f["layer_size"]), _p(f["ty"]), _p(f["l"]), _p(f["u"]), _p(f["v"]),
                               _p(f["inputs"]), 3396, _p(tr), C.byref(ps), C.byref(es))
    assert n == len(tr), (n, len(tr))
    return tr, ps.value, es.value


# ---------------------------------------------------------------- polynomial commitment, commit phase (pc_oracle.c)
def sha3_256(data):
    L = lib()
    out = (C.c_ubyte * 32)()
    buf = (C.c_ubyte * max(1, len(data))).from_buffer_copy(data if data else b"\0")
    L.opc_sha3_256.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
    L.opc_sha3_256.restype = None
    L.opc_sha3_256(buf, len(data), out)
    return bytes(out)


def pc_commit_private(array, log_len, mask=None, want_arrays=True):
    """poly_commit_prover::commit_private_array restated (poly_commit.h:41-124): -> dict(root, l_eval, leaf_hash, tree)"""
    L = lib()
    L.opc_commit_private.restype = C.c_long
    L.opc_commit_private.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    a = np.zeros(1 << log_len, F_DTYPE)
    a[:len(array)] = array
    m = np.zeros(1, F_DTYPE) if mask is None else np.ascontiguousarray(mask, dtype=F_DTYPE)
    ss = 1 << (log_len - 1)
    root = np.zeros(32, np.uint8)
    l_eval = np.zeros(65 * ss, F_DTYPE) if want_arrays else None
    leaf = np.zeros(ss // 2 * 32, np.uint8) if want_arrays else None
    tree = np.zeros(ss * 32, np.uint8) if want_arrays else None
    rc = L.opc_commit_private(_p(a), log_len, _p(m), len(m), _p(l_eval) if want_arrays else None,
                              _p(leaf) if want_arrays else None, _p(tree) if want_arrays else None, _p(root))
    assert rc == ss, rc
    return dict(root=root.tobytes(), l_eval=l_eval, leaf_hash=leaf, tree=tree, slice_size=ss)


def pc_commit_public(array, pub, log_len):
    """poly_commit_prover::commit_public_array restated (poly_commit.h:126-349, zero masks) -> dict(root_h, all_sum, h_eval, vow)"""
    L = lib()
    L.opc_commit_public.restype = C.c_long
    L.opc_commit_public.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    a = np.zeros(1 << log_len, F_DTYPE); a[:len(array)] = array
    q = np.zeros(1 << log_len, F_DTYPE); q[:len(pub)] = pub
    N = 1 << (log_len - 1)
    root = np.zeros(32, np.uint8)
    all_sum, h_eval, vow = np.zeros(65, F_DTYPE), np.zeros(65 * N, F_DTYPE), np.zeros(64 * N, F_DTYPE)
    rc = L.opc_commit_public(_p(a), _p(q), log_len, _p(all_sum), _p(h_eval), _p(vow), _p(root))
    assert rc == N, rc
    return dict(root_h=root.tobytes(), all_sum=all_sum, h_eval=h_eval, vow=vow, slice_size=N)


def pc_fri_commit_phase(vow, log_N, randomness):
    """fri::commit_phase_step restated (fri.cpp:289-418), one step per element of `randomness`, from the virtual oracle
    -> dict(roots [steps] bytes, codes [per level np array], trees [per level bytes])"""
    L = lib()
    L.opc_fri_commit_phase.restype = C.c_long
    L.opc_fri_commit_phase.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    N, steps = 1 << log_N, len(randomness)
    v = np.ascontiguousarray(vow, F_DTYPE)
    assert len(v) == 64 * N
    r = np.ascontiguousarray(randomness, F_DTYPE)
    sizes = [N >> (l + 1) for l in range(steps)]
    roots = np.zeros(32 * steps, np.uint8)
    codes = np.zeros(64 * sum(sizes), F_DTYPE)
    trees = np.zeros(32 * sum(sizes), np.uint8)
    rc = L.opc_fri_commit_phase(_p(v), log_N, _p(r), steps, _p(roots), _p(codes), _p(trees))
    assert rc == steps, rc
    out, o = dict(roots=[roots[32 * l:32 * l + 32].tobytes() for l in range(steps)], codes=[], trees=[]), 0
    for m in sizes:
        out["codes"].append(codes[64 * o:64 * (o + m)])
        out["trees"].append(trees[32 * o:32 * (o + m)].tobytes())
        o += m
    return out


REF_PC = os.path.join(REF_DIR, "ref_pc_commit")


def ref_pc_fri(array, pub, log_len, randomness):
    """the UNMODIFIED reference: commit_private_array, commit_public_array, then fri::commit_phase_step per element of
    `randomness` (log_len - 6 of them) -> dict(root_l, root_h, roots, codes, trees, seconds)"""
    import tempfile
    a = np.zeros(1 << log_len, F_DTYPE); a[:len(array)] = array
    q = np.zeros(1 << log_len, F_DTYPE); q[:len(pub)] = pub
    r = np.ascontiguousarray(randomness, F_DTYPE)
    steps, N = log_len - 6, 1 << (log_len - 1)
    assert len(r) == steps
    with tempfile.TemporaryDirectory() as td:
        fa, fq, fo, fo2, fr, fo3 = (os.path.join(td, x) for x in ("a.bin", "q.bin", "o.bin", "o2.bin", "r.bin", "o3.bin"))
        a.tofile(fa); q.tofile(fq); r.tofile(fr)
        pr = subprocess.run([REF_PC, str(log_len), fa, fo, fq, fo2, fr, fo3], capture_output=True, text=True, check=True)
        root_l = np.fromfile(fo, dtype=np.uint8, count=32).tobytes()
        root_h = np.fromfile(fo2, dtype=np.uint8, count=32).tobytes()
        raw = np.fromfile(fo3, dtype=np.uint8)
    out = dict(root_l=root_l, root_h=root_h, roots=[raw[32 * l:32 * l + 32].tobytes() for l in range(steps)], codes=[], trees=[],
               seconds=float(pr.stdout.split("fri_commit_seconds")[1]),
               commit_seconds=float(pr.stdout.split("commit_seconds")[1].split()[0]),
               commit_public_seconds=float(pr.stdout.split("commit_public_seconds")[1].split()[0]))
    o = 32 * steps
    for l in range(steps):
        m = N >> (l + 1)
        out["codes"].append(raw[o:o + 64 * m * 16].view(F_DTYPE)); o += 64 * m * 16
        out["trees"].append(raw[o:o + m * 32].tobytes()); o += m * 32
    return out


def ref_pc_commit_public(array, pub, log_len):
    """the UNMODIFIED reference commit_private_array + commit_public_array -> dict(root_h, all_sum, h_eval, vow, vow_msk, seconds)"""
    import tempfile
    a = np.zeros(1 << log_len, F_DTYPE); a[:len(array)] = array
    q = np.zeros(1 << log_len, F_DTYPE); q[:len(pub)] = pub
    N = 1 << (log_len - 1)
    with tempfile.TemporaryDirectory() as td:
        fa, fq, fo, fo2 = (os.path.join(td, x) for x in ("a.bin", "q.bin", "o.bin", "o2.bin"))
        a.tofile(fa); q.tofile(fq)
        r = subprocess.run([REF_PC, str(log_len), fa, fo, fq, fo2], capture_output=True, text=True, check=True)
        raw = np.fromfile(fo2, dtype=np.uint8)
    o = 32
    all_sum = raw[o:o + 65 * 16].view(F_DTYPE); o += 65 * 16
    h_eval = raw[o:o + 65 * N * 16].view(F_DTYPE); o += 65 * N * 16
    vow = raw[o:o + 64 * N * 16].view(F_DTYPE); o += 64 * N * 16
    vow_msk = raw[o:o + N * 16].view(F_DTYPE)
    return dict(root_h=raw[:32].tobytes(), all_sum=all_sum, h_eval=h_eval, vow=vow, vow_msk=vow_msk, slice_size=N,
                seconds=float(r.stdout.split("commit_public_seconds")[1]))


def ref_pc_commit(array, log_len):
    """the UNMODIFIED reference commit_private_array (oracle/_ref/ref_pc_commit) -> dict(root, l_eval, leaf_hash, tree, seconds)"""
    import tempfile
    a = np.zeros(1 << log_len, F_DTYPE)
    a[:len(array)] = array
    with tempfile.TemporaryDirectory() as td:
        fin, fout = os.path.join(td, "a.bin"), os.path.join(td, "o.bin")
        a.tofile(fin)
        r = subprocess.run([REF_PC, str(log_len), fin, fout], capture_output=True, text=True, check=True)
        ss = 1 << (log_len - 1)
        raw = np.fromfile(fout, dtype=np.uint8)
    o = 32
    l_eval = raw[o:o + 65 * ss * 16].view(F_DTYPE); o += 65 * ss * 16
    leaf = raw[o:o + ss // 2 * 32]; o += ss // 2 * 32
    tree = raw[o:o + ss * 32]
    return dict(root=raw[:32].tobytes(), l_eval=l_eval, leaf_hash=leaf, tree=tree, slice_size=ss,
                seconds=float(r.stdout.split("commit_seconds")[1]))


# ------------------------------------------------------------------ the polynomial commitment's inner GKR (fft_circuit_GKR)
def fft_gkr_rnd_count(lg):
    return lg + 64 + 2 * (lg + 10) + 2 * (lg + 6) + 2 * lg + lg * (2 * lg + 2)


def fft_gkr_poly_count(lg):
    return (lg + 6) + lg + 2 * lg * lg


def fft_gkr(lg, rnd, want_layers=True):
    """fft_circuit_GKR.cpp restated (fftgkr_oracle.c) -> dict(layers, polys [n, 3], claims, proof_size, ok)"""
    L = lib()
    L.ofg_run.restype = C.c_long
    L.ofg_run.argtypes = [C.c_int, C.c_void_p, C.c_long, C.c_void_p, C.c_void_p, C.c_long, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    rnd = np.ascontiguousarray(rnd, F_DTYPE)
    n = 1 << lg
    layers = np.zeros((lg + 2) * n + 64 * n + 64, F_DTYPE) if want_layers else None
    npoly = fft_gkr_poly_count(lg)
    polys = np.zeros(3 * npoly, F_DTYPE)
    claims = np.zeros(4 + lg + 2, F_DTYPE)
    n_polys, ps, ok = C.c_long(), C.c_int(), C.c_int()
    rc = L.ofg_run(lg, _p(rnd), len(rnd), _p(layers) if want_layers else None, _p(polys), npoly, _p(claims), C.byref(n_polys), C.byref(ps), C.byref(ok))
    assert rc == fft_gkr_rnd_count(lg), rc
    assert n_polys.value == npoly, (n_polys.value, npoly)
    return dict(layers=layers, polys=polys.reshape(npoly, 3), claims=claims, proof_size=ps.value, ok=bool(ok.value))


REF_FFTGKR = os.path.join(REF_DIR, "ref_fftgkr")


def ref_fft_gkr(lg, seed):
    """the UNMODIFIED reference fft_circuit_GKR driven stage by stage (oracle/_ref/ref_fftgkr)
    -> dict(rnd, layers, claims [a_0, add, mult, intermediate, ifft, alpha, beta], proof_size, ok, fft_gkr_ps, seconds)"""
    import tempfile
    n = 1 << lg
    with tempfile.TemporaryDirectory() as td:
        fo = os.path.join(td, "o.bin")
        pr = subprocess.run([REF_FFTGKR, str(lg), str(seed), fo], capture_output=True, text=True, check=True)
        raw = np.fromfile(fo, dtype=np.uint8)
    n_rnd = int(raw[:8].view(np.uint64)[0]); o = 8
    rnd = raw[o:o + 16 * n_rnd].view(F_DTYPE); o += 16 * n_rnd
    nl = (lg + 2) * n + 64 * n + 64
    layers = raw[o:o + 16 * nl].view(F_DTYPE); o += 16 * nl
    claims = raw[o:o + 16 * 7].view(F_DTYPE); o += 16 * 7
    ps, ok, ps2 = (int(x) for x in raw[o:o + 24].view(np.uint64))
    return dict(rnd=rnd, layers=layers, claims=claims, proof_size=ps, ok=bool(ok), fft_gkr_ps=ps2,
                seconds=float(pr.stdout.split("prover_seconds")[1].split()[0]))
