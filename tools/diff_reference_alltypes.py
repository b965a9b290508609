"""Live differential run, for the gate types / fields the .pws parser never produces (Addc, Mulc with real and complex constants,
Copy, Not of any wire, assert gates), of this repo's C oracle -- prover AND verifier -- against the UNMODIFIED reference
(oracle/_ref/ref_dump in its in-memory mode: stock prover + stock verifier incl. the polynomial commitment, proxy in between):
  * honest run: the stock verifier accepts, and every message it saw equals the oracle prover's transcript;
  * REF_TAMPER=k for every STRIDE-th message: the stock verifier's verdict (failing check, layer) == the oracle verifier's.
CPU container only.   python tools/diff_reference_alltypes.py [N_CIRCUITS] [STRIDE] [--write-golden]
--write-golden stores circuits + verdicts in tests/golden/verifier_verdicts_alltypes.json.xz for the CPU / GPU suites."""
import json, os, struct, subprocess, sys, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as E
import helpers as H
from diff_reference_verifier import PATTERNS, REF_DUMP   # noqa: E402
B, O = E.binding(), E.oracle()
_MUL, _ADD, _SUB, _ANTISUB, _NAAB, _ANTINAAB, _INPUT, _MULC, _ADDC, _XOR, _NOT, _COPY = range(12)   # inputCircuit.hpp:14-16


def make_arrays(seed, complex_consts, with_assert):
    """layered circuit over all gate types; >= 256 inputs (the reference's commitment needs bitLength(0) >= 8), every layer >= 2
    gates (a 1-gate layer makes the reference write out of bounds, prover.cpp:496)"""
    rng = np.random.default_rng(seed)
    n_layers = int(rng.integers(3, 7))
    sizes = [int(rng.integers(256, 300))] + [int(rng.integers(2, 40)) for _ in range(n_layers - 1)]
    ty, l, u, v, c, asr = [], [], [], [], [], []
    for g in range(sizes[0]):
        ty.append(_INPUT); l.append(-1); u.append(int(rng.integers(0, 1 << 31))); v.append(0); c.append((0, 0)); asr.append(0)
    kinds = [_MUL, _ADD, _SUB, _ANTISUB, _NAAB, _ANTINAAB, _MULC, _ADDC, _XOR, _NOT, _COPY]
    for i in range(1, n_layers):
        for g in range(sizes[i]):
            t = kinds[int(rng.integers(0, len(kinds)))]
            uu = int(rng.integers(0, sizes[i - 1]))
            cc, a = (0, 0), 0
            if t in (_MULC, _ADDC):
                cc = (int(rng.integers(1, B.P)), int(rng.integers(0, B.P)) if complex_consts else 0)
            if t in (_NOT, _COPY, _MULC, _ADDC):
                ll, vv = -1, 0
            else:
                ll = int(rng.integers(0, i))
                vv = int(rng.integers(0, sizes[ll]))
            if with_assert and g == 0:   # Sub(x, x) == 0: a legal assert gate
                t, ll, vv, cc, a = _SUB, i - 1, uu, (0, 0), 1
            ty.append(t); l.append(ll); u.append(uu); v.append(vv); c.append(cc); asr.append(a)
    return dict(sizes=sizes, ty=ty, l=l, u=u, v=v, c=c, is_assert=asr)


def circuit_of(a):
    cst = np.zeros(len(a["c"]), B.F_DTYPE)
    cst["re"] = [x[0] for x in a["c"]]
    cst["im"] = [x[1] for x in a["c"]]
    return B.Circuit.from_arrays(a["sizes"], a["ty"], a["l"], a["u"], a["v"], c=cst, is_assert=a["is_assert"] if any(a["is_assert"]) else None)


def mem_bytes(a):
    out = [struct.pack("<i", len(a["sizes"]))]
    k = 0
    for sz in a["sizes"]:
        out.append(struct.pack("<Q", sz))
        for _ in range(sz):
            out.append(struct.pack("<BiQQQQB", a["ty"][k], a["l"][k], a["u"][k], a["v"][k], a["c"][k][0], a["c"][k][1], a["is_assert"][k]))
            k += 1
    return b"".join(out)


def reference_run(path, td, k=None):
    env = dict(os.environ)
    if k is not None:
        env["REF_TAMPER"] = str(k)
    r = subprocess.run([REF_DUMP, path, os.path.join(td, "t")], capture_output=True, text=True, env=env)
    if "VERIFY 1" in r.stdout:
        return (True, 0, 0), open(os.path.join(td, "t.transcript.txt")).read()
    for line in r.stderr.split("\n"):
        for pat, code in PATTERNS:
            m = pat.search(line)
            if m:
                return (False, code, int(m.group(1)) if m.groups() else 0), None
    return (False, -1, -1), None


def main():
    args = [x for x in sys.argv[1:] if not x.startswith("--")]
    n_circ = int(args[0]) if args else 8
    stride = int(args[1]) if len(args) > 1 else 5
    gold, tot, bad = {}, 0, 0
    for seed in range(1, n_circ + 1):
        cc, wa = bool(seed & 1), bool(seed & 2)
        a = make_arrays(seed, cc, wa)
        circ = circuit_of(a)
        oc = O.OracleCircuit(circ.flat())
        tr, ch, _ = oc.prove()
        with tempfile.TemporaryDirectory() as td:
            path = os.path.join(td, "c.mem")
            open(path, "wb").write(mem_bytes(a))
            verdict, text = reference_run(path, td)
            honest = verdict == (True, 0, 0) and text == H.transcript_text(circ, tr, ch)
            tot += 1
            if not honest:
                bad += 1
                print(f"MISMATCH seed {seed} (complex {cc}, assert {wa}): honest run: reference verdict {verdict}, transcript equal {text == H.transcript_text(circ, tr, ch) if text else None}")
                continue
            cases = {}
            for k in range(0, len(tr), stride):
                t = tr.copy()
                t[k]["re"] = (int(t[k]["re"]) + 1) % B.P
                want, _ = reference_run(path, td, k)
                got = tuple(oc.verify(t))
                cases[str(k)] = [int(want[0]), want[1], want[2]]
                tot += 1
                if want != got:
                    bad += 1
                    print(f"MISMATCH seed {seed} message {k}: reference verifier {want}, oracle verifier {got}")
        import hashlib
        gold[str(seed)] = {"arrays": a, "complex_consts": cc, "with_assert": wa, "transcript_len": int(len(tr)), "verdicts": cases,
                           # the messages the stock verifier saw in the honest run (== the oracle's transcript, checked above)
                           "transcript_sha256": hashlib.sha256(np.ascontiguousarray(tr).tobytes()).hexdigest()}
        print(f"seed {seed}: layers {a['sizes']}, complex {cc}, assert {wa}: honest ok, {len(cases)} tampered messages", flush=True)
    print("all-type cases", tot, "mismatches", bad)
    if "--write-golden" in sys.argv:
        import lzma
        with lzma.open(os.path.join(ROOT, "tests", "golden", "verifier_verdicts_alltypes.json.xz"), "wt", preset=9) as f:
            json.dump(gold, f, sort_keys=True, separators=(",", ":"))
            f.write("\n")


if __name__ == "__main__":
    main()
