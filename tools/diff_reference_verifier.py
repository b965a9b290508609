"""Live differential run of the restated VERIFIER (oracle/gkr_oracle.c::ogkr_verify, which the device verifier vp_verify is
compared with) against the UNMODIFIED reference verifier (oracle/_ref/ref_dump with REF_TAMPER=k: the k-th prover message
reaches the stock verifier.cpp with 1 added): for every message index k of the golden small circuits, the stock verifier's
verdict -- which check fails and at which layer, read from its own stderr lines (verifier.cpp:164,211,251,299,329,385) -- must
be the (accept, code, layer) the oracle returns for the same tampered transcript.  CPU container only.
  python tools/diff_reference_verifier.py [STRIDE] [--write-golden]     (every STRIDE-th message index; default 1 = all)
--write-golden stores the stock verifier's verdicts in tests/golden/verifier_verdicts.json: {circuit: {message index: [accept,
code, layer]}} -- the fixture the CPU and GPU suites check the oracle verifier and vp_verify against where /root/reference is
not available."""
import lzma, os, re, subprocess, sys, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as E
B, O = E.binding(), E.oracle()
REF_DUMP = os.path.join(ROOT, "oracle", "_ref", "ref_dump")
PATTERNS = [(re.compile(r"Verification fail, phase1, circuit (\d+),"), 1), (re.compile(r"Verification fail, phase2, circuit level (\d+),"), 2),
            (re.compile(r"Verification fail, semi final, circuit level (\d+)"), 3), (re.compile(r"Liu fail, circuit (\d+), current bit"), 4),
            (re.compile(r"Liu fail, semi final, circuit (\d+)"), 5), (re.compile(r"Verification fail, final input check fail"), 6)]


def reference_verdict(pws_path, k, td):
    r = subprocess.run([REF_DUMP, pws_path, os.path.join(td, "t")], capture_output=True, text=True, env=dict(os.environ, REF_TAMPER=str(k)))
    if "VERIFY 1" in r.stdout:
        return (True, 0, 0)
    for line in r.stderr.split("\n"):
        for pat, code in PATTERNS:
            m = pat.search(line)
            if m:
                return (False, code, int(m.group(1)) if m.groups() else 0)
    return (False, -1, -1)


GOLD = {}


def run(name, stride):
    pws = lzma.open(os.path.join(ROOT, "tests", "golden", name + ".pws.xz")).read()
    circ = B.Circuit.from_pws_text(pws)
    oc = O.OracleCircuit(circ.flat())
    tr, _, _ = oc.prove()
    n = bad = 0
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "c.pws")
        open(path, "wb").write(pws)
        for k in range(0, len(tr), stride):
            t = tr.copy()
            t[k]["re"] = (int(t[k]["re"]) + 1) % B.P
            want, got = reference_verdict(path, k, td), tuple(oc.verify(t))
            n += 1
            GOLD.setdefault(name, {})[str(k)] = [int(want[0]), want[1], want[2]]
            if want != got:
                bad += 1
                print(f"MISMATCH {name} message {k}: reference verifier {want}, oracle verifier {got}")
    return n, bad


if __name__ == "__main__":
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    stride = int(args[0]) if args else 1
    tot = totbad = 0
    for name in ["small_allops", "small_chain", "small_notquirk", "small_random_a", "small_random_b", "small_random_c"]:
        n, bad = run(name, stride)
        print(name, n, "tampered messages,", bad, "mismatches", flush=True)
        tot += n; totbad += bad
    print("verifier cases", tot, "mismatches", totbad)
    if "--write-golden" in sys.argv:
        import json
        with open(os.path.join(ROOT, "tests", "golden", "verifier_verdicts.json"), "w") as f:
            json.dump(GOLD, f, sort_keys=True, separators=(",", ":"))
            f.write("\n")
