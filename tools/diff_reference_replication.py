"""Live differential run of template x K replication (vp_circuit_replicate + vp_circuit_expand: the index rules of DESIGN.md 3)
against the UNMODIFIED reference's own layering of the K-fold .pws (inputs first, instance-major: SURVEY 9.3) on seeded random
circuits, K in {2, 3, 5}: circuit dump after subsetInit and transcript, byte for byte. CPU container only.
Round 2: 40 cases, 0 mismatches."""
import importlib.util, os, random, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)
import helpers as H
import __graft_entry__ as E
B, O = E.binding(), E.oracle()
spec = importlib.util.spec_from_file_location("mg", os.path.join(ROOT, "tests/golden/make_golden.py")); mg = importlib.util.module_from_spec(spec); spec.loader.exec_module(mg)
n = bad = skipped = 0
for s in range(1, 41):
    rng = random.Random(s * 31337)
    n_in = rng.choice([70, 100, 128, 200]); n_g = rng.choice([60, 150, 400]); K = rng.choice([2, 3, 5])
    pws = mg.random_pws(s, n_in, n_g)
    big = mg.replicate_pws(pws, K)
    with tempfile.TemporaryDirectory() as td:
        p = os.path.join(td, "c.pws"); open(p, "wb").write(big)
        r = subprocess.run([mg.REF_DUMP, p, os.path.join(td, "c")], capture_output=True, text=True)
        if "VERIFY 1" not in r.stdout: skipped += 1; continue
        cb = open(os.path.join(td, "c.circuit.bin"), "rb").read(); trr = open(os.path.join(td, "c.transcript.txt")).read()
    circ = B.Circuit.from_pws_text(pws).replicate(K)
    ex = circ.expand()
    n += 1
    ok1 = H.circuit_dump(ex) == cb
    tr, ch, _ = O.OracleCircuit(ex.flat()).prove()
    ok2 = H.transcript_text(ex, tr, ch) == trr
    if not (ok1 and ok2): bad += 1; print("MISMATCH seed", s, "K", K, ok1, ok2)
print("replication cases", n, "skipped (reference aborts)", skipped, "mismatches", bad)
