"""Differential campaign against the LIVE unmodified reference (oracle/_ref/ref_dump; only where /root/reference was built):
seeded random .pws circuits -> the reference's layeredCircuit dump and transcript vs this repo's loader and C oracle.
  python tools/diff_reference_campaign.py FIRST_SEED COUNT
Circuits on which the reference itself aborts (its heap corruption on 1-gate layers, prover.cpp:496) are reported and
skipped. Round 2: seeds 1000-1039, 2000-2039, 3000-3039, 4000-4039, 5000-5059, 6000-6059, 7000-7059, 8000-8059: 0 mismatches."""
import importlib.util, os, random, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)
import helpers as H
import __graft_entry__ as E
B, O = E.binding(), E.oracle()
spec = importlib.util.spec_from_file_location("mg", os.path.join(ROOT, "tests/golden/make_golden.py")); mg = importlib.util.module_from_spec(spec); spec.loader.exec_module(mg)
seed0, n = int(sys.argv[1]), int(sys.argv[2])
bad = 0
for s in range(seed0, seed0 + n):
    rng = random.Random(s * 7919)
    n_in = rng.choice([200, 201, 255, 256, 257, 300, 511, 513])
    n_g = rng.choice([20, 40, 90, 200, 450, 900])
    pws = mg.random_pws(s, n_in, n_g)
    with tempfile.TemporaryDirectory() as td:
        p = os.path.join(td, "c.pws"); open(p, "wb").write(pws)
        r = subprocess.run([mg.REF_DUMP, p, os.path.join(td, "c")], capture_output=True, text=True)
        if "VERIFY 1" not in r.stdout or not os.path.exists(os.path.join(td, "c.transcript.txt")):
            print("seed", s, "reference failed/rc", r.returncode, r.stdout[-200:], r.stderr[-200:]); continue
        tr_ref = open(os.path.join(td, "c.transcript.txt")).read()
        cb_ref = open(os.path.join(td, "c.circuit.bin"), "rb").read()
    try:
        circ = B.Circuit.from_pws_text(pws)
    except B.VpError as e:
        print("seed", s, "OUR LOADER REJECTED:", e); bad += 1; continue
    if H.circuit_dump(circ) != cb_ref:
        print("seed", s, "CIRCUIT MISMATCH", n_in, n_g); bad += 1; continue
    oc = O.OracleCircuit(circ.flat())
    tr, ch, _ = oc.prove()
    if H.transcript_text(circ, tr, ch) != tr_ref:
        print("seed", s, "TRANSCRIPT MISMATCH", n_in, n_g); bad += 1; continue
print("seeds", seed0, "..", seed0 + n - 1, "mismatches", bad)
