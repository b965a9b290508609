"""GPU differential campaign: seeded random .pws circuits (tests/golden/make_golden.py::random_pws: every gate type the parser
emits, operands from any earlier layer, 1-gate layers included) x K data-parallel instances, proved on the device through the
C ABI (whole proof; every third case also method by method; every other case on 1 / 2 / 3 lanes) and compared bit for bit with
the C oracle on the materialised circuit; the device verifier must accept the transcript and give the oracle verifier's verdict (code, layer) on a tampered one.
  python tools/gpu_diff_campaign.py FIRST_SEED SECONDS        (runs until the time budget is used)"""
import importlib.util, os, random, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as E
B, O = E.binding(), E.oracle()
spec = importlib.util.spec_from_file_location("mg", os.path.join(ROOT, "tests/golden/make_golden.py"))
mg = importlib.util.module_from_spec(spec); spec.loader.exec_module(mg)

seed, budget = int(sys.argv[1]), float(sys.argv[2])
t_end = time.time() + budget
n = bad = inter = 0
gates = 0
while time.time() < t_end:
    rng = random.Random(seed * 104729)
    n_in = rng.choice([3, 8, 17, 64, 200, 257, 1000])
    n_g = rng.choice([1, 2, 7, 30, 120, 500, 2000])
    K = rng.choice([1, 1, 2, 3, 5, 8, 13, 33])
    circ = B.Circuit.from_pws_text(mg.random_pws(seed, n_in, n_g))
    if K > 1:
        circ = circ.replicate(K)
    flat = circ.expand() if K > 1 else circ
    oc = O.OracleCircuit(flat.flat())
    want, ch, _ = oc.prove()
    p = B.Prover(circ, device=0)
    if seed % 2:
        p.set_lanes(rng.choice([1, 2, 3]))
    got = p.prove(inputs=circ.inputs(), challenges=circ.draw_challenges())
    same = len(got) == len(want) and (got["re"] == want["re"]).all() and (got["im"] == want["im"]).all()
    ok = p.verify(got) == (True, 0, 0)
    t = got.copy(); k = rng.randrange(len(t)); t[k]["re"] = (int(t[k]["re"]) + 1) % B.P
    rej = tuple(p.verify(t)) == tuple(oc.verify(t))     # same verdict, failure code and layer as the oracle's verifier
    p.close()
    if seed % 3 == 0:
        p = B.Prover(circ, device=0)
        gi = B.prove_interactive(p, circ)
        same = same and (gi["re"] == want["re"]).all() and (gi["im"] == want["im"]).all()
        p.close()
        inter += 1
    n += 1; gates += circ.total_gates
    if not (same and ok and rej):
        bad += 1
        print(f"MISMATCH seed {seed}: n_in {n_in} gates {n_g} K {K} same {same} verifier-accepts {ok} tampered-verdict-equal {rej}", flush=True)
    seed += 1
print(f"gpu_diff_campaign: {n} random circuits ({inter} also method by method, {gates} gates in total), seeds {int(sys.argv[1])}..{seed - 1}: {bad} mismatches")
