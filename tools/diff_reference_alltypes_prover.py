"""Live differential run of the oracle PROVER against the UNMODIFIED reference prover (oracle/_ref/libref_gkr.so, entry point
ref_gkr_prove2: gate constants and assert flags) on random layered circuits over all 11 gate types -- Addc / Mulc with real and
complex constants, Copy, Not, assert gates -- plain (240 circuits) and replicated K = 2, 3, 5, 8 times through
vp_circuit_replicate + vp_circuit_expand (60 circuits; the reference recomputes subsetInit itself, so this also pins the
replication's subset numbering). CPU container only.   Round 2: 300 circuits, 0 mismatches."""
import importlib.util, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as E
B, O = E.binding(), E.oracle()
spec = importlib.util.spec_from_file_location("tg", os.path.join(ROOT, "tests", "test_gpu_parity.py"))
tg = importlib.util.module_from_spec(spec); spec.loader.exec_module(tg)


def same(flat):
    got, _, _ = O.OracleCircuit(flat).prove()
    want, _, _ = O.ref_prove(flat)
    return bool((got["re"] == want["re"]).all() and (got["im"] == want["im"]).all())


n = bad = 0
for seed in range(1, 61):
    for cc, wa in ((False, False), (True, False), (False, True), (True, True)):
        circ = tg._all_types_circuit(B, seed, n_layers=3 + seed % 5, max_size=8 + 5 * (seed % 7), complex_consts=cc, with_assert=wa)
        n += 1
        if not same(circ.flat()):
            bad += 1; print("MISMATCH seed", seed, "complex", cc, "assert", wa)
for seed in range(1, 61):
    K = [2, 3, 5, 8][seed % 4]
    circ = tg._all_types_circuit(B, 1000 + seed, n_layers=3 + seed % 4, max_size=6 + 4 * (seed % 5), complex_consts=bool(seed & 1), with_assert=bool(seed & 2))
    n += 1
    if not same(circ.replicate(K).expand().flat()):
        bad += 1; print("MISMATCH replicated seed", seed, "K", K)
print("all-gate-type circuits", n, "mismatches", bad)
