"""small workloads for compute-sanitizer (memcheck / racecheck / synccheck)"""
import os, sys, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry
B, O = entry.binding(), entry.oracle()
circ = B.Circuit.random(5, 9, 3).replicate(5)
want, _, _ = O.OracleCircuit(circ.expand().flat()).prove()
p = B.Prover(circ)
got = p.prove(inputs=circ.inputs(), challenges=circ.draw_challenges())
assert (got == want).all()
p.close()
p = B.Prover(circ)
got = B.prove_interactive(p, circ)
assert (got == want).all()
s = B.Sumcheck(13); s.fill_random(2)
r = np.zeros(13, B.F_DTYPE); r["re"] = np.arange(1, 14) * 987654321
a, _ = s.run(r); b, _ = s.run(r, fused=True)
assert (a == b).all()
print("sanitize target ok")
