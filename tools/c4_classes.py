"""Development probe: per-kernel-class times of the full-size C4 circuit (65 layers x 2^20 random add/mul gates)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "virgo-plus_b200"))
import binding as B
c = B.Circuit.random(65, 20, 1)
p = B.Prover(c); p.set_challenges(c.draw_challenges())
for lanes in (3, 1):
    p.set_lanes(lanes)
    for _ in range(2): p.prove()
    ms = []
    for _ in range(3):
        p.prove(); ms.append(p.last_prove_ms)
    print("lanes", lanes, "prove ms", round(min(ms), 3), "launches", p.last_prove_launches)
p.set_lanes(1); p.set_profiling(True)
for _ in range(2): p.prove()
prof = p.profile()
print({k: (round(v["ms"] / 2, 3), v["launches"] // 2) for k, v in prof.items() if v["launches"]})
