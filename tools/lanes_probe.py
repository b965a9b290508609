"""Development probe: C1 / C3 / C4 proof time with 3 and 6 lanes."""
import lzma, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "virgo-plus_b200"))
import binding as B
with lzma.open(os.path.join(ROOT, "tests/golden/SHA256_64.pws.xz")) as f:
    sha = B.Circuit.from_pws_text(f.read())
cases = [("C1 SHA256_64", sha), ("C3 SHA256_64 x 1024", sha.replicate(1024))]
if len(sys.argv) > 1: cases.append(("C4 65 x 2^20", B.Circuit.random(65, 20, 1)))
for name, c in cases:
    p = B.Prover(c); p.set_challenges(c.draw_challenges())
    for lanes in (3, 6):
        got = p.set_lanes(lanes)
        for _ in range(3): p.prove()
        ms = []
        for _ in range(6):
            p.prove(); ms.append(p.last_prove_ms)
        print(f"{name}: lanes {got}: {min(ms):.3f} ms ({p.last_prove_launches} launches)")
    p.close()
