import sys, os, lzma, numpy as np
ROOT="/root/repo"
sys.path.insert(0, os.path.join(ROOT, "virgo-plus_b200"))
import binding as B
with lzma.open(os.path.join(ROOT, "tests/golden/SHA256_64.pws.xz")) as f:
    c = B.Circuit.from_pws_text(f.read())
c = c.replicate(1024)
p = B.Prover(c); p.set_challenges(c.draw_challenges())
ms = []
for it in range(5):
    p.prove(); ms.append(p.last_prove_ms)
print(os.environ.get("VP_DFS_CAP"), os.environ.get("VP_ONE_LANE"), f"C3 {min(ms):.3f} ms")
