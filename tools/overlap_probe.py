"""Development probe: N independent provers (SHA256_64 x K/N each) proving concurrently on one GPU from N host
threads vs one prover with K instances -- how much does more kernel-level overlap buy?"""
import sys, os, lzma, time, threading
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "virgo-plus_b200"))
import binding as B
with lzma.open(os.path.join(ROOT, "tests/golden/SHA256_64.pws.xz")) as f:
    base = B.Circuit.from_pws_text(f.read())
K = 1024
for n in (1, 2, 4):
    cs = [base.replicate(K // n) for _ in range(n)]
    ps = [B.Prover(c) for c in cs]
    for p, c in zip(ps, cs):
        p.set_challenges(c.draw_challenges()); p.prove()
    best = 1e9
    for rep in range(4):
        th = [threading.Thread(target=lambda p=p: (p.prove(), p.prove())) for p in ps]
        t0 = time.perf_counter()
        [t.start() for t in th]; [t.join() for t in th]
        best = min(best, (time.perf_counter() - t0) / 2)
    print(f"{n} concurrent prover(s) x {K // n} instances: {best * 1e3:.2f} ms per {K}-instance batch (wall), single last_ms {ps[0].last_prove_ms:.2f}")
    for p in ps: p.close()
