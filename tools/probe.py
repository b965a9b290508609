"""Quick GPU probe: timings of C1, C2, C3 (not the bench; used during development)."""
import lzma, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "virgo-plus_b200"))
import binding as B

def c2(log_n=24):
    s = B.Sumcheck(log_n)
    s.fill_random(1)
    rng = np.random.default_rng(0)
    r = np.zeros(log_n, B.F_DTYPE); r["re"] = rng.integers(0, B.P, log_n, dtype=np.uint64); r["im"] = rng.integers(0, B.P, log_n, dtype=np.uint64)
    for it in range(4):
        out, ms = s.run(r)
    rm = s.round_ms()
    N = 1 << log_n
    print(f"C2 2^{log_n}: total {ms:.3f} ms; algorithmic {144*N/1e9:.3f} GB -> {144*N/ms/1e6:.0f} GB/s")
    for j in range(min(6, log_n)):
        nb = (48 * N if j == 0 else 72 * (N >> (j - 1)))
        print(f"  round {j+1}: {rm[j]*1e3:.1f} us  {nb/rm[j]/1e6:.0f} GB/s")
    for it in range(4):
        out2, ms2 = s.run(r, fused=True)
    assert (out2 == out).all()
    print(f"C2 2^{log_n} fused (two rounds per pass, one launch): {ms2:.3f} ms")
    s.close()

def gkr(K, reps=3):
    with lzma.open(os.path.join(ROOT, "tests/golden/SHA256_64.pws.xz")) as f:
        c = B.Circuit.from_pws_text(f.read())
    if K > 1:
        c = c.replicate(K)
    t = time.time(); p = B.Prover(c); print(f"K={K}: create {time.time()-t:.2f}s gates={c.total_gates}")
    ch = c.draw_challenges(); inp = c.inputs()
    p.set_challenges(ch)
    for it in range(reps):
        t = time.time(); p.prove(); w = time.time() - t
        print(f"  resident prove: {p.last_prove_ms:.3f} ms device, {w*1e3:.3f} ms wall, launches {p.last_prove_launches}, {c.total_gates/p.last_prove_ms/1e3:.2f} Mgates/s")
    t = time.time(); tr = p.prove(inputs=inp, challenges=ch); w = time.time() - t
    print(f"  host-io prove: {p.last_prove_ms:.3f} ms device, {w*1e3:.3f} ms wall")
    p.close()

def c4(n_layers=65, log_size=20, reps=3):
    t = time.time(); c = B.Circuit.random(n_layers, log_size, 1); print(f"C4 {n_layers} x 2^{log_size}: host circuit {time.time()-t:.1f}s gates={c.total_gates}")
    t = time.time(); p = B.Prover(c); print(f"  create {time.time()-t:.1f}s")
    p.set_challenges(c.draw_challenges())
    for it in range(reps):
        p.prove()
        print(f"  resident prove: {p.last_prove_ms:.3f} ms device, launches {p.last_prove_launches}, {c.total_gates/p.last_prove_ms/1e3:.2f} Mgates/s")
    p.close()


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what in ("all", "c2"): c2(24)
    if what in ("all", "c1"): gkr(1, 5)
    if what in ("all", "c3"): gkr(int(sys.argv[2]) if len(sys.argv) > 2 else 1024)
    if what == "c4": c4(int(sys.argv[2]) if len(sys.argv) > 2 else 65, int(sys.argv[3]) if len(sys.argv) > 3 else 20)
