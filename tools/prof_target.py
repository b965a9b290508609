"""Short workloads for ncu captures: `c2` = one 2^24 sumcheck, `gkr K` = one proof of SHA256_64 x K,
`fftgkr LG` = the commitment's inner GKR,
`pc LOG_LEN` = both commitments + the FRI commit phase of a random 2^LOG_LEN-entry array."""
import lzma, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "virgo-plus_b200"))
import binding as B

if sys.argv[1] == "c2":
    log_n = int(sys.argv[2]) if len(sys.argv) > 2 else 24
    s = B.Sumcheck(log_n); s.fill_random(1)
    r = np.zeros(log_n, B.F_DTYPE); r["re"] = np.arange(1, log_n + 1) * 1234567891; r["im"] = 77
    s.run(r); s.run(r)
    if len(sys.argv) > 3: s.run(r, fused=True); s.run(r, fused=True)
elif sys.argv[1] == "fftgkr":
    lg = int(sys.argv[2])
    rnd = B.draw_field(B.fft_gkr_rnd_count(lg), 9)
    B.fft_gkr(lg, rnd, want_layers=False)
    out = B.fft_gkr(lg, rnd, want_layers=False)
    print("fft_gkr ms", out["device_ms"], out["ok"])
elif sys.argv[1] == "pc":
    b = int(sys.argv[2])
    rng = np.random.default_rng(1)
    def rnd(n):
        a = np.zeros(n, B.F_DTYPE)
        a["re"] = rng.integers(0, (1 << 61) - 1, n, dtype=np.uint64); a["im"] = rng.integers(0, (1 << 61) - 1, n, dtype=np.uint64)
        return a
    out = B.pc_fri(rnd(1 << b), rnd(1 << b), b, rnd(b - 6), want_arrays=False)
    print("fri ms", out["ms"])
else:
    K = int(sys.argv[2])
    with lzma.open(os.path.join(ROOT, "tests/golden/SHA256_64.pws.xz")) as f:
        c = B.Circuit.from_pws_text(f.read()).replicate(K)
    p = B.Prover(c); p.set_challenges(c.draw_challenges())
    if len(sys.argv) > 3: p.set_lanes(int(sys.argv[3]))   # 1: like bench.py's instrumented pass (a launch has the GPU to itself)
    p.prove(); p.prove()
    print("ms", p.last_prove_ms)
