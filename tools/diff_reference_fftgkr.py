"""Live differential run of the inner-GKR oracle (oracle/fftgkr_oracle.c) against the unmodified reference functions
(oracle/_ref/ref_fftgkr; CPU container only): lg = 1..12 x 4 seeds -- randomness stream, all layer values, the running claims
the reference exposes, proof size and verdict. Round 2: 48 cases, 0 mismatches."""
import os
import sys, hashlib, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as E
O = E.oracle()
bad = n = 0
for lg in range(1, 13):
    for seed in (1, 2, 99, 12345):
        r = O.ref_fft_gkr(lg, seed)
        rnd = O.draw_challenges(O.fft_gkr_rnd_count(lg), seed=seed)
        assert (rnd["re"] == r["rnd"]["re"]).all() and (rnd["im"] == r["rnd"]["im"]).all()
        got = O.fft_gkr(lg, rnd)
        c = got["claims"]
        same = (np.ascontiguousarray(got["layers"]).tobytes() == np.ascontiguousarray(r["layers"]).tobytes()
                and all((int(c[i]["re"]), int(c[i]["im"])) == (int(x["re"]), int(x["im"])) for i, x in zip((0, 1, 2, 3, 3 + lg, 4 + lg, 5 + lg), r["claims"]))
                and got["proof_size"] == r["proof_size"] == r["fft_gkr_ps"] and got["ok"] and r["ok"])
        n += 1
        if not same: bad += 1; print("MISMATCH", lg, seed)
print("cases", n, "mismatches", bad)
