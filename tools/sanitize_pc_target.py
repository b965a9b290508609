"""small polynomial-commitment workloads for compute-sanitizer (memcheck / racecheck / synccheck)"""
import os, sys, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry
B, O = entry.binding(), entry.oracle()
rng = np.random.default_rng(4)
# commit_public + the FRI commit phase (shared-memory tiles of k_pc_dit_group / k_pc_vow, the all-level leaf
# hash), step by step and batched, small (everything in shared memory) and 2^18 entries (global stages)
if True:
    for b in (9, 18):
        a = np.zeros((1 << b) - 3, B.F_DTYPE); q = np.zeros(1 << b, B.F_DTYPE); r = np.zeros(b - 6, B.F_DTYPE)
        for x in (a, q, r):
            x["re"] = rng.integers(0, B.P, len(x), dtype=np.uint64); x["im"] = rng.integers(0, B.P, len(x), dtype=np.uint64)
        g = B.pc_fri(a, q, b, r, want_arrays=(b == 9))
        if b == 9:
            w = O.pc_fri_commit_phase(O.pc_commit_public(a, q, b)["vow"], b - 1, r)
            assert g["roots"] == w["roots"] and all((x == y).all() for x, y in zip(g["codes"], w["codes"]))
    os.environ["VP_FRI_STEPWISE"] = "1"
    g2 = B.pc_fri(a, q, b, r, want_arrays=False)
    assert g2["roots"] == g["roots"]
    print("sanitize target pc ok")
# the commitment's inner GKR: circuit kernels, table fills (incl. the device-side v_u feed), fused sumchecks on two streams
for lg in (1, 4, 9):
    rnd = B.draw_field(B.fft_gkr_rnd_count(lg), 5 + lg)
    g = B.fft_gkr(lg, rnd)
    w = O.fft_gkr(lg, rnd)
    assert g["ok"] and (g["polys"] == w["polys"]).all() and (g["layers"] == w["layers"]).all() and g["proof_size"] == w["proof_size"]
print("sanitize target fft_gkr ok")
