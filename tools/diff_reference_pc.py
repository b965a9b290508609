"""Live differential run of the polynomial-commitment oracle (oracle/pc_oracle.c) against the UNMODIFIED reference
(oracle/_ref/ref_pc_commit; CPU container only): random arrays of 2^9 .. 2^14 elements (log_len 7 and 8 are left out: the
reference's 4-point inverse FFT returns uninitialised memory there), several seeds each -- commit_private_array
(root, l_eval, leaf hashes), commit_public_array (root_h, all_sum, h_eval_arr, virtual oracle) and every level of the FRI commit
phase (roots, codewords, trees).   python tools/diff_reference_pc.py [SEEDS_PER_SIZE]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as E
O = E.oracle()
P = (1 << 61) - 1
per = int(sys.argv[1]) if len(sys.argv) > 1 else 3


def rnd(rng, n, ragged=False):
    a = np.zeros(n, O.F_DTYPE)
    m = n if not ragged else int(rng.integers(n // 2 + 1, n))
    a["re"][:m] = rng.integers(0, P, m, dtype=np.uint64)
    a["im"][:m] = rng.integers(0, P, m, dtype=np.uint64)
    return a


n = bad = 0
for b in range(9, 15):
    for s in range(per):
        rng = np.random.default_rng(1000 * b + s)
        a, q, r = rnd(rng, 1 << b, ragged=(s % 2 == 1)), rnd(rng, 1 << b), rnd(rng, b - 6)
        ref = O.ref_pc_fri(a, q, b, r)
        o1 = O.pc_commit_private(a, b)
        o2 = O.pc_commit_public(a, q, b)
        o3 = O.pc_fri_commit_phase(o2["vow"], b - 1, r)
        rp = O.ref_pc_commit_public(a, q, b)
        same = (bytes(o1["root"]) == bytes(ref["root_l"]) and bytes(o2["root_h"]) == bytes(ref["root_h"])
                and (o2["all_sum"] == rp["all_sum"]).all() and (o2["h_eval"] == rp["h_eval"]).all() and (o2["vow"] == rp["vow"]).all()
                and [bytes(x) for x in o3["roots"]] == [bytes(x) for x in ref["roots"]]
                and all((x == y).all() for x, y in zip(o3["codes"], ref["codes"]))
                and all(bytes(x)[32:] == bytes(y)[32:] for x, y in zip(o3["trees"], ref["trees"])))
        n += 1
        if not same:
            bad += 1
            print("MISMATCH log_len", b, "seed", s)
print("pc cases", n, "mismatches", bad)
