#!/usr/bin/env python3
"""Golden vectors for the FRI commit phase of the polynomial commitment (fri::commit_phase_step): runs the
UNMODIFIED reference (oracle/_ref/ref_pc_commit: commit_private_array, commit_public_array, then one commit_phase_step per
fold challenge, log_len - 6 of them, as poly_commit_prover::commit_phase does) and stores in pc_fri.json per case: the root
of every level, SHA-256 of all level codewords, of all level trees (node 0 of each heap excluded: never read) and of the
final 32-point codewords.
  arrays as in make_golden_pc_public.py; fold challenges = numpy default_rng(seed + 200) (sha256_64: default_rng(2224))
Only runnable where /root/reference exists.   usage: make_golden_pc_fri.py [--full]   (--full adds sha256_64_x1024: minutes)"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
import __graft_entry__ as entry  # noqa: E402
import make_golden_pc_public as mkp  # noqa: E402

P = (1 << 61) - 1
CASES = ["random_9_3", "random_10_4", "random_12_6", "sha256_64", "random_16_9"]
FULL = ["sha256_64_x1024"]


def case_inputs(B, O, name):
    a, q, b = mkp.case_arrays(B, O, name)
    seed = 2224 if name.startswith("sha256_64") else int(name.split("_")[2]) + 200
    rng = np.random.default_rng(seed)
    r = np.zeros(b - 6, O.F_DTYPE)
    r["re"] = rng.integers(0, P, b - 6, dtype=np.uint64)
    r["im"] = rng.integers(0, P, b - 6, dtype=np.uint64)
    return a, q, b, r


def digest_of(roots, codes, trees):
    hc, ht = hashlib.sha256(), hashlib.sha256()
    for c in codes:
        hc.update(np.ascontiguousarray(c).tobytes())
    for t in trees:
        ht.update(bytes(t)[32:])
    return {"roots": [bytes(x).hex() for x in roots], "codes_sha256": hc.hexdigest(), "trees_sha256": ht.hexdigest(),
            "final_sha256": hashlib.sha256(np.ascontiguousarray(codes[-1]).tobytes()).hexdigest()}


def main():
    B, O = entry.binding(), entry.oracle()
    path = os.path.join(HERE, "pc_fri.json")
    out = json.load(open(path)) if os.path.exists(path) else {}
    for name in CASES + (FULL if "--full" in sys.argv else []):
        a, q, b, r = case_inputs(B, O, name)
        ref = O.ref_pc_fri(a, q, b, r)
        out[name] = dict(digest_of(ref["roots"], ref["codes"], ref["trees"]), log_len=b, root_l=ref["root_l"].hex(), root_h=ref["root_h"].hex(),
                         reference_fri_commit_seconds=ref["seconds"], reference_commit_seconds=ref["commit_seconds"],
                         reference_commit_public_seconds=ref["commit_public_seconds"])
        print(name, {k: v for k, v in out[name].items() if k != "roots"})
    with open(path, "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
        f.write("\n")


if __name__ == "__main__":
    main()
