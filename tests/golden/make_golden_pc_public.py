#!/usr/bin/env python3
"""Golden vectors for commit_public_array (the second half of the polynomial commitment's commit work): runs the
UNMODIFIED reference (oracle/_ref/ref_pc_commit: commit_private_array, then commit_public_array with a public array, zero
masks, target sum = <array, public>) and stores in pc_commit_public.json per case: root_h, SHA-256 of all_sum, of h_eval_arr
and of virtual_oracle_witness.
  random_<b>_<s>   2^b random F_p^2 elements for both arrays (numpy default_rng(s) / default_rng(s + 100))
  sha256_64        the input layer of SHA256_64 against the eq table of a fixed random point (what verifyPoly commits, verifier.cpp:367-383)
b = 7 is left out on purpose (2n = 4: the reference's 4-point inverse FFT returns uninitialised memory, RS_polynomial.cpp:100).
Only runnable where /root/reference exists."""
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
import make_golden_pc as mk  # noqa: E402

P = (1 << 61) - 1
CASES = ["random_6_1", "random_9_3", "random_10_4", "random_12_6", "sha256_64"]


def case_arrays(B, O, name):
    a, b = mk.case_array(B, O, name)
    if name.startswith("sha256_64"):
        rng = np.random.default_rng(2024)
        r = np.zeros(b, O.F_DTYPE)
        r["re"] = rng.integers(0, P, b, dtype=np.uint64)
        r["im"] = rng.integers(0, P, b, dtype=np.uint64)
        q = O.beta_table(r)                      # initBetaTable(output, bitLength, r_liu, F_ONE)
    else:
        rng = np.random.default_rng(int(name.split("_")[2]) + 100)
        q = np.zeros(1 << b, O.F_DTYPE)
        q["re"] = rng.integers(0, P, 1 << b, dtype=np.uint64)
        q["im"] = rng.integers(0, P, 1 << b, dtype=np.uint64)
    return a, q, b


def digest_of(r):
    h = lambda x: hashlib.sha256(np.ascontiguousarray(x).tobytes()).hexdigest()
    return {"root_h": r["root_h"].hex(), "all_sum_sha256": h(r["all_sum"]), "h_eval_sha256": h(r["h_eval"]), "vow_sha256": h(r["vow"]),
            "slice_size": int(r["slice_size"])}


def main():
    B, O = entry.binding(), entry.oracle()
    out = {}
    for name in CASES:
        a, q, b = case_arrays(B, O, name)
        r = O.ref_pc_commit_public(a, q, b)
        assert (r["vow_msk"]["re"] == 0).all() and (r["vow_msk"]["im"] == 0).all()
        out[name] = dict(digest_of(r), log_len=b, reference_commit_public_seconds=r["seconds"])
        print(name, out[name])
    with open(os.path.join(HERE, "pc_commit_public.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
        f.write("\n")


if __name__ == "__main__":
    main()
