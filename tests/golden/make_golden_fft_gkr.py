#!/usr/bin/env python3
"""Golden vectors for the polynomial commitment's inner GKR (fft_circuit_GKR): runs the UNMODIFIED reference
functions in engage_gkr's order (oracle/_ref/ref_fftgkr <lg> <seed>) and stores in fft_gkr.json per case: the SHA-256 of the
randomness stream the run consumed (glibc random() after srand(seed), through fieldElement::random()), of all layer values,
the running claim after each stage, the final alpha / beta, proof_size, the verdict, and fft_gkr's own proof size on the same
stream. The randomness itself is regenerated in the tests with oracle.py's restatement of that stream (ogkr_seed /
ogkr_random_field), so the hash also pins the restated generator.
Only runnable where /root/reference exists."""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402

CASES = [(1, 3), (2, 5), (5, 77), (7, 3396), (10, 11), (13, 2024), (17, 7)]


def fe_hex(x):
    return "%016x%016x" % (int(x["re"]), int(x["im"]))


def main():
    O = entry.oracle()
    out = {}
    for lg, seed in CASES:
        r = O.ref_fft_gkr(lg, seed)
        out["lg%d_seed%d" % (lg, seed)] = {
            "lg": lg, "seed": seed, "rnd_sha256": hashlib.sha256(np.ascontiguousarray(r["rnd"]).tobytes()).hexdigest(),
            "layers_sha256": hashlib.sha256(np.ascontiguousarray(r["layers"]).tobytes()).hexdigest(),
            "claims": [fe_hex(x) for x in r["claims"]], "proof_size": r["proof_size"], "ok": r["ok"], "fft_gkr_ps": r["fft_gkr_ps"],
            "reference_prover_seconds": r["seconds"]}
        print(lg, seed, out["lg%d_seed%d" % (lg, seed)]["proof_size"], r["ok"], r["seconds"])
    with open(os.path.join(HERE, "fft_gkr.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
        f.write("\n")


if __name__ == "__main__":
    main()
