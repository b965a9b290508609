#!/usr/bin/env python3
"""Golden vectors for the commit phase of the polynomial commitment: runs the UNMODIFIED reference
poly_commit_prover::commit_private_array (oracle/_ref/ref_pc_commit, built from /root/reference by `make -C oracle ref`,
SHA3 from the reference's prebuilt XKCP) and stores in pc_commit.json, per case: the Merkle root, and the SHA-256 of the
codeword array l_eval, of the leaf hashes and of the Merkle tree (nodes 1..) it produced.

  sha256_64        the input layer of data/SHA256_64.pws (7226 values, padded to 2^13): what prover::commit_private commits
  sha256_64_x16    16 instances (115616 values, 2^17)
  sha256_64_x1024  the C3 benchmark circuit's input layer (7.4 M values, 2^23; 135 s in the reference; `--full`)
  random_<b>_<s>   2^b random F_p^2 elements, numpy default_rng(s); b = 10 has all-zero slices (the reference short-cuts them)
b = 8 is left out on purpose: the reference's 4-point inverse FFT runs zero iterations of its unrolled stage
(RS_polynomial.cpp:100) and returns uninitialised memory.
Only runnable where /root/reference exists."""
import hashlib
import json
import lzma
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402

P = (1 << 61) - 1


def case_array(B, O, name):
    if name.startswith("sha256_64"):
        with lzma.open(os.path.join(HERE, "SHA256_64.pws.xz"), "rb") as f:
            c = B.Circuit.from_pws_text(f.read())
        if "_x" in name:
            c = c.replicate(int(name.split("_x")[1]))
        a = np.zeros(c.num_inputs, O.F_DTYPE)
        a["re"] = c.inputs()
        return a, c.bit_length(0)
    _, b, s = name.split("_")
    b, s = int(b), int(s)
    rng = np.random.default_rng(s)
    a = np.zeros(1 << b, O.F_DTYPE)
    a["re"] = rng.integers(0, P, 1 << b, dtype=np.uint64)
    a["im"] = rng.integers(0, P, 1 << b, dtype=np.uint64)
    if b == 10:
        a[64:128] = 0
        a[512:] = 0
    return a, b


CASES = ["random_6_1", "random_7_2", "random_9_3", "random_10_4", "random_11_5", "random_12_6", "sha256_64", "sha256_64_x16"]


def digest_of(r):
    return {"root": r["root"].hex(), "l_eval_sha256": hashlib.sha256(np.ascontiguousarray(r["l_eval"]).tobytes()).hexdigest(),
            "leaf_sha256": hashlib.sha256(np.ascontiguousarray(r["leaf_hash"]).tobytes()).hexdigest(),
            "tree_sha256": hashlib.sha256(np.ascontiguousarray(r["tree"][32:]).tobytes()).hexdigest(), "slice_size": int(r["slice_size"])}


def main():
    B, O = entry.binding(), entry.oracle()
    path = os.path.join(HERE, "pc_commit.json")
    out = json.load(open(path)) if os.path.exists(path) else {}
    for name in CASES + (["sha256_64_x1024"] if "--full" in sys.argv else []):
        a, b = case_array(B, O, name)
        r = O.ref_pc_commit(a, b)
        out[name] = dict(digest_of(r), log_len=b, n=int(len(a)), reference_commit_seconds=r["seconds"])
        print(name, out[name])
    with open(os.path.join(HERE, "pc_commit.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
        f.write("\n")


if __name__ == "__main__":
    main()
