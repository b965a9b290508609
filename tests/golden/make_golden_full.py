#!/usr/bin/env python3
"""Full-size golden hashes: runs the UNMODIFIED reference prover (oracle/_ref/libref_gkr.so, built from
/root/reference by `make -C oracle ref`; no polynomial commitment) on the BENCHMARK-size circuits and stores the
SHA-256 of the transcript (the prover's messages as little-endian {u64 re, u64 im} pairs, vp_transcript_len layout)
in full_size.json. The -m gpu tests and bench.py compare the GPU transcript's hash with these.

  sha256_64_x1024   BASELINE.json configs[2] (C3): SHA256_64 x 1024 instances, 94.9 M gates   (~3 min, ~25 GB)
  sha256_64_x2048   the N=2 weak-scaling workload of bench.py                                  (~6 min, ~50 GB)
  random_65x14      configs[3] shape (C4) at 65 layers x 2^14 gates, seed 7
  random_65x20      configs[3] (C4) itself: 65 layers x 2^20 random add/mul gates, seed 7 -- only with --c4-full

Inputs are the circuit's own (drawn like main.cpp:188 draws them), challenges F::random() after srand(3396).
Only runnable where /root/reference exists (this container); usage: make_golden_full.py [case ...] [--c4-full]
"""
import hashlib
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402

OUT = os.path.join(HERE, "full_size.json")


def build_case(B, name):
    import lzma
    if name.startswith("sha256_64_x"):
        k = int(name.split("x")[-1])
        with lzma.open(os.path.join(HERE, "SHA256_64.pws.xz"), "rb") as f:
            c = B.Circuit.from_pws_text(f.read())
        return c.replicate(k) if k > 1 else c
    if name.startswith("random_"):
        n, lg = name.split("_")[1].split("x")
        return B.Circuit.random(int(n), int(lg), 7)
    raise SystemExit("unknown case " + name)


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    cases = args or ["random_65x14", "sha256_64_x1024"]
    if "--c4-full" in sys.argv:
        cases.append("random_65x20")
    B, O = entry.binding(), entry.oracle()
    assert O.ref_available(), "build oracle/_ref first (make -C oracle ref)"
    res = json.load(open(OUT)) if os.path.exists(OUT) else {}
    devnull = os.open(os.devnull, os.O_WRONLY)
    os.dup2(devnull, 2)   # the reference prints per-layer progress on stderr
    for name in cases:
        t0 = time.time()
        circ = build_case(B, name)
        flat = (circ.expand() if circ.instances > 1 else circ).flat()
        tr, prove_s, eval_s = O.ref_prove(flat)
        res[name] = {
            "gates": int(circ.total_gates), "instances": int(circ.instances), "n_layers": int(circ.n_layers),
            "transcript_len": int(len(tr)), "transcript_sha256": hashlib.sha256(tr.tobytes()).hexdigest(),
            "vres": [int(tr[0]["re"]), int(tr[0]["im"])], "input_mle": [int(tr[-1]["re"]), int(tr[-1]["im"])],
            "reference_prove_seconds": round(prove_s, 3), "reference_evaluate_seconds": round(eval_s, 3),
            "generator": "tests/golden/make_golden_full.py (unmodified reference prover, libref_gkr.so)",
        }
        print(name, res[name], "wall %.0f s" % (time.time() - t0), flush=True)
        with open(OUT, "w") as f:
            json.dump(res, f, indent=1, sort_keys=True)
            f.write("\n")
        del flat, tr, circ


if __name__ == "__main__":
    main()
