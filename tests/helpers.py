"""Shared test helpers: canonical circuit dump and transcript text in the formats of
oracle/ref_harness/ref_dump.cpp (what the golden fixtures were recorded in)."""
import hashlib
import lzma
import os
import struct

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
INT_MIN = -(1 << 31)


def golden_bytes(name):
    with lzma.open(os.path.join(GOLDEN, name), "rb") as f:
        return f.read()


def golden_text(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return f.read()


def ceil_log2(x):
    if x == 0:
        return -1
    b = x.bit_length() - 1
    return b + 1 if (1 << b) < x else b


def circuit_dump(circ):
    """bytes identical to ref_dump.cpp::dump_circuit for the reference's layeredCircuit (instances == 1)."""
    assert circ.instances == 1
    n = circ.n_layers
    inputs = circ.inputs()
    out = [struct.pack("<i", n)]
    rec = np.dtype([("ty", "u1"), ("l", "<i4"), ("u", "<u8"), ("v", "<u8"), ("lv", "<u8")])
    for i in range(n):
        L = circ.export_layer(i)
        size = circ.layer_size(i)
        dad = [circ.export_dad(i, l) for l in range(i)]
        dbl = [ceil_log2(len(d)) if len(d) else INT_MIN for d in dad]  # (int)log2(0) == INT_MIN, circuit.cpp:73
        mdb = max([-1] + dbl)
        out.append(struct.pack("<Qii", size, ceil_log2(size), mdb))
        g = np.zeros(size, rec)
        g["ty"] = L["ty"]
        g["l"] = L["l"]
        g["u"] = inputs if i == 0 else L["u"]
        g["v"] = L["v"]
        g["lv"] = L["lv"]
        out.append(g.tobytes())
        for l in range(i):
            out.append(struct.pack("<Qi", len(dad[l]), dbl[l]))
            out.append(dad[l].astype("<u8").tobytes())
    return b"".join(out)


def circuit_digest(circ):
    return hashlib.sha256(circuit_dump(circ)).hexdigest()


def transcript_text(circ, tr, ch):
    """The "TAG real img" dump ref_dump writes, rebuilt from a transcript + the challenge stream."""
    n = circ.n_layers
    max_bl = max(circ.bit_length(i) for i in range(n))
    lines = []
    fe = lambda tag, x: lines.append(f"{tag} {int(x['re'])} {int(x['im'])}")
    zero = {"re": 0, "im": 0}
    ti = ci = 0
    ci += circ.bit_length(n - 1)
    fe("VRES", tr[ti]); ti += 1

    def rounds(count, r):
        nonlocal ti
        prev = zero
        for j in range(count):
            fe("CH", prev)
            fe("PA", tr[ti]); fe("PB", tr[ti + 1]); fe("PC", tr[ti + 2])
            ti += 3
            prev = r[j]
        return prev

    for i in range(n - 1, 0, -1):
        pb, m = circ.bit_length(i - 1), circ.max_dad_bit_length(i)
        r_u = ch[ci:ci + max_bl]; ci += max_bl
        ci += 1  # assert_random
        prev = rounds(pb, r_u)
        fe("CH", prev); fe("CLAIM_U", tr[ti]); ti += 1
        if m != -1:
            r_v = ch[ci:ci + m]; ci += m
            rounds(m, r_v)
            for l in range(i):
                fe("CLAIM_V", tr[ti]); ti += 1
        ci += n  # sig
        r_liu = ch[ci:ci + max_bl]; ci += max_bl
        prev = rounds(pb, r_liu)
        fe("CH", prev); fe("CLAIM_LIU", tr[ti]); ti += 1
    fe("INPUT_MLE", tr[ti]); ti += 1
    assert ti == len(tr) and ci == len(ch)
    return "\n".join(lines) + "\n"
