"""Live differential run of the .pws STATEMENT GRAMMAR against the reference's regex parser (main.cpp:160-204; oracle/_ref/ref_dump):
valid random circuits with match-preserving variations (leading zeros on ids) and inserted lines the reference skips (wrong spacing,
tabs, trailing characters, CR, lower-case operators, truncated statements, overlong numbers on malformed lines, binary bytes):
circuit dump and transcript must be identical. CPU container only.   python tools/diff_reference_pws_format.py FIRST_SEED COUNT
Round 2: found that a MALFORMED line holding an overlong id made the loader reject the file (the reference skips the line); fixed;
seeds 1-50, 101-150, 201-250, 301-350: 0 mismatches."""
import importlib.util, os, random, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)
import helpers as H
import __graft_entry__ as E
B, O = E.binding(), E.oracle()
spec = importlib.util.spec_from_file_location("mg", os.path.join(ROOT, "tests/golden/make_golden.py")); mg = importlib.util.module_from_spec(spec); spec.loader.exec_module(mg)
JUNK = ["", " ", "# comment", "P V1 = V0 +  V0 E", "P V1 = V0 + V0 E ", " P V1 = V0 + V0 E", "P V1 = V0 + V0 E\r", "P V1 = V0 xor V0 E", "P V1 = V0 / V0 E",
        "P V1 = V0 + V0", "V1 = V0 + V0 E", "P V1 = V0 + I0 E", "P V1 = I0 + V0 E", "P V-1 = V0 + V0 E", "P V1 = V0 NOT E", "P  V1 = V0 + V0 E",
        "P V1 = V0\t+ V0 E", "P V1 = V0 + V0 E E", "p V1 = V0 + V0 E", "P V1.0 = V0 + V0 E", "P V1 = V0 MINUS V0 E", "P O = V0 E", "P V = I0 E",
        "P V99999999999999999999999 = I0 E x", "\x00", "P V1 = V0 + V0 E#", "P V1 = V0 +V0 E", "P V1 = I E"]
n = bad = skipped = 0
for s in range(int(sys.argv[1]), int(sys.argv[1]) + int(sys.argv[2])):
    rng = random.Random(s * 977)
    lines = mg.random_pws(s, rng.choice([200, 230, 256]), rng.choice([40, 120, 300])).decode().split("\n")
    lines = [l for l in lines if l]
    out = []
    for l in lines:
        t = l.split()
        r = rng.random()
        if r < 0.10 and len(t) >= 5:            # match-preserving: leading zeros on ids
            t = [("V" + "0" * rng.randint(1, 3) + x[1:]) if (x[0] == "V" and x[1:].isdigit()) else x for x in t]
            l = " ".join(t)
        out.append(l)
        if rng.random() < 0.15:
            out.append(rng.choice(JUNK))
    if rng.random() < 0.5: out.append(rng.choice(JUNK))
    data = ("\n".join(out) + ("\n" if rng.random() < 0.8 else "")).encode("latin-1")
    with tempfile.TemporaryDirectory() as td:
        p = os.path.join(td, "c.pws"); open(p, "wb").write(data)
        r = subprocess.run([mg.REF_DUMP, p, os.path.join(td, "c")], capture_output=True, text=True, errors="replace")
        if "VERIFY 1" not in r.stdout: skipped += 1; continue
        cb = open(os.path.join(td, "c.circuit.bin"), "rb").read(); trr = open(os.path.join(td, "c.transcript.txt")).read()
    n += 1
    try:
        circ = B.Circuit.from_pws_text(data)
    except B.VpError as e:
        bad += 1; print("seed", s, "OUR LOADER REJECTED what the reference accepted:", e); continue
    ok1 = H.circuit_dump(circ) == cb
    if not ok1: bad += 1; print("seed", s, "CIRCUIT MISMATCH"); continue
    tr, ch, _ = O.OracleCircuit(circ.flat()).prove()
    if H.transcript_text(circ, tr, ch) != trr: bad += 1; print("seed", s, "TRANSCRIPT MISMATCH")
print("format cases", n, "skipped", skipped, "mismatches", bad)
