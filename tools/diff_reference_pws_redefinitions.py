"""Live differential run of .pws REDEFINITIONS and statement ORDER against the reference parser (buildGate / buildInput overwrite,
every input line draws one random() value in file order, layering is topological: main.cpp:15-137,176-231): inputs and gates defined
again, statements swapped. Circuit dump and transcript must be identical. CPU container only.
  python tools/diff_reference_pws_redefinitions.py FIRST_SEED COUNT      Round 2: 146 comparable cases, 0 mismatches."""
import importlib.util, os, random, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)
import helpers as H
import __graft_entry__ as E
B, O = E.binding(), E.oracle()
spec = importlib.util.spec_from_file_location("mg", os.path.join(ROOT, "tests/golden/make_golden.py")); mg = importlib.util.module_from_spec(spec); spec.loader.exec_module(mg)
n = bad = skipped = 0
for s in range(int(sys.argv[1]), int(sys.argv[1]) + int(sys.argv[2])):
    rng = random.Random(s * 7331)
    n_in = rng.choice([200, 230, 256]); n_g = rng.choice([40, 120, 300])
    lines = [l for l in mg.random_pws(s, n_in, n_g).decode().split("\n") if l]
    ops = ["+", "*", "XOR", "minus", "NAAB"]
    for _ in range(rng.randint(1, 6)):
        k = rng.random()
        pos = rng.randrange(len(lines) + 1)
        if k < 0.4:      # an input defined again (draws another random value; the later line wins)
            i = rng.randrange(n_in); lines.insert(pos, f"P V{i} = I{rng.randrange(1000)} E")
        elif k < 0.8:    # a gate defined again with other operands / operator (operands below its id: still a DAG)
            g = rng.randrange(n_in, n_in + n_g); a, b = rng.randrange(g), rng.randrange(g)
            lines.insert(pos, f"P V{g} = V{a} {rng.choice(ops)} V{b} E")
        else:            # statements in another order
            i, j = rng.randrange(len(lines)), rng.randrange(len(lines)); lines[i], lines[j] = lines[j], lines[i]
    data = ("\n".join(lines) + "\n").encode()
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
    if H.circuit_dump(circ) != cb: bad += 1; print("seed", s, "CIRCUIT MISMATCH"); continue
    tr, ch, _ = O.OracleCircuit(circ.flat()).prove()
    if H.transcript_text(circ, tr, ch) != trr: bad += 1; print("seed", s, "TRANSCRIPT MISMATCH")
print("redefinition cases", n, "skipped", skipped, "mismatches", bad)
