"""Run the quick probe (C1, C2 fused, C3, C4-shape) once per library variant under build/variants/ (development tool)."""
import glob, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
libs = ["default"] + sorted(glob.glob(os.path.join(ROOT, "build/variants/*.so")))
pat = sys.argv[1] if len(sys.argv) > 1 else ""
code = r'''
import sys, os, lzma, numpy as np
sys.path.insert(0, os.path.join(%r, "virgo-plus_b200"))
import binding as B
for log_n in (24,):
    s = B.Sumcheck(log_n); s.fill_random(1)
    r = np.zeros(log_n, B.F_DTYPE); r["re"] = np.arange(1, log_n + 1) * 1234567891; r["im"] = 77
    ms = [s.run(r, fused=True)[1] for _ in range(8)]
    print(f"  C2 2^{log_n} fused {min(ms):.4f} ms", end=";")
    s.close()
with lzma.open(os.path.join(%r, "tests/golden/SHA256_64.pws.xz")) as f:
    c1 = B.Circuit.from_pws_text(f.read())
def best(c, n=6):
    p = B.Prover(c); p.set_challenges(c.draw_challenges())
    ms = []
    for it in range(n):
        p.prove(); ms.append(p.last_prove_ms)
    p.close()
    return min(ms)
print(f"  C1 {best(c1, 10):.4f} ms; C3 {best(c1.replicate(1024)):.3f} ms; 65x2^16 {best(B.Circuit.random(65, 16, 7)):.3f} ms")
''' % (ROOT, ROOT)
for lib in libs:
    if pat and pat not in lib and lib != "default":
        continue
    env = dict(os.environ)
    if lib != "default":
        env["VP_LIB"] = lib
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
    print(os.path.basename(lib), out.stdout.strip(), out.stderr.strip()[-300:])
