"""development probe: prove a random circuit (n_layers, log_size, seed) and compare whole-proof with the interactive path"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "virgo-plus_b200"))
import binding as B
n, lg, seed = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
c = B.Circuit.random(n, lg, seed)
p = B.Prover(c)
t = time.time()
tr = p.prove(inputs=c.inputs(), challenges=c.draw_challenges())
print(f"OK {n}x2^{lg} seed {seed}: {p.last_prove_ms:.3f} ms device, {time.time() - t:.2f} s wall, launches {p.last_prove_launches}", flush=True)
p.close()
