import os, sys, ctypes, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "virgo-plus_b200"))
import binding as B
lib = sys.argv[1]
if lib != "default":
    B.LIB_PATH = lib
for log_n in (20, 22, 24):
    s = B.Sumcheck(log_n); s.fill_random(1)
    r = np.zeros(log_n, B.F_DTYPE); r["re"] = np.arange(1, log_n + 1) * 1234567891; r["im"] = 77
    ms = [s.run(r, fused=True)[1] for _ in range(6)]
    print(lib.split("/")[-1], log_n, f"{min(ms):.4f} ms")
    st = s.pass_stamps().astype(np.int64)
    t0 = st[0, 0]
    print("   pass: start  work_done  barrier  end   (us, relative)")
    for i, row in enumerate(st):
        print(f"   {i:2d}: " + "  ".join(f"{(x - t0) / 1e3:8.1f}" if x else "       -" for x in row))

    bs = s.block_stamps.astype(np.int64)
    bs = bs[bs[:, 0] > 0]
    rel = (bs[:, 0] - t0) / 1e3
    order = np.argsort(rel)
    print("   blocks", len(bs), "finish us: min %.1f median %.1f max %.1f" % (rel.min(), np.median(rel), rel.max()))
    sm = bs[:, 1]
    import collections
    per_sm = collections.Counter(sm.tolist())
    print("   blocks per SM histogram:", collections.Counter(per_sm.values()))
    print("   slowest 8 (block, sm, us):", [(int(i), int(sm[i]), round(float(rel[i]), 1)) for i in order[-8:]])
    print("   fastest 8 (block, sm, us):", [(int(i), int(sm[i]), round(float(rel[i]), 1)) for i in order[:8]])
    # finish time vs blocks-on-that-SM
    for cnt in sorted(set(per_sm.values())):
        sel = [rel[i] for i in range(len(bs)) if per_sm[int(sm[i])] == cnt]
        print(f"   SMs with {cnt} block(s): mean finish {np.mean(sel):.1f} us over {len(sel)} blocks")
    s.close()
