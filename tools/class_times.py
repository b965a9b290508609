"""Development probe: per-kernel-class device times of one C3 proof on ONE lane (no overlap), and the 3-lane total."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for env in ({"VP_ONE_LANE": "1"}, {}):
    e = dict(os.environ); e.update(env)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "3", "--warmup", "3", "--no-extras"], env=e, capture_output=True, text=True).stdout
    d = json.loads(out.strip().splitlines()[-1])
    print(env, round(d["ms_per_step"], 3), {k: round(v["ms_per_step"], 2) for k, v in d["kernel_classes"].items() if isinstance(v, dict)})
