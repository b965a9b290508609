#!/usr/bin/env python3
"""Count SASS instructions per kernel and opcode class: tools/sass_count.py file.cubin [kernel-substring]"""
import collections, re, subprocess, sys

def main():
    out = subprocess.run(["cuobjdump", "-sass", sys.argv[1]], capture_output=True, text=True, check=True).stdout
    want = sys.argv[2] if len(sys.argv) > 2 else ""
    fn, counts = None, {}
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            fn = m.group(1); counts[fn] = collections.Counter(); continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and fn:
            op = m.group(1)
            base = op.split(".")[0]
            if base == "IMAD":
                base = "IMAD.WIDE" if ".WIDE" in op else ("IMAD.MOV/IADD/SHL" if any(x in op for x in (".MOV", ".IADD", ".SHL")) else ("IMAD.HI" if ".HI" in op else "IMAD"))
            counts[fn][base] += 1
    for fn, c in counts.items():
        if want not in fn: continue
        tot = sum(c.values())
        print(f"{fn}: total {tot}")
        print("   " + "  ".join(f"{k}={v}" for k, v in sorted(c.items(), key=lambda kv: -kv[1])))

main()
