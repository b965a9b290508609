#!/usr/bin/env python3
"""Per-launch ncu counters of k_phase_dfs over whole proofs -> profiles/r2_dfs_counters.json (read by bench.py).

  ncu --metrics <METRICS below> --clock-control none -k regex:k_phase_dfs --csv --log-file gpurun_out/r2_dfs_counters.csv \
      python tools/prof_target.py gkr 1024
  python tools/dfs_counters.py gpurun_out/r2_dfs_counters.csv 1024 profiles/r2_dfs_counters.json

Averages are over all captured launches (the 42 phase kernels of one SHA256_64 x K proof, as many proofs as captured),
like bench.py's `roofline.achieved`. The pipe utilisations are ncu's own "% of peak sustained" figures, weighted by each
launch's duration."""
import collections, csv, json, sys

METRICS = ("gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,"
           "sm__inst_executed_pipe_alu.sum,sm__inst_executed_pipe_fma.sum,sm__inst_executed_pipe_fmaheavy.sum,"
           "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,"
           "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,"
           "sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,sm__cycles_elapsed.avg")


def num(v, unit):
    x = float(v.replace(",", ""))
    return x * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1, "usecond": 1e-6,
                "nsecond": 1e-9, "msecond": 1e-3, "second": 1}.get(unit, 1)


def main():
    if len(sys.argv) < 4:
        print(METRICS)
        return
    rows = list(csv.reader(open(sys.argv[1])))
    hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    h = rows[hi]
    ik, im, iu, iv, iid = (h.index(x) for x in ("Kernel Name", "Metric Name", "Metric Unit", "Metric Value", "ID"))
    per = collections.defaultdict(dict)
    for r in rows[hi + 1:]:
        if len(r) > iv and "k_phase_dfs" in r[ik]:
            try:
                per[int(r[iid])][r[im]] = num(r[iv], r[iu])
            except ValueError:
                pass
    n = len(per)
    tot = lambda k: sum(d.get(k, 0.0) for d in per.values())
    t = tot("gpu__time_duration.sum")
    wavg = lambda k: sum(d.get(k, 0.0) * d.get("gpu__time_duration.sum", 0.0) for d in per.values()) / t if t else None
    out = {
        "kernel": "k_phase_dfs", "workload": f"SHA256_64 x {sys.argv[2]}", "instances": int(sys.argv[2]), "launches": n, "n_sm": 148,
        "ncu_seconds_per_launch": t / n,
        "traffic_bytes_per_launch": (tot("dram__bytes_read.sum") + tot("dram__bytes_write.sum")) / n,
        "dram_bytes_read_per_launch": tot("dram__bytes_read.sum") / n, "dram_bytes_write_per_launch": tot("dram__bytes_write.sum") / n,
        "warp_insts_per_launch": tot("smsp__inst_executed.sum") / n,
        "alu_pipe_insts_per_launch": tot("sm__inst_executed_pipe_alu.sum") / n,
        "fma_pipe_insts_per_launch": tot("sm__inst_executed_pipe_fma.sum") / n,
        "fmaheavy_pipe_insts_per_launch": tot("sm__inst_executed_pipe_fmaheavy.sum") / n,
        "issue_active_pct": wavg("smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "alu_pipe_pct": wavg("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active"),
        "fma_pipe_pct": wavg("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
        "fmaheavy_pipe_pct": wavg("sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active"),
        "warps_active_pct": wavg("sm__warps_active.avg.pct_of_peak_sustained_active"),
        "registers_per_thread": max((d.get("launch__registers_per_thread", 0) for d in per.values()), default=None),
        "source": "ncu --metrics (tools/dfs_counters.py METRICS) --clock-control none -k regex:k_phase_dfs on tools/prof_target.py gkr K",
    }
    # the integer roofline of THIS instruction mix: one warp instruction per scheduler per clock (4 / clk / SM); the ALU pipe
    # and the FMA-heavy pipe (IMAD*) each retire 2 warp instructions / clk / SM (profiles/r1_microbench_int_pipes.txt)
    w, a, f = out["warp_insts_per_launch"], out["alu_pipe_insts_per_launch"], out["fmaheavy_pipe_insts_per_launch"] or out["fma_pipe_insts_per_launch"]
    if w:
        cycles = max(w / 4.0, a / 2.0, f / 2.0)
        out["mix_limited_ipc_per_sm"] = w / cycles
        out["mix_limit"] = "issue" if cycles == w / 4.0 else ("alu pipe" if cycles == a / 2.0 else "fma-heavy pipe")
    json.dump(out, open(sys.argv[3], "w"), indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
