#!/usr/bin/env python3
"""bench.py -- GKR prover throughput (gates/s) and sumcheck ms/proof on B200, beside the CPU reference.

  python bench.py --gpus N --steps K --warmup W            our arm (CUDA, through the C ABI)
  python bench.py --impl reference --gpus N --steps K ...  the reference's CPU prover on the host cores

One step = one complete GKR proof (evaluate + every phase-1 / phase-2 / Liu sumcheck of every layer +
input-layer MLE) of the workload circuit on synthetic inputs:
  N = 1 : BASELINE.json configs[2], SHA256_64 x 1024 data-parallel instances (94.9 M gates).
  N > 1 : weak scaling, 1024 instances per GPU (see DESIGN.md "Multi-GPU").
`value` = gates/s with inputs and challenges resident in HBM; `e2e` = the same through
vp_prove(host_io=1): inputs + challenges copied from pinned host memory and the transcript copied
back inside the timed region. Prints ONE JSON line on rank 0.
"""
import argparse
import ctypes
import json
import lzma
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402

SHA_PWS = os.path.join(ROOT, "tests", "golden", "SHA256_64.pws.xz")
METRIC = "gkr_prover_gates_per_s"
UNIT = "gates/s"


def _jd(o):
    if isinstance(o, (np.floating,)):
        return float(o)
    if isinstance(o, (np.integer,)):
        return int(o)
    raise TypeError(str(type(o)))


def load_sha(B):
    with lzma.open(SHA_PWS, "rb") as f:
        return B.Circuit.from_pws_text(f.read())


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            j = json.load(open(p))
            return float(j["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)"
        except Exception:
            pass
    return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons while the timed region runs."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu = gpu_index
        self.samples = []
        self.stop_flag = threading.Event()

    def run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.gpu)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.1)

    def summary(self):
        self.stop_flag.set()
        self.join(timeout=3)
        sm, mx, reasons = [], [], set()
        for s in self.samples:
            try:
                sm.append(float(s[1]))
                mx.append(float(s[2]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------- our arm
def run_ours(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run --nproc-per-node {args.gpus}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product has no CPU path)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    B = entry.binding()

    inst = args.instances
    tmpl = load_sha(B)
    # weak scaling: `inst` instances per GPU; ONE proof of the (inst * world)-instance circuit, its sumcheck tables
    # dealt out block-cyclically to the ranks (DESIGN.md "Multi-GPU")
    circ = tmpl.replicate(inst * world)
    gates = circ.total_gates // world
    if world == 1:
        prover = B.Prover(circ, device=local_rank)
    else:
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt.copy_(torch.from_numpy(B.nccl_unique_id()))
        dist.broadcast(idt, 0)
        prover = B.Prover(circ, device=local_rank, rank=rank, world=world, nccl_id=idt.cpu().numpy())
    stream = torch.cuda.Stream(device=local_rank)
    prover.set_stream(stream.cuda_stream)

    ch = circ.draw_challenges()
    n_in = circ.num_inputs
    pin_in = torch.empty(n_in, dtype=torch.int64).pin_memory()
    pin_ch = torch.empty(len(ch) * 2, dtype=torch.int64).pin_memory()
    pin_tr = torch.empty(circ.transcript_len * 2, dtype=torch.int64).pin_memory()
    np_in = pin_in.numpy().view(np.uint64)
    np_in[:] = circ.inputs()
    np_ch = pin_ch.numpy().view(np.uint64).view(B.F_DTYPE)
    np_ch[:] = ch
    np_tr = pin_tr.numpy().view(np.uint64).view(B.F_DTYPE)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---------------- resident: inputs + challenges already in HBM
    prover.set_inputs(np_in)
    prover.set_challenges(np_ch)
    for _ in range(args.warmup):
        prover.prove()
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record(stream)
        for _ in range(args.steps):
            prover.prove()
        e1.record(stream)
    barrier()
    ms_resident = max_over_ranks(e0.elapsed_time(e1)) / args.steps
    launches = prover.last_prove_launches
    # the same K steps again with a CUDA-event pair around every kernel launch (per-class device time for the
    # roofline object; the event pairs cost a few % so `value` comes from the un-instrumented pass above)
    prover.set_profiling(True)
    lanes = prover.set_lanes(1)   # one lane: a launch's event-pair duration is its own (no other phase's kernels on the SMs)
    barrier()
    with torch.cuda.stream(stream):
        e0.record(stream)
        for _ in range(args.steps):
            prover.prove()
        e1.record(stream)
    barrier()
    ms_profiled = max_over_ranks(e0.elapsed_time(e1)) / args.steps
    prof = prover.profile()
    prover.set_profiling(False)
    lanes = prover.set_lanes(6)   # the default: 6 on one GPU, 3 on a sharded context

    # ---------------- e2e: host buffers in, transcript out, every step
    for _ in range(max(1, args.warmup // 2)):
        prover.prove(inputs=np_in, challenges=np_ch, transcript=np_tr)
    barrier()
    with torch.cuda.stream(stream):
        e0.record(stream)
        for _ in range(args.steps):
            prover.prove(inputs=np_in, challenges=np_ch, transcript=np_tr)
        e1.record(stream)
    barrier()
    ms_e2e = max_over_ranks(e0.elapsed_time(e1)) / args.steps
    clocks = sampler.summary()
    tr_e2e = np_tr.copy()

    total_gates = gates * world
    line = {
        "metric": METRIC, "value": total_gates / (ms_resident * 1e-3), "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_resident, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u64 limbs (F_p^2, p=2^61-1)", "data": "synthetic",
        "config": {
            "workload": f"SHA256_64 x {inst * world} data-parallel instances ({inst} per GPU; BASELINE.json configs[2] at N=1, "
                        f"configs[4] family at N>1), {gates} gates per GPU, one full GKR proof per step",
            "instances_per_gpu": inst, "gates_per_gpu": gates, "rounds": None,
            "l2": "tables + values are several GB per proof, far larger than the 126 MB L2 (no flush needed)",
            "lanes": lanes,
            "parallelism": "1 GPU" if world == 1 else f"one proof sharded over {world} GPUs: tables block-cyclic by index, local rounds "
                           f"without communication, one NCCL all-gather per sumcheck phase, evaluate replicated",
        },
        "e2e": {"value": total_gates / (ms_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e,
                "h2d_bytes_per_step": int(np_in.nbytes + np_ch.nbytes) * world, "d2h_bytes_per_step": int(np_tr.nbytes) * world},
        "gpu_launches": int(launches) * args.steps,
        "clocks": clocks,
    }
    # roofline of the dominant kernel (K6d, k_phase_dfs: all rounds of one sumcheck phase) from live CUDA-event timing of
    # every launch. achieved = SURVEY 8(d) algorithmic bytes (144 B per live table entry per sumcheck) / device time.
    peak, peak_src = measured_peaks()
    rf = prof["round_fold"]
    if rf["launches"]:
        ach = rf["bytes"] / (rf["ms"] * 1e-3) / 1e9
        step_share = rf["ms"] / sum(v["ms"] for v in prof.values())   # share of the summed kernel time (the two lanes overlap)
        line["roofline"] = {"bound": "hbm", "kernel": "k_phase_dfs (all rounds of one sumcheck phase in one cooperative launch: fused fold + round polynomials, two rounds per pass)",
                            "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": dfs_traffic(inst, world),
                            "peak_source": peak_src, "launches_per_step": rf["launches"] // args.steps,
                            "avg_launch_us": rf["ms"] * 1e3 / rf["launches"], "share_of_step": step_share,
                            "ms_per_step_instrumented": ms_profiled,
                            "bytes_model": "144 B per live table entry per sumcheck (SURVEY 8d: 3 tables x 16 B, read N_k + write N_k/2 "
                                           "per round); the kernel itself moves ~80 B per entry (two rounds per pass) and is bound by "
                                           "integer-pipe latency, see DESIGN.md and profiles/",
                            "alg_bytes_per_launch": rf["bytes"] / rf["launches"],
                            "note": "launch durations come from the instrumented pass, which runs the phases on ONE stream "
                                    "(vp_set_lanes(1)); `value` is the un-instrumented pass with the lanes overlapped (six streams on one GPU, three per rank when sharded), so the "
                                    "per-class times add up to more than the step. `traffic` = DRAM bytes read + written per launch "
                                    "(ncu, profiles/r1_dfs_traffic.json), averaged over the 42 launches of one proof like `achieved`"}
    line["kernel_classes"] = {k: {"ms_per_step": v["ms"] / args.steps, "GBps": (v["bytes"] / (v["ms"] * 1e-3) / 1e9 if v["ms"] else None),
                                  "launches_per_step": v["launches"] // args.steps} for k, v in prof.items() if v["launches"]}

    if rank == 0 and world == 1 and not args.no_extras:
        line["sumcheck_c2"] = run_c2(B, peak)
        line["single_proof_c1"] = run_c1(B, tmpl)
        line["cpu_baseline"] = cpu_baseline_sample(args)
        # parity spot check of what was just timed: the oracle verifier accepts a K-instance sample? The
        # full-size transcript is checked by tests (size-independent properties); here only sanity.
        line["transcript_nonzero"] = bool(np.any(tr_e2e["re"]))
        # full-size acceptance: the device-side verifier (vp_verify: verifier.cpp's checks, O(#gates) sums on the GPU,
        # no code or tables shared with the prover) on the transcript the e2e pass just produced
        t0 = time.time()
        ok, code, layer = prover.verify(tr_e2e)
        line["verifier"] = {"accept": bool(ok), "fail_code": int(code), "fail_layer": int(layer), "wall_ms": (time.time() - t0) * 1e3,
                            "what": "vp_verify on the SHA256_64 x %d transcript of the e2e pass" % (inst * world)}
    prover.close()
    if world > 1:
        dist.destroy_process_group()
    if rank == 0:
        emit(line)


def dfs_traffic(inst, world):
    """ncu-measured DRAM bytes per k_phase_dfs launch for this workload (tools/dfs_traffic.py), or None"""
    try:
        with open(os.path.join(ROOT, "profiles", "r1_dfs_traffic.json")) as f:
            d = json.load(f)
        return d["traffic_bytes_per_launch"] if d.get("instances") == inst and world == 1 else None
    except Exception:
        return None


def run_c2(B, peak):
    """BASELINE.json configs[1]: stand-alone multilinear sumcheck, random F_p^2 tables of 2^24 entries."""
    log_n = 24
    s = B.Sumcheck(log_n)
    s.fill_random(1)
    rng = np.random.default_rng(0)
    r = np.zeros(log_n, B.F_DTYPE)
    r["re"] = rng.integers(0, B.P, log_n, dtype=np.uint64)
    r["im"] = rng.integers(0, B.P, log_n, dtype=np.uint64)
    for _ in range(3):
        s.run(r)
    ms = [s.run(r)[1] for _ in range(10)]
    rm = s.round_ms()
    for _ in range(3):
        s.run(r, fused=True)
    msf = [s.run(r, fused=True)[1] for _ in range(10)]
    N = 1 << log_n
    alg = 144 * N - 144
    out = {"workload": "3 tables x 2^24 random F_p^2 entries, 24 rounds",
           "ms_per_proof": statistics.median(msf), "ms_per_proof_best": min(msf),
           "mode": "one cooperative launch, two rounds per pass (challenges known up front, as in vp_prove)",
           "algorithmic_GB": alg / 1e9, "achieved_GBps": alg / (statistics.median(msf) * 1e-3) / 1e9,
           "frac_of_peak": alg / (statistics.median(msf) * 1e-3) / 1e9 / peak,
           "one_round_per_launch": {"ms_per_proof": statistics.median(ms), "what": "the interactive path: 24 launches, each round waits for its challenge",
                                    "round2_GBps": 72 * N / (rm[1] * 1e-3) / 1e9, "round2_frac_of_peak": 72 * N / (rm[1] * 1e-3) / 1e9 / peak}}
    s.close()
    return out


def run_c1(B, tmpl):
    """BASELINE.json configs[0] on the GPU: one SHA256_64 proof (latency-bound: 439 dependent rounds)."""
    p = B.Prover(tmpl)
    p.set_challenges(tmpl.draw_challenges())
    for _ in range(3):
        p.prove()
    ms = []
    for _ in range(10):
        p.prove()
        ms.append(p.last_prove_ms)
    out = {"workload": "SHA256_64, 1 instance, 92723 gates", "ms_per_proof": statistics.median(ms),
           "gates_per_s": 92723 / (statistics.median(ms) * 1e-3), "launches": p.last_prove_launches}
    p.close()
    return out


# ---------------------------------------------------------------------------------------- CPU reference
def _ref_worker(k_inst, reps, out_q):
    """one process: the unmodified reference prover (oracle/_ref/libref_gkr.so) -- or the C oracle port when
    the compiled reference is unavailable -- on SHA256_64 x k_inst; reports seconds per proof."""
    devnull = os.open(os.devnull, os.O_WRONLY)
    os.dup2(devnull, 2)  # the reference prints per-layer progress lines on stderr
    B, O = entry.binding(), entry.oracle()
    flat = load_sha(B).replicate(k_inst).expand().flat()
    kind = "reference" if O.ref_available() else "port"
    secs = []
    for _ in range(reps):
        if kind == "reference":
            _, ps, es = O.ref_prove(flat)
            secs.append(ps + es)
        else:
            _, _, s = O.OracleCircuit(flat).prove()
            secs.append(s)
    out_q.put((kind, secs))


def cpu_run(k_inst, procs, reps):
    import multiprocessing as mp
    ctx = mp.get_context("fork")
    q = ctx.Queue()
    ps = [ctx.Process(target=_ref_worker, args=(k_inst, reps, q)) for _ in range(procs)]
    t0 = time.time()
    for p in ps:
        p.start()
    res = [q.get() for _ in ps]
    for p in ps:
        p.join()
    wall = time.time() - t0
    return res, wall


def cpu_baseline_sample(args):
    """bounded sample for the `cpu_baseline` object of our line: 1 core, SHA256_64 x 16 instances, 1 proof"""
    k = 16
    res, _ = cpu_run(k, 1, 1)
    kind, secs = res[0]
    g = 92723 * k
    return {"value": g / secs[0], "unit": UNIT, "cores": 1, "kind": kind,
            "sample": f"SHA256_64 x {k} instances ({g} gates), one proof, prover methods + evaluate timed (the reference's `Prove Time`)",
            "seconds": secs[0]}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    procs = max(1, min(os.cpu_count() or 1, 64))
    k = 8
    g = 92723 * k
    # per step: every host core proves an independent SHA256_64 x k batch (the reference is single-threaded)
    total = args.warmup + args.steps
    res, _ = cpu_run(k, procs, total)
    kind = res[0][0]
    per_step = [max(r[1][s] for r in res) for s in range(total)][args.warmup:]
    sec = sum(per_step) / len(per_step)
    value = procs * g / sec
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u64 limbs (F_p^2, p=2^61-1)", "data": "synthetic",
        "config": {"workload": f"SHA256_64 x {k} instances per process ({g} gates), {procs} independent processes "
                               f"(bounded sample of the SHA256_64 x 1024 workload; the reference prover is single-threaded)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": procs, "kind": kind,
                         "sample": f"{procs} processes x SHA256_64 x {k} instances per step, max time over processes"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--instances", type=int, default=1024, help="SHA256_64 instances per GPU")
    ap.add_argument("--no-extras", action="store_true", help="skip the C1/C2 side measurements and the CPU sample")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    # The contract is ONE JSON line on stdout. Libraries write there too (NCCL prints its version banner on stdout at
    # communicator creation): keep the real stdout aside, point fd 1 at stderr for the run, print the line at the end.
    sys.stdout.flush()
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    global _OUT
    _OUT = real_stdout
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
    real_stdout.flush()


_OUT = None


def emit(line):
    out = _OUT or sys.stdout
    out.write(json.dumps(line, default=_jd) + "\n")
    out.flush()


if __name__ == "__main__":
    main()
