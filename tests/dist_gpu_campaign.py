"""Launched by torchrun (one process per GPU): a differential campaign on a context SHARDED over WORLD_SIZE GPUs. Seeded
random circuits -- random .pws DAGs (every gate type, operands from any earlier layer) and layered random add/mul circuits with
multi-source wiring -- replicated to K instances so that their sumcheck phases really shard (>= 2^12 table entries), proved as a
whole (vp_prove, NVLink exchange), from the rank's witness slice (vp_prove_local), every third case method by method (per-round
48-byte exchange), and verified by the sharded device verifier (honest + tampered): all compared with the C oracle on every rank.
usage: torchrun --nproc-per-node N tests/dist_gpu_campaign.py FIRST_SEED SECONDS"""
import importlib.util
import os
import random
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    B, O = entry.binding(), entry.oracle()
    spec = importlib.util.spec_from_file_location("mg", os.path.join(ROOT, "tests", "golden", "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    seed, budget = int(sys.argv[1]), float(sys.argv[2])
    t_end = time.time() + budget
    n = bad = sharded_total = 0
    while True:
        go = torch.tensor([1 if time.time() < t_end else 0], device="cuda")
        dist.broadcast(go, 0)                      # rank 0's clock decides: every rank runs the same cases
        if int(go.item()) == 0:
            break
        rng = random.Random(seed * 15485863)
        if seed % 3 == 2:
            circ = B.Circuit.random(rng.choice([3, 5, 8]), rng.choice([4, 7, 9]), seed)
            what = "layered"
        else:
            circ = B.Circuit.from_pws_text(mg.random_pws(seed, rng.choice([17, 64, 257, 1000]), rng.choice([30, 120, 500, 2000])))
            what = "pws"
        K = rng.choice([1, 7, 32, 100, 257, 600])
        if K > 1:
            circ = circ.replicate(K)
        flat = circ.expand() if K > 1 else circ
        oc = O.OracleCircuit(flat.flat())
        want, _, _ = oc.prove()
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt.copy_(torch.from_numpy(B.nccl_unique_id()))
        dist.broadcast(idt, 0)
        p = B.Prover(circ, device=local, rank=rank, world=world, nccl_id=idt.cpu().numpy())
        sharded = sum(t["sharded"] for i in range(1, circ.n_layers) for ph in (1, 2, 3) for t in B.shard_describe(circ, world, rank, i, ph)[:1])
        inp, ch = circ.inputs(), circ.draw_challenges()
        eq = lambda g: bool(len(g) == len(want) and (g["re"] == want["re"]).all() and (g["im"] == want["im"]).all())
        ok = eq(p.prove(inputs=inp, challenges=ch))
        lo, hi = p.input_range()
        s0 = circ.num_inputs // circ.instances
        ok_local = eq(p.prove_local(inp[lo * s0:hi * s0], ch))
        t = want.copy()
        k = rng.randrange(len(t))
        t[k]["re"] = (int(t[k]["re"]) + 1) % B.P
        ok_v = tuple(p.verify(want)) == (True, 0, 0) and tuple(p.verify(t)) == tuple(oc.verify(t))
        ok_i = eq(B.prove_interactive(p, circ)) if seed % 3 == 0 else True
        p.close()
        n += 1
        sharded_total += sharded
        if not (ok and ok_local and ok_v and ok_i):
            bad += 1
            print(f"[rank {rank}/{world}] MISMATCH seed {seed} ({what}, K {K}, gates {circ.total_gates}, sharded phases {sharded}): "
                  f"whole {ok} local {ok_local} verifier {ok_v} interactive {ok_i}", flush=True)
        seed += 1
    t = torch.tensor([bad], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    dist.destroy_process_group()
    if rank == 0:
        print(f"dist_gpu_campaign on {world} GPUs: {n} random circuits ({sharded_total} sharded phases), seeds {int(sys.argv[1])}..{seed - 1}: "
              f"{int(t.item())} mismatches (summed over ranks)", flush=True)
        print("DIST_CAMPAIGN", "PASS" if int(t.item()) == 0 else "FAIL", flush=True)
    sys.exit(0 if int(t.item()) == 0 else 1)


if __name__ == "__main__":
    main()
