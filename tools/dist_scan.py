"""development probe (torchrun): for SHA256_64 x K on WORLD_SIZE GPUs, compare every rank's transcript (resident proof and
vp_prove_local) with the same circuit proved on one GPU by rank 0, under kernel toggles.
usage: torchrun --nproc-per-node N tools/dist_scan.py K [K ...]"""
import hashlib, lzma, os, sys
import numpy as np, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
B = entry.binding()
with lzma.open(os.path.join(ROOT, "tests", "golden", "SHA256_64.pws.xz"), "rb") as f:
    sha = B.Circuit.from_pws_text(f.read())
sha_h = lambda tr: hashlib.sha256(np.ascontiguousarray(tr).tobytes()).hexdigest()[:12]
for K in [int(a) for a in sys.argv[1:]]:
    circ = sha.replicate(K)
    inp, ch = circ.inputs(), circ.draw_challenges()
    ref = [None]
    if rank == 0:
        os.environ["VP_ONE_LANE"] = "1"
        p1 = B.Prover(circ, device=local)
        ref[0] = sha_h(p1.prove(inputs=inp, challenges=ch))
        p1.close()
        del os.environ["VP_ONE_LANE"]
    dist.broadcast_object_list(ref, 0)
    for toggles in ({}, {"VP_LAYERS_TOP_DOWN": "1"}, {"VP_NO_EXTRAS": "1"}, {"VP_NCCL_EXCHANGE": "1"}):
        for k, v in toggles.items():
            os.environ[k] = v
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt.copy_(torch.from_numpy(B.nccl_unique_id()))
        dist.broadcast(idt, 0)
        p = B.Prover(circ, device=local, rank=rank, world=world, nccl_id=idt.cpu().numpy())
        for k in toggles:
            del os.environ[k]
        p.set_inputs(inp); p.set_challenges(ch)
        res = []
        for rep in range(2):
            p.prove(); res.append(sha_h(p.transcript()))
        lo, hi = p.input_range(); s0 = circ.num_inputs // K
        for rep in range(2):
            res.append(sha_h(p.prove_local(inp[lo * s0:hi * s0], ch)))
        res.append(sha_h(p.prove(inputs=inp, challenges=ch)))
        allr = [None] * world
        dist.all_gather_object(allr, res)
        if rank == 0:
            print(f"K={K} toggles={toggles} ref={ref[0]}")
            for r, x in enumerate(allr):
                print(f"   rank {r}: resident {[h == ref[0] for h in x[:2]]} local {[h == ref[0] for h in x[2:4]]} host_io {x[4] == ref[0]}", flush=True)
        p.close()
        dist.barrier()
dist.destroy_process_group()
