#!/usr/bin/env python
"""Multi-GPU check of the sharded C-ABI entry point (run under torchrun, one rank per GPU):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29540 tools/shard_check.py
Several steps over two alternating gather buffers with different inputs per step and rank; after every step each rank compares its whole
gather buffer with the records the ranks computed on their own (exchanged with torch.distributed).  Both exchange paths: registered
buffers (copy-engine peer copies) and unregistered ones (ncclAllGather).  Prints one JSON line on rank 0; exit code 1 on mismatch."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from cpi_b200 import preint, shard, synth

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
import datetime
dist.init_process_group("nccl", device_id=torch.device("cuda", local), timeout=datetime.timedelta(seconds=120))
comm = shard.Communicator()
report = {"world": world, "cases": []}
bad = 0
for model, dtype, n, ns in ((1, torch.float64, 1503, 60), (2, torch.float64, 700, 80), (1, torch.float32, 1503, 60)):
    rd = 290 if model == 1 else 308
    for registered in (True, False):
        gathers = [torch.zeros((world, n, rd), dtype=dtype, device="cuda") for _ in range(2)]
        torch.cuda.synchronize()
        mode = [comm.register(g) for g in gathers] if registered else [False, False]
        worst = 0.0
        for step in range(5):
            S, L = synth.make_windows(n, ns, first_window=(step * world + rank) * n)
            dS, dL = torch.from_numpy(S).to(dtype).cuda(), torch.from_numpy(L).to(dtype).cuda()
            g = gathers[step & 1]
            comm.step(model, dS, dL, synth.SIGMAS, 0, g, ns=ns)
            comm.wait()
            got = g.clone()
            torch.cuda.synchronize()                       # the product communicator is idle before torch's group runs a collective
            own = preint.preintegrate(model, dS, dL, synth.SIGMAS, 0, ns=ns)
            ref = torch.empty_like(g)
            dist.all_gather_into_tensor(ref, own)
            torch.cuda.synchronize()
            if not torch.equal(got, ref):
                bad += 1
                worst = max(worst, float((got.double() - ref.double()).abs().max()))
        torch.cuda.synchronize()
        comm.unregister()
        dist.barrier()
        torch.cuda.synchronize()
        report["cases"].append({"model": model, "dtype": str(dtype), "windows_per_rank": n, "registered": registered, "peer_copies": mode, "steps": 5,
                                "max_abs_diff_vs_own_results": worst})
t = torch.tensor([bad], device="cuda")
torch.cuda.synchronize()
dist.all_reduce(t)
report["mismatching_steps_over_all_ranks"] = int(t.item())
if rank == 0:
    print(json.dumps(report))
comm.close()
dist.destroy_process_group()
sys.exit(1 if int(t.item()) else 0)
