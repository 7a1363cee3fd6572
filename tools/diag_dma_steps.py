"""Per-rank, per-step view of the DMA-shipped sharded join (torchrun, one process per GPU).

    torchrun --nproc-per-node 8 --master-addr 127.0.0.1 tools/diag_dma_steps.py [log2n=27] [steps=12] [free]

With `free` the steps run back to back as in bench.py's timed loop (no barrier or synchronize between them; a step's time
is the distance between the events recorded in front of consecutive steps).

bench.py reports one average over the max of all ranks; this prints, for every rank, the device time of every step, the
kernels the library launched in it (a fixed-capacity overflow shows up as extra launches: the second passes are redone)
and the plan bits (4 / 8 = build / probe slot took the histogram-free second pass), so that a slow average can be told
apart from one slow rank or a few slow steps.  Environment switches (RHJ_NO_SHARD_OPT2, RHJ_SHARD_OPT2_WORLD, ...) apply.
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from radixhashjoin_b200 import RadixHashJoin, workloads as W
from radixhashjoin_b200.distributed import DmaShardedJoin


def main():
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 27
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 12
    torch.cuda.set_device(lr)
    dev = f"cuda:{lr}"
    dist.init_process_group("nccl", device_id=torch.device(dev))
    n_local = 1 << log2n
    gbits = log2n + (world.bit_length() - 1)
    eng = RadixHashJoin(lr)
    w = W.uniform_unique(log2n, dev, row_offset=rank * n_local, log2_global=gbits)
    slack = int(n_local * 1.05) + 4096
    out = torch.empty((slack, 2), dtype=torch.int64, device=dev)
    dj = DmaShardedJoin(eng, world, rank, n_local * world, n_local * world, n_local, slack)
    free = len(sys.argv) > 3 and sys.argv[3] == "free"
    rows, evs = [], []
    dist.barrier()
    torch.cuda.synchronize()
    for it in range(steps):
        if not free:
            dist.barrier()
            torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True)
        e0.record()
        evs.append(e0)
        pairs, count, (nB, nP) = dj.step(w.R, w.S, out)
        plan = eng.last_plan()
        rows.append({"step": it, "launches": plan["kernel_launches"], "bits": plan["optimistic_pass1"], "count": int(count),
                     "recv": [int(nB), int(nP)]})
        if not free:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            torch.cuda.synchronize()
            rows[-1]["ms"] = round(e0.elapsed_time(e1), 3)
    e_end = torch.cuda.Event(enable_timing=True)
    e_end.record()
    evs.append(e_end)
    torch.cuda.synchronize()
    if free:
        for it in range(steps):
            rows[it]["ms"] = round(evs[it].elapsed_time(evs[it + 1]), 3)
    gathered = [None] * world
    dist.all_gather_object(gathered, rows)
    if rank == 0:
        print("step " + " ".join(f"r{r:<9d}" for r in range(world)) + "  max")
        for it in range(steps):
            cells = [gathered[r][it] for r in range(world)]
            print(f"{it:4d} " + " ".join(f"{c['ms']:6.2f}/{c['launches']:2d}/{c['bits']:<2d}".ljust(10) for c in cells) +
                  f"  {max(c['ms'] for c in cells):6.2f}")
        print("cells: ms / kernel launches / plan bits; " + ("steps back to back" if free else "steps separated by a barrier"))
        print(json.dumps({"world": world, "log2n": log2n, "per_rank": gathered}))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
