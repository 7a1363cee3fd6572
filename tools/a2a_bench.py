#!/usr/bin/env python
"""All-to-all microbench of the hand-written copy kernel (k_pipe_ship) alone.

    torchrun --nproc-per-node N tools/a2a_bench.py [--log2n 27] [--chunks 4] [--ctas 48] [--reps 5]

Every rank partitions its two local relations (pass 1 of the pipelined exchange, untimed), then -- between a
barrier and CUDA events on the copy stream -- ships all chunks of both relations to their destinations, nothing
else running on the GPU.  The step is then completed normally (pass 2, join) so that the flag protocol stays in
step.  One JSON line per run (rank 0): GB/s per GPU and direction = remote bytes sent / max-over-ranks time.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def one(args, torch, dist, eng, R, S, out, rank, world, dev, n, ctas, stage_kb, stages):
    from radixhashjoin_b200.distributed import PipeShardedJoin
    pj = PipeShardedJoin(eng, world, rank, n * world, n * world, n, n, chunks=args.chunks)
    rels = (R, S)
    times = []
    for rep in range(args.reps + 2):
        pj.epoch += 1
        eng.pipe_begin(pj.epoch)
        for rel in pj.order:
            rows = pj.chunk_rows[rel]
            for c in range(pj.chunks):
                eng.pipe_pass1(rel, c, rels[rel][c * rows:(c + 1) * rows])
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        cs = pj.copy_stream
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(cs)
        for rel in pj.order:
            for c in range(pj.chunks):
                eng.pipe_ship(rel, c, stream=cs)
        e1.record(cs)
        for rel in pj.order:
            for c in range(pj.chunks):
                eng.pipe_pass2(rel, c)
        eng.pipe_post()
        pairs, count, status = eng.pipe_join(out)
        torch.cuda.synchronize()
        assert status == 0, status
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rep >= 2:
            times.append(float(t.item()))
    cnt = torch.tensor([count], dtype=torch.int64, device=dev)
    dist.all_reduce(cnt)
    if rank == 0:
        sent = 2 * n * 16 * (world - 1) / world
        ms = sorted(times)[len(times) // 2]
        print(json.dumps({"bench": "k_pipe_ship all-to-all", "n_gpus": world, "tuples_per_relation_per_gpu": n, "chunks": args.chunks,
                          "ship_ctas": ctas, "stage_kb": stage_kb, "stages": stages, "remote_bytes_sent_per_gpu": sent, "ms_median": ms,
                          "ms_all": [round(x, 3) for x in times], "GBps_per_gpu_per_direction": sent / ms / 1e6,
                          "total_pairs": int(cnt.item()), "expected_pairs": n * world}), flush=True)
    del pj
    torch.cuda.empty_cache()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log2n", type=int, default=27)
    ap.add_argument("--chunks", type=int, default=4)
    ap.add_argument("--ctas", type=int, default=0)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--sweep", default="", help="comma list of ctas:stage_kb:stages to run one after the other")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    from radixhashjoin_b200 import RadixHashJoin
    from radixhashjoin_b200 import workloads as W
    from radixhashjoin_b200.distributed import PipeShardedJoin

    rank, world, local_rank = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local_rank)
    dev = f"cuda:{local_rank}"
    dist.init_process_group("nccl", device_id=torch.device(dev))
    eng = RadixHashJoin(local_rank)
    n = 1 << args.log2n
    gbits = args.log2n + (world.bit_length() - 1)
    w = W.uniform_unique(args.log2n, dev, row_offset=rank * n, log2_global=gbits)
    R, S = w.R, w.S
    out = torch.empty((int(n * 1.05) + 4096, 2), dtype=torch.int64, device=dev)
    configs = [tuple(int(x) for x in c.split(":")) for c in args.sweep.split(",") if c] or [(args.ctas or 48, 8, 8)]
    for ctas, stage_kb, stages in configs:
        os.environ["RHJ_PIPE_SHIP_CTAS"], os.environ["RHJ_PIPE_STAGE_KB"], os.environ["RHJ_PIPE_STAGES"] = str(ctas), str(stage_kb), str(stages)
        one(args, torch, dist, eng, R, S, out, rank, world, dev, n, ctas, stage_kb, stages)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
