#!/usr/bin/env python
"""The pipelined multi-GPU exchange with all ranks EMULATED on one GPU (plain tensors stand in for the symmetric
blocks, the phases run rank after rank so that no kernel ever waits for another one).

    python tools/pipe_emulated_step.py --world 8 --log2n 24 --wire 12 --steps 2

Same kernels, same launch shapes and the same radix plan as a real N-GPU run with 2^log2n + 2^log2n tuples per rank --
this is what `ncu` is pointed at for the sharded kernels (k_scatter<kDigitShard, LIMIT>, k_pipe_ship / k_pipe_ship12,
k_pipe_arrive, the segmented LIMIT pass 2), and it gives their stand-alone durations, i.e. without NVLink traffic next
to them.  Prints per-phase CUDA-event times of rank 0 and verifies the union of the ranks' results (count + digest).
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--world", type=int, default=8)
    ap.add_argument("--log2n", type=int, default=24, help="tuples per relation per (emulated) rank")
    ap.add_argument("--chunks", type=int, default=4)
    ap.add_argument("--wire", type=int, default=12)
    ap.add_argument("--steps", type=int, default=2)
    args = ap.parse_args()
    import torch
    from radixhashjoin_b200 import RadixHashJoin
    from radixhashjoin_b200 import workloads as W

    dev = "cuda:0"
    world, n, C = args.world, 1 << args.log2n, args.chunks
    gbits = args.log2n + (world.bit_length() - 1)
    engines = [RadixHashJoin(0) for _ in range(world)]
    plan = engines[0].shard_plan(world * n, world * n, world)
    nbytes = engines[0].pipe_sym_bytes(plan, engines[0].pipe_cfg(plan, 0, C, n, n, wire_bytes=args.wire))
    syms = [torch.zeros(nbytes // 8, dtype=torch.int64, device=dev) for _ in range(world)]
    ptrs = [s.data_ptr() for s in syms]
    for r in range(world):
        engines[r].pipe_open(plan, engines[r].pipe_cfg(plan, r, C, n, n, ptrs, wire_bytes=args.wire))
    shards = []
    for r in range(world):
        w = W.uniform_unique(args.log2n, dev, row_offset=r * n, log2_global=gbits)
        shards.append((w.R, w.S))
    rows = (n + C - 1) // C
    out = torch.empty((int(n * 1.1) + 4096, 2), dtype=torch.int64, device=dev)
    exp = W.uniform_unique_global_digest(gbits, dev)

    def ev():
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        return e

    for step in range(1, args.steps + 1):
        t = {}
        for r in range(world):
            e0 = ev()
            engines[r].pipe_begin(step)
            e1 = ev()
            for rel in (0, 1):
                for c in range(C):
                    engines[r].pipe_pass1(rel, c, shards[r][rel][c * rows:(c + 1) * rows])
            e2 = ev()
            if r == 0:
                t["begin"], t["pass1"] = (e0, e1), (e1, e2)
        for r in range(world):
            e0 = ev()
            for rel in (0, 1):
                for c in range(C):
                    engines[r].pipe_ship(rel, c)
            if r == 0:
                t["ship"] = (e0, ev())
        for r in range(world):
            e0 = ev()
            for rel in (0, 1):
                for c in range(C):
                    engines[r].pipe_pass2(rel, c)
            engines[r].pipe_post()
            if r == 0:
                t["pass2"] = (e0, ev())
        cnt, s, x = 0, 0, 0
        for r in range(world):
            e0 = ev()
            pairs, k, status = engines[r].pipe_join(out)
            if r == 0:
                t["join"] = (e0, ev())
            assert status == 0, status
            d = engines[r].pairs_digest(pairs)
            cnt, s, x = cnt + d[0], (s + d[1]) & ((1 << 64) - 1), x ^ d[2]
        torch.cuda.synchronize()
        ms = {k: round(a.elapsed_time(b), 4) for k, (a, b) in t.items()}
        print(json.dumps({"tool": "pipe_emulated_step", "world": world, "tuples_per_relation_per_rank": n, "chunks": C,
                          "wire_bytes": args.wire, "radix_bits": [plan.bits_pass1, plan.bits_pass2], "step": step,
                          "rank0_phase_ms": ms, "verified": (cnt, s, x) == tuple(exp)}), flush=True)
    for e in engines:
        e.close()


if __name__ == "__main__":
    main()
