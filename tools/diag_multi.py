"""Diagnostic for the sharded join: per-rank expected count/digest from the closed form."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from radixhashjoin_b200 import RadixHashJoin, workloads as W
from radixhashjoin_b200.distributed import ShardedJoin

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = f"cuda:{lr}"
dist.init_process_group("nccl", device_id=torch.device(dev))
log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 22
gbits = log2n + (world.bit_length() - 1)
n_local = 1 << log2n
eng = RadixHashJoin(lr)
w = W.uniform_unique(log2n, dev, row_offset=rank * n_local, log2_global=gbits)

def hash64_t(v):
    c = W._s64(0xd6e8feb86659fd93)
    v = v ^ W.lsr(v, 32); v = v * c; v = v ^ W.lsr(v, 32); v = v * c; v = v ^ W.lsr(v, 32)
    return v
def rank_of(v):
    bits = world.bit_length() - 1
    return W.lsr(hash64_t(v), 64 - bits) if bits else torch.zeros_like(v)

slack = int(n_local * 1.05) + 4096
recvR = torch.empty((slack, 2), dtype=torch.int64, device=dev); recvS = torch.empty_like(recvR); out = torch.empty_like(recvR)
sj = ShardedJoin(world, rank, lambda T: eng.shuffle_partition(T, world), lambda a, b: eng.join_device(a, b, out=out))
for it in range(3):
    gR, cR = eng.shuffle_partition(w.R, world)
    # grouped check: ranks of grouped values are sorted and counts right
    rk = rank_of(gR[:, 1])
    okg = bool((rk[1:] >= rk[:-1]).all().item()); cnts = torch.bincount(rk, minlength=world).tolist()
    pairs, count, (nR, nS) = sj.step(w.R, w.S, recvR, recvS)
    okR = bool((rank_of(recvR[:nR, 1]) == rank).all().item()); okS = bool((rank_of(recvS[:nS, 1]) == rank).all().item())
    # expected for this rank: probe rows j (global) whose value maps here
    n = 1 << gbits
    a, b = W._s64(W.A_MUL), W.B_ADD
    cnt_e, s_e, x_e = 0, 0, 0
    for lo in range(0, n, 1 << 24):
        j = torch.arange(lo, min(n, lo + (1 << 24)), dtype=torch.int64, device=dev)
        pi = (j * a + b) & (n - 1)
        val = W.mix64(pi + W.SEED)
        m = rank_of(val) == rank
        h = W.pair_hash(pi[m], j[m])
        cnt_e += int(m.sum().item()); s_e = (s_e + int(h.sum().item())) & ((1 << 64) - 1)
        if h.numel(): x_e ^= int(W._xor_reduce(h).item()) & ((1 << 64) - 1)
    dig = eng.pairs_digest(pairs)
    print(f"[rank {rank} it {it}] grouped_sorted={okg} counts_match={cnts == cR} recvR_ok={okR} recvS_ok={okS} nR={nR} nS={nS} "
          f"count={count} expected={cnt_e} digest_ok={(dig[1], dig[2]) == (s_e, x_e)} plan={eng.last_plan()}", flush=True)
dist.destroy_process_group()
