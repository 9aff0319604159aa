"""BASELINE config 3 on N GPUs (torchrun): exact flat L2 kNN over N x 12.5M x 128-d fp32 rows, k=100.
Each rank holds its shard; a step = ShardedFlatL2Index.query for one batch (local exact top-k, all-gather,
merge).  Device-timed, max over ranks; a few queries are verified against a float64 brute force over ALL
shards (each rank checks its own rows, the minima are all-reduced)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from smqtk_indexing_b200.distributed import ShardedFlatL2Index

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rows = int(float(sys.argv[1])) if len(sys.argv) > 1 else 12_500_000
k = int(sys.argv[2]) if len(sys.argv) > 2 else 100
g = torch.Generator(device="cuda"); g.manual_seed(100 + rank)
x = torch.rand((rows, 128), device="cuda", generator=g)
idx = ShardedFlatL2Index()
idx.build(x)
gq = torch.Generator(device="cuda"); gq.manual_seed(7)
q_all = torch.rand((4096, 128), device="cuda", generator=gq)
for Q in (4096, 256, 32, 1):
    q = q_all[:Q].contiguous()
    for _ in range(2):
        r, d = idx.query(q, k)
    dist.barrier(); torch.cuda.synchronize()
    steps = 5
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        r, d = idx.query(q, k)
    e1.record(); torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / steps], device="cuda")
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        print("flat L2 %d x %d rows x 128-d, k=%d, Q=%d: %.3f ms/batch  %.1f queries/s (max over %d ranks)"
              % (world, rows, k, Q, ms.item(), Q / ms.item() * 1e3, world), flush=True)
    if Q == 4096:
        # float64 check of 2 queries: global k-th distance from per-shard brute force
        for qi in (0, 4095):
            dd = ((x.double() - q[qi].double()[None, :]) ** 2).sum(1).sqrt() if rows <= 2_000_000 else None
            if dd is None:
                acc = []
                for s0 in range(0, rows, 2_000_000):
                    xs = x[s0:s0 + 2_000_000].double()
                    acc.append(((xs - q[qi].double()[None, :]) ** 2).sum(1).sqrt().topk(k, largest=False).values)
                    del xs
                loc = torch.cat(acc).topk(k, largest=False).values
            else:
                loc = dd.topk(k, largest=False).values
            allv = [torch.empty_like(loc) for _ in range(world)]
            dist.all_gather(allv, loc)
            want = torch.cat(allv).topk(k, largest=False).values
            assert torch.allclose(d[qi], want, rtol=1e-9, atol=0), (qi, (d[qi] - want).abs().max().item())
        if rank == 0:
            print("   verified 2 queries against a float64 brute force over all shards (rtol 1e-9)", flush=True)
dist.destroy_process_group()
