"""Step-by-step check of the peer-memory path on N GPUs (torchrun): IPC mapping, sb_rerank_peer, sharded index
in every mode, CUDA-graph replay.  Prints after every step (with a device synchronise) so a fault is attributable."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)


def say(msg):
    torch.cuda.synchronize()
    dist.barrier()
    if rank == 0:
        print(msg, flush=True)


from smqtk_indexing_b200 import device, peer  # noqa: E402
from smqtk_indexing_b200.distributed import ShardedLshIndex  # noqa: E402
from smqtk_indexing_b200.impls.lsh_functor.itq import ItqFunctor  # noqa: E402

n_local, D = 5000, 64
x = torch.arange(n_local * D, device=dev, dtype=torch.float32).reshape(n_local, D) / (n_local * D) + rank
bounds = [r * n_local for r in range(world + 1)]
say("alloc_conf=%r" % os.environ.get("PYTORCH_CUDA_ALLOC_CONF"))
sh = peer.share_rows(x, bounds)
say("share_rows ok: ptrs=%s aligned=%s ld=%d" % ([hex(p) for p in sh.ptrs], sh.aligned16, sh.ld))
# one query (zeros), candidates = row 3 of every shard
q = torch.zeros((1, D), device=dev)
cand = torch.tensor([r * n_local + 3 for r in range(world)], dtype=torch.int64, device=dev)
off = torch.tensor([0, world], dtype=torch.int64, device=dev)
d = device.rerank_peer(sh, q, cand, off, "euclidean")
want = [float(np.sqrt(((np.arange(3 * D, 4 * D, dtype=np.float64).astype(np.float32) / np.float32(n_local * D) + np.float32(r)).astype(np.float64) ** 2).sum())) for r in range(world)]
say("rerank_peer: got %s want %s" % (d.tolist(), want))

rng = np.random.RandomState(0)
xs = rng.rand(world * 20000, D).astype(np.float32)
f = ItqFunctor(bit_length=32, itq_iterations=3, random_seed=0)
f.fit_matrix(xs[:5000])
say("fit ok")
xl = torch.from_numpy(xs[rank * 20000:(rank + 1) * 20000]).to(dev)
qs = torch.from_numpy(rng.rand(1024, D).astype(np.float32)).to(dev)
ref = None
for rerank in ("allreduce", "peer"):
    for part in ("rows", "queries"):
        for graph in (False, True):
            idx = ShardedLshIndex(f, "euclidean", scan_partition=part, rerank=rerank, graph=graph)
            idx.build(xl)
            say("build %s/%s/graph=%s ok (peers=%s err=%s)" % (rerank, part, graph, idx.peers is not None, idx.peer_error))
            for it in range(4):
                rows, dd = idx.query(qs, 10)
                say("  query %d ok" % it)
            if ref is None:
                ref = (rows.clone(), dd.clone())
            else:
                assert torch.equal(rows, ref[0]) and torch.equal(dd, ref[1]), "mismatch in %s/%s/%s" % (rerank, part, graph)
fs = ItqFunctor(bit_length=32, itq_iterations=3, random_seed=0)
fs.fit_matrix(xl[:2500], want_codes=False, group=dist.group.WORLD)
say("sharded fit ok")
say("ALL OK")
torch.cuda.synchronize(); dist.barrier(); os._exit(0)
