import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, torch.distributed as dist
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
MODE = os.environ.get("MODE", "local")
def say(msg):
    torch.cuda.synchronize(); dist.barrier()
    print("[r%d] %s" % (rank, msg), flush=True)
from smqtk_indexing_b200 import device, peer, _lib
n_local, D = 5000, 64
x = torch.arange(n_local * D, device=dev, dtype=torch.float32).reshape(n_local, D) / (n_local * D) + rank
bounds = [r * n_local for r in range(world + 1)]
other = (rank + 1) % world
say("can_access_peer(%d->%d)=%s" % (lr, other, torch.cuda.can_device_access_peer(lr, other)))
sh = peer.share_rows(x, bounds)
say("ptrs=%s" % [hex(p) for p in sh.ptrs])
if MODE == "torchcopy":
    st = None
    t = torch.empty(0, dtype=torch.float32, device=st.device).set_(st, x.storage_offset(), (n_local, D))
    say("peer tensor device=%s ptr=%s" % (t.device, hex(t.data_ptr())))
    loc = t.to(dev)
    say("copied: first=%g expect=%g" % (float(loc[0, 0]), float(other)))
q = torch.zeros((1, D), device=dev)
rows = {"local": [rank * n_local + 3], "remote": [other * n_local + 3], "torchcopy": [other * n_local + 3]}[MODE]
cand = torch.tensor(rows, dtype=torch.int64, device=dev)
off = torch.tensor([0, len(rows)], dtype=torch.int64, device=dev)
d = device.rerank_peer(sh, q, cand, off, "euclidean")
say("rerank_peer(%s): %s" % (MODE, d.tolist()))
torch.cuda.synchronize(); dist.barrier(); os._exit(0)
