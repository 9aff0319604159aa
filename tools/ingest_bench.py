"""Scratch: incremental ingest rate (update_index_matrix batches into a growing index)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import smqtk_indexing_b200  # noqa
from smqtk_dataprovider.impls.key_value_store.memory import MemoryKeyValueStore
from smqtk_descriptors.impls.descriptor_set.memory import MemoryDescriptorSet
from smqtk_indexing_b200.impls.hash_index.linear import LinearHashIndex
from smqtk_indexing_b200.impls.lsh_functor.itq import ItqFunctor
from smqtk_indexing_b200.impls.nn_index.lsh import LSHNearestNeighborIndex

n0, nb, batches, D, b = 8_000_000, 250_000, 8, 512, 256
f = ItqFunctor(bit_length=b, itq_iterations=3, random_seed=0)
f.fit_matrix(torch.rand((200_000, D), device="cuda"), want_codes=False)
idx = LSHNearestNeighborIndex(f, MemoryDescriptorSet(), MemoryKeyValueStore(), LinearHashIndex(), "euclidean")
x0 = torch.rand((n0, D), device="cuda")
torch.cuda.synchronize(); t0 = time.perf_counter()
idx.build_index_matrix(x0)
torch.cuda.synchronize(); print("build %d rows: %.3f s" % (n0, time.perf_counter() - t0))
for i in range(batches):
    xb = torch.rand((nb, D), device="cuda")
    torch.cuda.synchronize(); t0 = time.perf_counter()
    idx.update_index_matrix(xb)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print("update +%d rows -> %d rows: %.1f ms (%.2f M rows/s)" % (nb, idx.count_rows(), dt * 1e3, nb / dt / 1e6))
gone = torch.randperm(idx.count_rows())[:500_000].tolist()
torch.cuda.synchronize(); t0 = time.perf_counter()
idx.remove_from_index_matrix(gone)
torch.cuda.synchronize(); dt = time.perf_counter() - t0
print("remove 500000 rows -> %d live: %.1f ms" % (idx.count_rows(), dt * 1e3))
r, d = idx.nn_batch(torch.rand((256, D), device="cuda"), 10)
print("query after ingest ok", r.shape)
