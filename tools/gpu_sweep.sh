timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_fullsize.py -m gpu -x -q -k "hamming or sharded or c2_scan or c2_few or c2_shard" 2>&1 | tail -3
for U in 1.25e6 1e7; do for s in 4096 8192 16384; do echo -n "U=$U seed=$s: "; SB_SEED_ROWS=$s timeout 120 python tools/scan_bench.py $U 4096 2>&1 | grep "variant 0" | sed 's/.*scan kernel//'; done; done
python - <<'PY'
import torch, sys
sys.path.insert(0,'.')
from smqtk_indexing_b200 import device as D, _lib
db = torch.randint(-2**31, 2**31-1, (1_250_000, 8), dtype=torch.int32, device="cuda")
q = torch.randint(-2**31, 2**31-1, (4096, 8), dtype=torch.int32, device="cuda")
for _ in range(3): D.hamming_scan_keys(db, q, 10)
_lib.profile_fetch(); _lib.profile_enable(True)
for _ in range(5): D.hamming_scan_keys(db, q, 10)
_lib.profile_enable(False)
agg={}
for n,ms in _lib.profile_fetch(): agg.setdefault(n,[]).append(ms)
print({k: sum(v)/len(v) for k,v in agg.items()})
PY
