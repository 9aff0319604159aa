import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import numpy as np, torch
from smqtk_indexing_b200 import device as dev, fit as fitops
import np_oracle as O
np.set_printoptions(linewidth=250, precision=2, suppress=True)

def run(x, bits):
    n, D = x.shape; b = bits.shape[1]
    W = 1 if b <= 32 else (2 if b <= 64 else (4 if b <= 128 else 8))
    codes = dev.codes_to_device(O.pack_codes(bits, W))
    xt = torch.from_numpy(x.astype(np.float32)).cuda()
    got = fitops.gram_bits_tc(codes, b, xt, None, None).cpu().numpy()
    ref = np.where(bits, 1.0, -1.0).T @ x.astype(np.float64)
    return got, ref

# 1. all bits set, X = distinct value per (row, d): G[m][d] = sum_r x[r][d]
n, D, b = 16, 32, 32
x = np.zeros((n, D)); x[:, :] = np.arange(D)[None, :] + 1          # column d has value d+1 in every row
got, ref = run(x, np.ones((n, b), bool))
print("test1 (col sums, expect row = 16*(d+1)):\n got[0] ", got[0], "\n got[5] ", got[5], "\n ref[0] ", ref[0])
# 2. single nonzero X element at (row 3, d 5) = 1, all bits set
x = np.zeros((n, D)); x[3, 5] = 1.0
got, ref = run(x, np.ones((n, b), bool))
print("test2 nonzero positions got:", np.argwhere(np.abs(got) > 1e-6)[:10].tolist(), " ref:", np.argwhere(np.abs(ref) > 1e-6)[:5].tolist())
# 3. X all ones in column 0; bits: only bit m=2 set in row 0, everything else clear -> G[m][0] = -16 except m=2: -14
x = np.zeros((n, D)); x[:, 0] = 1.0
bits = np.zeros((n, b), bool); bits[0, 2] = True
got, ref = run(x, bits)
print("test3 got[:,0]", got[:, 0], "\n      ref[:,0]", ref[:, 0])
# 4. row dependence: X[r][0] = r+1, bits: row r has bit (r % b) set
x = np.zeros((n, D)); x[:, 0] = np.arange(n) + 1
bits = np.zeros((n, b), bool); bits[np.arange(n), np.arange(n) % b] = True
got, ref = run(x, bits)
print("test4 got[:,0]", got[:, 0], "\n      ref[:,0]", ref[:, 0])
