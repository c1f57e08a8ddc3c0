"""Source order under the LINE cost model (the memory system counts 128-byte lines):
lines per gathered row in both passes for the order that ships and for re-arrangements of the 128
sources INSIDE a group.  BA(N, 8), CPU only.   python scripts/sim_source_order3.py [N]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import scipy.sparse as sp
from scipy.sparse.csgraph import shortest_path
from rustworkx_b200 import _lib
N = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
cap = 8 * N
src = np.zeros(cap, np.uint32); dst = np.zeros(cap, np.uint32)
m = _lib.gen_lib().rxc_gen_barabasi_albert(N, 8, 1, _lib.p_u32(src), _lib.p_u32(dst), cap)
src = src[:m].astype(np.int64); dst = dst[:m].astype(np.int64)
A = sp.coo_matrix((np.ones(m), (src, dst)), shape=(N, N)).tocsr(); A = A + A.T
deg = np.asarray(A.sum(axis=1)).ravel()
key = np.asarray(A @ deg).ravel()
u = np.concatenate([src, dst]); w = np.concatenate([dst, src])
indptr, indices = A.indptr, A.indices

def eval_group(idx, D=None):
    """idx: 128 sources in SORTED position order (position i -> lane i // 4, slot i % 4)."""
    if D is None:
        D = shortest_path(A, method='D', unweighted=True, indices=idx).astype(np.int8).T.copy()  # [N][128] by position
    rows = useful = sec_b = line_b = line_f = half_f = 0
    chunk = 400_000
    for i in range(0, u.size, chunk):
        Du = D[u[i:i+chunk]]; Dw = D[w[i:i+chunk]]
        dag = (Dw == Du + 1)
        for L in range(0, 8):
            mL = dag & (Du == L)
            anyrow = mL.any(axis=1)
            if not anyrow.any(): continue
            mL = mL[anyrow]
            rows += mL.shape[0]; useful += int(mL.sum())
            lanes = mL.reshape(-1, 32, 4).any(axis=2)          # lane l = positions 4l..4l+3
            sec_b += int(lanes.sum())                          # f64: one 32-byte sector per lane
            line_b += int(lanes.reshape(-1, 8, 4).any(axis=2).sum())   # f64: 4 lanes per 128-byte line
            half_f += int(lanes.sum())                         # uint32: 16 bytes per lane
            line_f += int(lanes.reshape(-1, 4, 8).any(axis=2).sum())   # uint32: 8 lanes per line
    return np.array([rows, useful, sec_b, line_b, line_f], dtype=np.float64), D

def show(label, r, ng):
    print(f"{label:46s} rows {r[0]/ng/1e6:5.2f}M/group  useful/row {r[1]/r[0]:5.1f}  sweep: sectors/row {r[2]/r[0]:5.2f} lines/row {r[3]/r[0]:4.2f}"
          f"  forward: lines/row {r[4]/r[0]:4.2f}", flush=True)

top1 = np.zeros(N, np.int64)
for v in range(N):
    nb = indices[indptr[v]:indptr[v+1]]
    top1[v] = nb[np.lexsort((nb, -deg[nb]))][0]
rng = np.random.default_rng(1)
pool = rng.permutation(N)[:min(65536, N // 2)]
n = len(pool)
order = pool[np.lexsort((-key[pool], top1[pool]))]   # (hub1, key): ships
centers = [int(0.15 * n), int(0.5 * n), int(0.85 * n)]
groups = [order[max(0, min(n - 128, c - 64)):][:128] for c in centers]
hubs = np.argsort(-deg, kind='stable')[:16]
acc = {}
for g in groups:
    r, D = eval_group(g)
    acc.setdefault("ships: consecutive sorted sources share a lane", []).append(r)
    variants = {}
    variants["random inside the group"] = rng.permutation(128)
    variants["interleaved (neighbours in different lines)"] = np.arange(128).reshape(16, 8).T.ravel()
    meand = D.astype(np.float64).mean(axis=0)
    variants["by exact mean distance"] = np.argsort(meand, kind='stable')
    # greedy: grow lines of 16 sources by Hamming distance of their level vectors on a vertex sample
    S = D[rng.choice(N, size=4000, replace=False)].astype(np.int16)      # [4000][128]
    left = list(range(128)); perm = []
    while left:
        seed = left.pop(0); line = [seed]
        while len(line) < 16 and left:
            d = [(np.abs(S[:, line].astype(np.int32) - S[:, [c]]) != 0).sum() for c in left]
            line.append(left.pop(int(np.argmin(d))))
        perm += line
    variants["greedy lines by level-vector agreement"] = np.array(perm)
    for name, p in variants.items():
        r2, _ = eval_group(g[p], D[:, p].copy())
        acc.setdefault(name, []).append(r2)
for name, rs in acc.items():
    show(name, sum(rs), len(rs))
