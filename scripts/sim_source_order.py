"""CPU simulation of how the ORDER of betweenness sources changes the traffic of the row gathers.

No GPU needed (scipy BFS).  For a BA(N, 8) graph it forms groups of 128 sources under several
orderings, lays them out as the kernels do (slot 32k + l = source k of lane l, consecutive sources share
a lane's 32-byte sector) and counts, over all (vertex, level) entries and arcs, the rows gathered, the
useful (arc, source) pairs and the 32-byte sectors touched in the f64 coefficient rows (dependency
sweep) and the uint32 sigma rows (forward pass), plus the level-queue entries per vertex.

    python scripts/sim_source_order.py [N]

Results for N = 200 000 (third session of round 1; "rows" per group of 128 sources):
  random order                    rows 6.90M   bwd 22.1 B per useful 8 B   fwd 16.2 B per useful 4 B
  pool 1024 random  V1 key        rows 5.67M   bwd 20.4                    fwd 14.5      <- shipped (bc_order)
  all sources       V1 key        rows 5.61M   bwd 20.4                    fwd 14.4
  all sources       V2 BFS rank   rows 4.60M   bwd 15.7                    fwd 10.6
  all sources       V6 dom. hub   rows 4.81M   bwd 16.3                    fwd 11.1      <- next: full runs
"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import scipy.sparse as sp
from scipy.sparse.csgraph import shortest_path
from rustworkx_b200 import _lib
N = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
cap = 8 * N
src = np.zeros(cap, np.uint32); dst = np.zeros(cap, np.uint32)
m = _lib.gen_lib().rxc_gen_barabasi_albert(N, 8, 1, _lib.p_u32(src), _lib.p_u32(dst), cap)
src = src[:m].astype(np.int64); dst = dst[:m].astype(np.int64)
A = sp.coo_matrix((np.ones(m), (src, dst)), shape=(N, N)).tocsr(); A = A + A.T
deg = np.asarray(A.sum(axis=1)).ravel()
key = np.asarray(A @ deg).ravel()
u = np.concatenate([src, dst]); w = np.concatenate([dst, src])
indptr, indices = A.indptr, A.indices

def eval_group(idx):
    slot_src = np.empty(128, np.int64)
    for i in range(128):
        slot_src[32 * (i % 4) + i // 4] = idx[i]
    D = shortest_path(A, method='D', unweighted=True, indices=slot_src).astype(np.int8).T.copy()
    useful = rows = bwd32 = fwd32 = 0
    ent = sum(int(((D == L).any(axis=1)).sum()) for L in range(0, 10))
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
            sec = mL.reshape(-1, 4, 32).any(axis=1)
            bwd32 += int(sec.sum()) * 32
            fwd32 += int(sec.reshape(-1, 16, 2).any(axis=2).sum()) * 32
    return np.array([rows, useful, bwd32, fwd32, ent], dtype=np.float64)

def report(label, groups):
    r = sum(eval_group(g) for g in groups)
    print(f"{label:52s} rows {r[0]/len(groups)/1e6:6.2f}M/group  useful/row {r[1]/r[0]:5.1f}  bwd {r[2]/r[1]:5.1f} B/useful8  fwd {r[3]/r[1]:5.1f} B/useful4  entries/vertex {r[4]/len(groups)/N:4.2f}", flush=True)

root = int(np.argmax(deg))
lvl = shortest_path(A, method='D', unweighted=True, indices=[root])[0].astype(np.int64)
# parent = min-id neighbour on the previous level; domhub = max-degree neighbour
parent = np.full(N, -1, np.int64); domhub = np.zeros(N, np.int64)
for v in range(N):
    nb = indices[indptr[v]:indptr[v+1]]
    domhub[v] = nb[np.argmax(deg[nb])]
    if v != root:
        cand = nb[lvl[nb] == lvl[v] - 1]
        parent[v] = cand.min()
# exact BFS rank (queue order, neighbours in CSR order)
rank = np.full(N, -1, np.int64); q = [root]; rank[root] = 0; c = 1; head = 0
while head < len(q):
    x = q[head]; head += 1
    for y in indices[indptr[x]:indptr[x+1]]:
        if rank[y] < 0:
            rank[y] = c; c += 1; q.append(y)

rng = np.random.default_rng(1)
for label, pool in (("pool 1024 random", rng.permutation(N)[:1024]), ("pool 8192 random", rng.permutation(N)[:8192]), ("all sources", np.arange(N))):
    n = len(pool)
    centers = [int(0.2 * n), int(0.5 * n), int(0.8 * n)]
    def groups(order):
        return [order[max(0, min(n - 128, c - 64)):][:128] for c in centers]
    o1 = pool[np.argsort(-key[pool], kind='stable')]
    o2 = pool[np.argsort(rank[pool], kind='stable')]
    o5 = pool[np.lexsort((-key[pool], parent[pool], lvl[pool]))]
    o6 = pool[np.lexsort((-key[pool], domhub[pool]))]
    o7 = pool[np.lexsort((-key[pool], deg[domhub[pool]] * -1))]   # by degree of dominant hub, then key
    print(label)
    report("  V1 key (now)", groups(o1))
    report("  V2 BFS rank from the top hub", groups(o2))
    report("  V5 (level, min-id parent, key)", groups(o5))
    report("  V6 (dominant hub id, key)", groups(o6))
    report("  V7 (dominant hub degree desc, key)", groups(o7))
