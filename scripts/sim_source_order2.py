"""Follow-up to sim_source_order.py: lexicographic orders on the top-k neighbours by degree.
Same counting (rows gathered, bytes per useful value in both passes) on BA(N, 8), CPU only.
    python scripts/sim_source_order2.py [N]"""
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

# landmark signatures: hop distances to the K highest-degree vertices
hubs = np.argsort(-deg, kind='stable')[:64]
DH = shortest_path(A, method='D', unweighted=True, indices=hubs).astype(np.int64).T  # [N][64]
top1 = np.zeros(N, np.int64)
for v in range(N):
    nb = indices[indptr[v]:indptr[v+1]]
    top1[v] = nb[np.lexsort((nb, -deg[nb]))][0]
rng = np.random.default_rng(1)
pools = [("pool 65536 random" if N >= 65536 * 2 else "pool N/4 random", rng.permutation(N)[:min(65536, N // 4)]), ("all sources", np.arange(N))]
if os.environ.get("POOL"):
    pools = [p for p in pools if os.environ["POOL"] in p[0]]
for label, pool in pools:
    n = len(pool)
    centers = [int(0.15 * n), int(0.4 * n), int(0.65 * n), int(0.9 * n)]
    def groups(order):
        return [order[max(0, min(n - 128, c - 64)):][:128] for c in centers]
    print(label)
    o6 = pool[np.lexsort((-key[pool], top1[pool]))]
    report("  V6 (hub1, key)  [ships]", groups(o6))
    for K in (8, 16, 32, 64):
        sig = DH[pool][:, :K]
        o = pool[np.lexsort(tuple([-key[pool]] + [sig[:, j] for j in range(K - 1, -1, -1)]))]
        report(f"  V10 landmark signature, K={K}, then key", groups(o))
