"""CPU simulation: groups of 128 vs 256 sources (sorted by the two-hop key) on BA(200k, 8) -- rows gathered,
level-mask look-ups, queue entries and sector bytes PER SOURCE.  Result (third session of round 1):
  BG=128: rows 44.5k  look-ups 70.7k  queue entries 4.5k  bwd 14.13 MB  fwd 10.12 MB
  BG=256: rows 25.2k  look-ups 38.2k  queue entries 2.4k  bwd 14.09 MB  fwd 10.07 MB
i.e. 43-47 % less instruction work at equal DRAM bytes.  python scripts/sim_group_width.py"""
import sys, numpy as np
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
import scipy.sparse as sp
from scipy.sparse.csgraph import shortest_path
from rustworkx_b200 import _lib
N = 200_000
cap = 8 * N
src = np.zeros(cap, np.uint32); dst = np.zeros(cap, np.uint32)
m = _lib.gen_lib().rxc_gen_barabasi_albert(N, 8, 1, _lib.p_u32(src), _lib.p_u32(dst), cap)
src = src[:m].astype(np.int64); dst = dst[:m].astype(np.int64)
A = sp.coo_matrix((np.ones(m), (src, dst)), shape=(N, N)).tocsr(); A = A + A.T
deg = np.asarray(A.sum(axis=1)).ravel()
key = np.asarray(A @ deg).ravel()
u = np.concatenate([src, dst]); w = np.concatenate([dst, src])

def eval_group(idx, BG):
    """sorted position i -> lane i // K, k = i % K  (K = BG/32 sources per lane, consecutive)"""
    K = BG // 32
    D = shortest_path(A, method='D', unweighted=True, indices=idx).astype(np.int8).T.copy()  # [N, BG] in sorted order
    useful = rows = bwd32 = fwd32 = lookups = 0
    ent = sum(int(((D == L).any(axis=1)).sum()) for L in range(0, 10))
    chunk = 200_000
    for i in range(0, u.size, chunk):
        Du = D[u[i:i+chunk]]; Dw = D[w[i:i+chunk]]
        dag = (Dw == Du + 1)
        for L in range(0, 8):
            atL = (Du == L)
            lookups += int(atL.any(axis=1).sum())          # backward: one P look-up per (entry, arc)
            mL = dag & atL
            anyrow = mL.any(axis=1)
            if not anyrow.any(): continue
            mL = mL[anyrow]
            rows += mL.shape[0]; useful += int(mL.sum())
            sec = mL.reshape(-1, BG // 4, 4).any(axis=2)   # f64: 4 consecutive sources per 32-byte sector
            bwd32 += int(sec.sum()) * 32
            fwd32 += int(mL.reshape(-1, BG // 8, 8).any(axis=2).sum()) * 32  # u32: 8 consecutive per sector
    return np.array([rows, useful, bwd32, fwd32, ent, lookups], dtype=np.float64)

order = np.argsort(-key, kind='stable')
for BG in (128, 256):
    r = sum(eval_group(order[c:c+BG], BG) for c in (int(0.3*N), int(0.6*N)))
    per_src = 2 * BG
    print(f"BG={BG}: per source: rows {r[0]/per_src/1e3:7.1f}k  look-ups {r[5]/per_src/1e3:7.1f}k  queue entries {r[4]/per_src/1e3:6.1f}k  bwd bytes {r[2]/per_src/1e6:6.2f} MB ({r[2]/r[1]:4.1f} B/useful8)  fwd bytes {r[3]/per_src/1e6:6.2f} MB ({r[3]/r[1]:4.1f} B/useful4)", flush=True)
