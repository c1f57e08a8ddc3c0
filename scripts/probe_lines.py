"""What the betweenness kernels ask of the memory system, counted by the RXC_PROBE build
(make -C rustworkx_b200/csrc variant VAR=probe DEFS=-DRXC_PROBE=1): level-mask look-ups, gathered rows and
the 32-byte sectors / 128-byte lines of those rows, per source, on the bench graph.
python scripts/probe_lines.py [S] > profiles/r02_probe_lines.json   (the probe build is never timed)"""
import json, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
S = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
child = f"""
import sys; sys.path.insert(0, {ROOT!r})
import numpy as np, rustworkx_b200 as rx
g = rx.barabasi_albert_graph(1_000_000, 8, seed=1)
dg = rx.device_snapshot(g)
plan = np.random.default_rng(12345).permutation(g.num_nodes()).astype(np.uint32)
dg.betweenness_partial(plan[:{S}], False)
st = dg.stats()
print("STATS", st["te"], st["te_dag"], st["v"], st["levels"], st["batches"])
"""
env = dict(os.environ, RXCUDA_LIB=os.path.join(ROOT, "rustworkx_b200", "librxcuda_probe.so"))
r = subprocess.run([sys.executable, "-c", child], env=env, capture_output=True, text=True)
probe = [l for l in r.stderr.splitlines() if l.startswith("[rxc probe]")]
stats = [l for l in r.stdout.splitlines() if l.startswith("STATS")]
if not probe or not stats:
    sys.stderr.write(r.stdout + r.stderr)
    sys.exit(1)
nums = [int(x) for x in re.findall(r"(?:lookups|rows|sectors|lines128) (\d+)", probe[-1])]
te, dag, v, levels, batches = (int(x) for x in stats[-1].split()[1:])
keys = ("lookups", "rows", "sectors32", "lines128")
out = {"workload": f"betweenness_partial, barabasi_albert_graph(1000000, 8, seed=1), first {S} sources of the bench plan",
       "sources": S, "te": te, "te_dag": dag, "v": v, "levels": levels, "batches": batches,
       "bc_forward_kernel": {k: nums[i] / S for i, k in enumerate(keys)},
       "bc_backward_kernel": {k: nums[4 + i] / S for i, k in enumerate(keys)},
       "unit": "per source; a look-up reads one 32-byte level-mask entry (one 128-byte line request), rows / sectors32 / "
               "lines128 are the gathered rows and the sectors and lines their copies request"}
print(json.dumps(out, indent=1))
