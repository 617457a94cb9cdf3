"""Determinism check of a kernel form on a deep circuit: run the same encrypted input several times, compare every wire
ciphertext with the first run (reference form = GPC_REF, default 8) and report the first level whose outputs differ."""
import os, sys, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bfhe_loader
B = bfhe_loader.load_package()
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
V = json.load(open(os.path.join(ROOT, "tests", "golden", "vectors.json")))
name = sys.argv[1] if len(sys.argv) > 1 else "md5"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
ctx = B.Context(B.STD128_OPT, B.GINX, 0)
ctx.keygen(1); ctx.btkeygen(2)
c = B.Circuit(ctx)
c.load_npz(os.path.join(ROOT, "tests", "golden", "circuits", name + ".npz"))
v = V[name]["vectors"][0]
def run(gpc):
    ctx.dbg_set_gates_per_cta(gpc)
    c.Reset(); c.setEncrypted(True)
    c.SetInput(v["inputs"], seed=7)
    out = c.Clock()[0]
    return out, c.download_slab()
runs = []
if os.environ.get("TEST_FIRST"):  # the form under test runs before anything else touched the GPU
    runs = [run(int(os.environ.get("GPC", "32"))) for _ in range(reps)]
ref_out, ref = run(int(os.environ.get("GPC_REF", "8")))
print("reference ok", ref_out == v["golden"], flush=True)
misc = c.plan_misc()
firsts = [c.level_plan(L, 0, 1)[1] for L in range(misc["n_levels"])]
for r in range(reps):
    out, slab = runs[r] if runs else run(int(os.environ.get("GPC", "32")))
    diff = np.nonzero((slab != ref).any(axis=1))[0]
    if diff.size == 0:
        print("run", r, "identical", flush=True)
        continue
    row = int(diff[0])
    L = max(i for i, f in enumerate(firsts) if f <= row)
    g, first, _ = c.level_plan(L, 0, 1)
    idx = row - first
    nd = int((slab[row] != ref[row]).sum())
    print("run", r, "DIFF rows", diff.size, "first row", row, "level", L, "width", len(g), "index", idx, "words differing", nd,
          "gate", g[idx] if idx < len(g) else None, "kat", out == v["golden"], flush=True)
    bad = np.nonzero(slab[row] != ref[row])[0]
    print("   differing word indices (first 12):", bad[:12].tolist(), "got", slab[row][bad[:6]].tolist(), "want", ref[row][bad[:6]].tolist(), flush=True)
