"""Latency of one wave of bootstraps per kernel form: `python tools/perf_lat.py 8:148 32:74 32:37 16:296` (form:gates)."""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bfhe_loader
B = bfhe_loader.load_package()
if os.environ.get('BFHE_LIB'):
    B.LIB_PATH = os.path.join(os.path.dirname(B.LIB_PATH), os.environ['BFHE_LIB'])
ctx = B.Context(B.STD128_OPT, B.GINX, 0)
ctx.keygen(1); ctx.btkeygen(2)
n_in = 1024
bits = np.random.default_rng(0).integers(0, 2, n_in)
cts = ctx.encrypt(bits, seed=1)
for spec in sys.argv[1:] or ["8:148", "32:74"]:
    gpc, count = (int(x) for x in spec.split(":"))
    slab = ctx.slab(n_in + count); slab.upload(cts)
    g = np.zeros(count, dtype=B.GATE_DTYPE)
    g["op"] = B.NAND; g["in0"] = np.arange(count) % n_in; g["in1"] = (np.arange(count) * 7 + 1) % n_in; g["out"] = n_in + np.arange(count)
    ctx.dbg_set_gates_per_cta(gpc)
    ctx.eval_bingate_batch(slab, g); ctx.sync()
    best = 1e9
    for _ in range(3):
        ctx.profile_enable(True)
        ctx.eval_bingate_batch(slab, g); ctx.sync()
        br, _ = ctx.profile_read(0); ks, _ = ctx.profile_read(1); ctx.profile_enable(False)
        best = min(best, br)
    dec = ctx.decrypt(slab.download(n_in, count))
    ok = bool(np.array_equal(dec, 1 - (bits[g["in0"]] & bits[g["in1"]])))
    print(json.dumps(dict(form=gpc, gates=count, br_ms=best, ks_ms=ks, ok=ok)), flush=True)
    slab.free()
