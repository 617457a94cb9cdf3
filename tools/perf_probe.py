"""Quick throughput probe: independent NAND gates, STD128_OPT GINX, device-resident inputs."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bfhe_loader

B = bfhe_loader.load_package()
ps = B.STD128_OPT if "--toy" not in sys.argv else B.TOY
ctx = B.Context(ps, B.GINX, 0)
t = time.time()
ctx.keygen(1)
ctx.btkeygen(2)
print("keygen+upload s", time.time() - t, flush=True)
n_in = 4096
bits = np.random.default_rng(0).integers(0, 2, n_in)
cts = ctx.encrypt(bits, seed=1)
res = []
for gpc, count in ((8, 1), (8, 100), (8, 148), (8, 200), (8, 296), (8, 592), (1, 148), (2, 296), (4, 592), (4, 2368), (4, 9472)):
    slab = ctx.slab(n_in + count)
    slab.upload(cts)
    g = np.zeros(count, dtype=B.GATE_DTYPE)
    g["op"] = B.NAND
    g["in0"] = np.arange(count) % n_in
    g["in1"] = (np.arange(count) * 7 + 1) % n_in
    g["out"] = n_in + np.arange(count)
    ctx.dbg_set_gates_per_cta(gpc)
    ctx.eval_bingate_batch(slab, g[: min(count, 148 * gpc)])
    ctx.sync()
    ctx.profile_enable(True)
    t = time.time()
    ctx.eval_bingate_batch(slab, g)
    ctx.sync()
    dt = time.time() - t
    br, nbr = ctx.profile_read(0)
    ks, nks = ctx.profile_read(1)
    ctx.profile_enable(False)
    out = slab.download(n_in, count)
    dec = ctx.decrypt(out)
    ok = bool(np.array_equal(dec, 1 - (bits[g["in0"]] & bits[g["in1"]])))
    r = dict(gpc=gpc, gates=count, wall_s=dt, gates_per_s=count / dt, blind_rotate_ms=br, keyswitch_ms=ks, ok=ok)
    print(json.dumps(r), flush=True)
    res.append(r)
    slab.free()
json.dump(res, open("gpurun_out/perf_probe.json", "w"))
