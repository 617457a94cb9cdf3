"""AP-method throughput (BASELINE 'next' row f-1): independent NAND gates, STD128_OPT AP, device-resident."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bfhe_loader
B = bfhe_loader.load_package()
ctx = B.Context(B.STD128_OPT, B.AP, 0)
t = time.time(); ctx.keygen(1); ctx.btkeygen(2); print("AP keygen+upload s", round(time.time() - t, 1), flush=True)
n_in = 4096
bits = np.random.default_rng(0).integers(0, 2, n_in)
cts = ctx.encrypt(bits, seed=1)
for gpc, count in ((128, 8), (128, 33), (8, 33), (8, 148), (4, 592), (4, 4736)):
    slab = ctx.slab(n_in + count); slab.upload(cts)
    g = np.zeros(count, dtype=B.GATE_DTYPE)
    g["op"] = B.NAND; g["in0"] = np.arange(count) % n_in; g["in1"] = (np.arange(count) * 7 + 1) % n_in; g["out"] = n_in + np.arange(count)
    ctx.dbg_set_gates_per_cta(gpc)
    ctx.eval_bingate_batch(slab, g[:min(count, 148)]); ctx.sync()
    ctx.profile_enable(True); ctx.eval_bingate_batch(slab, g); ctx.sync()
    br, _ = ctx.profile_read(0); ks, _ = ctx.profile_read(1); ctx.profile_enable(False)
    dec = ctx.decrypt(slab.download(n_in, count))
    ok = bool(np.array_equal(dec, 1 - (bits[g["in0"]] & bits[g["in1"]])))
    print(json.dumps(dict(method="AP", gpc=gpc, gates=count, blind_rotate_ms=br, keyswitch_ms=ks, gates_per_s=count / ((br + ks) * 1e-3), ok=ok)), flush=True)
    slab.free()
