"""Small fixed invocation for ncu: STD128_OPT GINX, one full wave of 4-gate tiles (148 CTAs x 4 gates), NAND."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bfhe_loader

B = bfhe_loader.load_package()
ctx = B.Context(B.STD128_OPT, B.GINX, 0)
ctx.keygen(1)
ctx.btkeygen(2)
count = int(sys.argv[1]) if len(sys.argv) > 1 else 592
gpc = int(sys.argv[2]) if len(sys.argv) > 2 else 0
n_in = 2 * count
bits = np.random.default_rng(0).integers(0, 2, n_in)
slab = ctx.slab(n_in + count)
slab.upload(ctx.encrypt(bits, seed=1))
g = np.zeros(count, dtype=B.GATE_DTYPE)
g["op"] = B.NAND
g["in0"] = 2 * np.arange(count)
g["in1"] = 2 * np.arange(count) + 1
g["out"] = n_in + np.arange(count)
ctx.dbg_set_gates_per_cta(gpc)
for _ in range(3):
    ctx.eval_bingate_batch(slab, g)
ctx.sync()
dec = ctx.decrypt(slab.download(n_in, count))
assert np.array_equal(dec, 1 - (bits[g["in0"]] & bits[g["in1"]]))
print("ok", count)
