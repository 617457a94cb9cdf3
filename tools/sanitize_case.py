"""Tiny TOY-parameter invocation of every hot-path kernel, for compute-sanitizer (memcheck / racecheck)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bfhe_loader

B = bfhe_loader.load_package()
for method in (B.GINX, B.AP):
    ctx = B.Context(B.TOY, method, 0)
    ctx.keygen(1)
    ctx.btkeygen(2)
    bits = np.array([1, 0, 1, 1, 0, 1], dtype=np.uint8)
    cts = ctx.encrypt(bits, seed=3)
    g = np.array([(B.NAND, 0, 1, 6), (B.XOR, 2, 3, 7), (B.AND | B.NEG1, 4, 5, 8), (B.BOOTSTRAP, 1, 1, 9), (B.OR, 0, 4, 10)],
                 dtype=B.GATE_DTYPE)
    for gpc in (8, 4, 2, 1):
        ctx.dbg_set_gates_per_cta(gpc)
        out = ctx.eval_bingate_host(g, cts, 5)
        assert ctx.decrypt(out).tolist() == [1, 0, 0, 0, 1], (method, gpc, ctx.decrypt(out))
    s = ctx.slab(8)
    s.upload(cts)
    ctx.eval_not_batch(s, [0, 1], [6, 7])
    assert ctx.decrypt(s.download(6, 2)).tolist() == [0, 1]
    ctx.close()
print("sanitize case ok")
