"""Measure the integer-pipe instruction rates the roofline is quoted against."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bfhe_loader

B = bfhe_loader.load_package()
ctx = B.Context(B.TOY, B.GINX, 0)
names = ["imad", "imad_hi", "imad_wide", "alu3", "shoup_mul_as_imad_class"]
out = {}
for w, nm in enumerate(names):
    vals = [ctx.microbench_int(w) for _ in range(5)]
    out[nm + "_ginstr_per_s"] = max(vals)
    out[nm + "_all"] = vals
print(json.dumps(out))
