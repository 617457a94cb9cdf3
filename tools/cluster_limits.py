import sys, os
sys.path.insert(0, os.getcwd())
import bfhe_loader
B = bfhe_loader.load_package()
ctx = B.Context(B.STD128_OPT, B.GINX, 0)
print(ctx.dbg_cluster_limits())
