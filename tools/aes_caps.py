"""AES-128 wall time under forced wave capacities (single GPU or sharded under torchrun): which schedule is really fastest?"""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bfhe_loader
B = bfhe_loader.load_package()
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
V = json.load(open(os.path.join(ROOT, "tests", "golden", "vectors.json")))
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
if world > 1:
    import torch, torch.distributed as dist
    torch.cuda.set_device(local)
    dist.init_process_group("nccl")


def fresh_uid():  # one NCCL communicator per circuit
    t = torch.from_numpy(B.nccl_unique_id() if rank == 0 else np.zeros(128, dtype=np.uint8)).cuda()
    dist.broadcast(t, 0)
    return t.cpu().numpy()
ctx = B.Context(B.STD128_OPT, B.GINX, local)
ctx.keygen(1); ctx.btkeygen(2)
name = os.environ.get("CIRCUIT", "AES-non-expanded")
v = V[name]["vectors"][0]
for cap in [int(a) for a in sys.argv[1:]] or [-1]:
    c = B.Circuit(ctx)
    c.load_npz(os.path.join(ROOT, "tests", "golden", "circuits", name + ".npz"))
    if world > 1:
        c.set_sharding(rank, world, fresh_uid())
    c.set_wave_capacity(cap)
    best = 1e9
    for rep in range(2):
        c.Reset(); c.setEncrypted(True)
        c.SetInput(v["inputs"], seed=rep)
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        out = c.Clock()[0]
        best = min(best, time.perf_counter() - t0)
    if rank == 0:
        print(json.dumps(dict(circuit=name, world=world, cap=cap, waves=c.plan_misc()["n_levels"] - 1, wall_ms=1e3 * best, kat_ok=out == v["golden"])), flush=True)
    c.close()
