"""Row e on hardware: a circuit sharded over 2 GPUs (one process per GPU, NCCL for set-up) gives, on every rank, the same wire
ciphertexts as the unsharded evaluation -- with the exchange fused into the key switch (stores into the peers' slabs) and with the
ncclAllGather fallback.  Needs two CUDA devices: skipped on the one-GPU test box (bench.py's aux reports the same check,
`sharded_slab_equal`, at every N > 1)."""
import hashlib
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _digest(c, slab):
    """the bootstrap-output rows of every level, in level order: independent of how the rows are padded for sharding"""
    h = hashlib.sha256()
    w = c.ctx.p.ct_words
    for L in range(c.plan_misc()["n_levels"]):
        first, n = None, 0
        for r in range(max(1, c.world_size)):
            g, f, rpr = c.level_plan(L, r, c.world_size)
            first = f
            n += len(g) if rpr or c.world_size == 1 else (len(g) if r == 0 else 0)
        h.update(np.ascontiguousarray(slab[first:first + n, :w]).tobytes())
    return h.hexdigest()


def _worker(rank, world, port, mode, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import bfhe_loader
    from helpers import VECTORS, load_circuit
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    if mode == "nccl":
        os.environ["BFHE_EXCHANGE"] = "nccl"
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world)
    B = bfhe_loader.load_package()
    ctx = B.Context(B.STD128_OPT, B.GINX, rank)
    if rank == 0:  # keys are generated once and distributed as the BFHEKEY1 blob (include/bfhe.h, randomness contract)
        ctx.keygen(31)
        ctx.btkeygen(32)
        blob = torch.from_numpy(ctx.export_keys()).cuda()
        size = torch.tensor([blob.numel()], device="cuda")
    else:
        size = torch.tensor([0], device="cuda")
    dist.broadcast(size, 0)
    if rank != 0:
        blob = torch.empty(int(size), dtype=torch.uint8, device="cuda")
    dist.broadcast(blob, 0)
    if rank != 0:
        ctx.import_keys(blob.cpu().numpy())
    ok, modes = True, []
    for name, thr in (("mult_32x32", 0), ("adder_32bit", 0), ("mult_32x32", -1)):  # thr 0 = shard every level; -1 = cost model
        v = VECTORS[name]["vectors"][0]
        c = load_circuit(B, ctx, name)
        uid = torch.from_numpy(B.nccl_unique_id() if rank == 0 else np.zeros(128, dtype=np.uint8)).cuda()
        dist.broadcast(uid, 0)
        c.set_sharding(rank, world, uid.cpu().numpy())
        c.set_shard_threshold(thr)
        for rep in range(2):  # the second Clock re-launches the captured graph: epochs, end-of-Clock signals
            c.Reset()
            c.setEncrypted(True)
            c.SetInput(v["inputs"], seed=9)
            out = c.Clock()[0]
            ok = ok and out == v["golden"]
        modes.append(c.exchange_mode())
        cap = c.schedule()["wave_cap"]
        mine = _digest(c, c.download_slab())
        c1 = load_circuit(B, ctx, name)  # the same schedule, unsharded, on this GPU
        c1.set_wave_capacity(cap)
        c1.Reset()
        c1.setEncrypted(True)
        c1.SetInput(v["inputs"], seed=9)
        ok = ok and c1.Clock()[0] == v["golden"] and _digest(c1, c1.download_slab()) == mine
        c1.close()
        c.close()
    q.put((rank, ok, modes))
    dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["peer", "nccl"])
def test_sharded_circuit_two_gpus(mode):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29620 + (0 if mode == "peer" else 1)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, mode, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in res), res
    want = 2 if mode == "peer" else 1
    assert all(m == want for _, _, ms in res for m in ms), res
