"""N > 1 host logic on CPU: world_size-2 gloo.  Each rank evaluates its block of every level of the product's plan
(oracle as the gate executor -- there is no GPU here) and the level's output rows are exchanged with all_gather,
exactly the pattern the CUDA path runs with ncclAllGather (host/circuit.cpp enqueue_levels)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, name, nvec, q, cap=0, thr=-1):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import bfhe_loader
    from helpers import VECTORS, load_circuit, oracle_run_plan
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    B, O = bfhe_loader.load_package(), bfhe_loader.load_oracle()
    # keys: rank 0 generates with the product's host keygen, everyone imports the same blob
    ctx = B.Context(B.TOY, B.GINX, device=-1)
    if rank == 0:
        ctx.keygen(21)
        ctx.btkeygen(22)
        blob = torch.from_numpy(ctx.export_keys())
        size = torch.tensor([blob.numel()])
    else:
        size = torch.tensor([0])
    dist.broadcast(size, 0)
    if rank != 0:
        blob = torch.empty(int(size), dtype=torch.uint8)
    dist.broadcast(blob, 0)
    o = O.Oracle(O.TOY, O.GINX)
    o.import_keys(blob.numpy())
    circ = load_circuit(B, ctx, name)
    circ.set_sharding(rank, world)
    circ.set_wave_capacity(cap)  # 0 = the reference's ASAP waves; > 0 = packed waves (must be identical on every rank)
    if thr >= 0:  # levels narrower than thr are evaluated redundantly by every rank and skip the exchange
        circ.set_shard_threshold(thr)
        widths = [len(circ.level_plan(L, rank, world)[0]) for L in range(circ.plan_misc()["n_levels"])]
        unsharded = [L for L in range(len(widths)) if circ.level_plan(L, rank, world)[2] == 0 and widths[L]]
        assert unsharded and len(unsharded) < len(widths), "the threshold must split the levels into both kinds"

    def gather(blk, r, rpr):
        mine = torch.from_numpy(np.ascontiguousarray(blk[r * rpr:(r + 1) * rpr]).astype(np.int32))
        outs = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(outs, mine)
        for k in range(world):
            blk[k * rpr:(k + 1) * rpr] = outs[k].numpy().astype(np.uint32)

    ok = True
    slab_sum = 0
    for v in VECTORS[name]["vectors"][:nvec]:
        out, slab = oracle_run_plan(circ, o, v["inputs"], seed=6, rank=rank, world=world, gather=gather)
        ok = ok and out == v["golden"]
        slab_sum += int(slab.astype(np.uint64).sum())
    # replicas must agree on every wire ciphertext after the exchange
    t = torch.tensor([slab_sum], dtype=torch.int64)
    lst = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(lst, t)
    ok = ok and all(int(x) == int(lst[0]) for x in lst)
    q.put((rank, ok))
    dist.destroy_process_group()


@pytest.mark.parametrize("name,nvec,cap,thr", [("adder_2bit", 3, 0, -1), ("parity", 2, 0, -1), ("parity", 2, 4, -1), ("adder_2bit", 2, 0, 4)])
def test_world2_gloo_sharded_levels(name, nvec, cap, thr):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (os.getpid() % 200) + (0 if name == "adder_2bit" else 1) + 2 * (cap > 0) + 4 * (thr >= 0)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, name, nvec, q, cap, thr)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


def test_sharded_plan_equals_unsharded_results(bfhe, orc):
    """Single process emulation of 2 ranks: the union of both ranks' rows equals the world-1 evaluation bit for bit."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import VECTORS, load_circuit, oracle_run_plan
    ctx = bfhe.Context(bfhe.TOY, bfhe.GINX, device=-1)
    ctx.keygen(1)
    ctx.btkeygen(2)
    o = orc.Oracle(orc.TOY, orc.GINX)
    o.import_keys(ctx.export_keys())
    v = VECTORS["adder_2bit"]["vectors"][1]
    c1 = load_circuit(bfhe, ctx, "adder_2bit")
    out1, slab1 = oracle_run_plan(c1, o, v["inputs"], seed=3)
    c2 = load_circuit(bfhe, ctx, "adder_2bit")
    c2.set_sharding(0, 2)
    misc = c2.plan_misc()
    slab = o.new_slab(misc["total_rows"])
    flat = np.concatenate(v["inputs"])
    slab[misc["fresh_base"]:misc["fresh_base"] + flat.size] = o.encrypt(flat, seed=3)
    for L in range(misc["n_levels"]):
        for r in range(2):  # both "ranks" in turn on the shared slab == all_gather
            g, _, _ = c2.level_plan(L, r, 2)
            if len(g):
                o.eval_gates(g, slab)
        for a, b in c2.plan_misc(L)["nots"]:
            slab[b] = o.eval_not(slab[a])
    outs = [int(o.decrypt(slab[int(r) & 0x7fffffff])) ^ (int(r) >> 31) for r in misc["out_rows"]]
    assert outs == out1 == v["golden"]
    # same ciphertexts on the output wires, whatever the row numbering
    r1 = c1.plan_misc()["out_rows"]
    for a, b in zip(r1, misc["out_rows"]):
        assert np.array_equal(slab1[int(a) & 0x7fffffff], slab[int(b) & 0x7fffffff])
