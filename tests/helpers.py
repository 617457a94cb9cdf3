"""Shared test helpers.  The oracle (oracle/) is used here ONLY as the checker / stand-in executor of tests."""
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")
VECTORS = json.load(open(os.path.join(GOLDEN, "vectors.json")))


def circuit_path(name):
    return os.path.join(GOLDEN, "circuits", name + ".npz")


def load_circuit(bfhe, ctx, name):
    c = bfhe.Circuit(ctx)
    c.load_npz(circuit_path(name))
    return c


def oracle_run_plan(circ, o, inputs, seed=0, rank=0, world=1, gather=None, fresh=None, state=None):
    """Evaluate the product's level plan with the oracle as the gate executor (CPU).  Returns (out_bits, slab).
    gather(slab_block) -> exchanges a level's output rows between ranks (None for world == 1).
    state: the slab returned by the previous clock of a circuit with flip-flops (None = first clock after Reset)."""
    misc = circ.plan_misc()
    slab = o.new_slab(misc["total_rows"]) if state is None else state.copy()
    flat = np.concatenate([np.asarray(i, dtype=np.uint8) for i in inputs])
    fb = misc["fresh_base"]
    slab[fb:fb + flat.size] = fresh if fresh is not None else o.encrypt(flat, seed=seed)
    dff = circ.dff_plan()  # per flip-flop: D row | neg << 31, state row, latch row, fresh row
    if len(dff) and state is None:  # power-up: Bootstrap(Encrypt(0)) into the latch rows
        zeros = o.encrypt(np.zeros(len(dff), dtype=np.uint8), seed=seed + 77)
        for i, (_, _, latch, fr) in enumerate(dff):
            slab[fr] = zeros[i]
            slab[latch] = o.bootstrap(slab[fr])
    for _, st_row, latch, _ in dff:  # the clock starts by moving the latched values into the state rows
        slab[st_row] = slab[latch]
    for L in range(misc["n_levels"]):
        gates, first, rpr = circ.level_plan(L, rank, world)
        if len(gates):
            o.eval_gates(gates, slab)
        if world > 1 and rpr:
            blk = slab[first:first + rpr * world]
            gather(blk, rank, rpr)
        for a, b in circ.plan_misc(L)["nots"]:
            slab[b] = o.eval_not(slab[a])
    for d, _, latch, _ in dff:  # latch D at the end of the clock
        src = slab[int(d) & 0x7fffffff]
        slab[latch] = o.eval_not(src) if int(d) >> 31 else src
    outs = []
    for r in misc["out_rows"]:
        bit = o.decrypt(slab[int(r) & 0x7fffffff])
        outs.append(int(bit) ^ (int(r) >> 31))
    return outs, slab
