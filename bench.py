#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 gate-bootstrapping engine.

Metric (BASELINE.json): bootstrapped gates/sec, STD128_OPT GINX, on the synthetic wavefront of independent
NAND / AND / XOR gates (config 3: gate i has type i mod 3; the reference's composite XOR = 3 bootstraps,
src/gate.cpp:198-202), plus AES-128 whole-circuit wall time (config 5) as an auxiliary figure.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--gates G] [--impl reference]

One process per GPU (torchrun for N > 1; torch.distributed is plumbing only: barrier + max-reduce of times).
A step = one pass of the hot path (bfhe_eval_bingate_batch: blind rotation + key switch) over one batch of G gates
per rank whose input ciphertexts are already resident in HBM.  `e2e` is the same step through
bfhe_eval_bingate_host with pinned HOST buffers (H2D of the inputs and D2H of the outputs inside the timed region).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import bfhe_loader  # noqa: E402

METRIC = "bootstrapped gates/sec (STD128_OPT GINX)"
UNIT = "bootstrapped gates/s"
# algorithmic work of one bootstrapped gate (STD128_OPT GINX): 45.25 M modular multiplications
# = 502 x [10 NTTs x 5120 butterflies + 36 N pointwise + 2 N scaling] + 3 NTT-equivalents; x3 IMAD-class instructions each
W_MODMUL = 45_250_000
W_IMAD = 3 * W_MODMUL
# algorithmic bytes per bootstrapped gate (DESIGN.md "roofline"): BK once per CTA tile of G gates + KSK row gather + I/O
BK_BYTES = 65_798_144
KS_ROW_BYTES = 1024 * 2 * 1024  # 2048 padded 1 KiB rows
IO_BYTES = 3 * 504 * 4


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--gates", type=int, default=0,
                    help="gates per rank per step (config 3 mix); 0 = 42 full waves of 4-gate CTAs (24 864 on a 148-SM B200)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-aux", action="store_true", help="skip the AES-128 / SHA-256 circuit wall-time figures")
    ap.add_argument("--cpu-sample", type=int, default=0, help="gates in the CPU baseline sample (0 = 16 per core; reference arm: 8 per core per step)")
    return ap.parse_args()


def make_workload(B, count, n_in):
    g = np.zeros(count, dtype=B.GATE_DTYPE)
    idx = np.arange(count)
    g["op"] = np.array([B.NAND, B.AND, B.XOR], dtype=np.uint32)[idx % 3]
    g["in0"] = (2 * idx) % n_in
    g["in1"] = (2 * idx + 1) % n_in
    g["out"] = n_in + idx
    boots = int(np.where(g["op"] == B.XOR, 3, 1).sum())
    return g, boots


def expected_bits(B, g, bits):
    a, b = bits[g["in0"]], bits[g["in1"]]
    return np.where(g["op"] == B.NAND, 1 - (a & b), np.where(g["op"] == B.AND, a & b, a ^ b))


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (profiling recipe's clocks line)."""

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) >= 9:
                for nm, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_baseline(sample_gates, seed_keys=(1, 2), threads=0):
    """The oracle (CPU restatement of OpenFHE binfhe 1.0.x, oracle/bfhe_oracle.c) driven like the reference drives
    OpenFHE: one OpenMP task per ready gate of a wavefront (src/circuit.cpp:698-710), all host cores."""
    O = bfhe_loader.load_oracle()
    O.build()
    cores = threads or os.cpu_count() or 1
    o = O.Oracle(O.STD128_OPT, O.GINX)
    o.keygen(seed_keys[0])
    n_in = 64
    rng = np.random.default_rng(42)
    bits = rng.integers(0, 2, n_in)
    slab = o.new_slab(n_in + sample_gates)
    slab[:n_in] = o.encrypt(bits, seed=42)
    g = np.zeros(sample_gates, dtype=O.GATE_DTYPE)
    idx = np.arange(sample_gates)
    g["op"] = np.array([O.NAND, O.AND, O.XOR], dtype=np.uint32)[idx % 3]
    g["in0"] = (2 * idx) % n_in
    g["in1"] = (2 * idx + 1) % n_in
    g["out"] = n_in + idx
    boots = int(np.where(g["op"] == O.XOR, 3, 1).sum())
    return o, g, slab, boots, cores, bits


def run_reference(args, rank):
    """--impl reference: the reference's CPU implementation of the path = the oracle port (OpenFHE itself cannot be
    built here: not vendored, no network), all host threads, same workload mix, bounded sample per step."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    sample = args.cpu_sample or 8 * cores  # ~0.7 s per step on 16 cores; XOR gates are 3 serial bootstraps inside one task
    o, g, slab, boots, cores, bits = cpu_baseline(sample)
    times = []
    for it in range(args.warmup + args.steps):
        t = time.perf_counter()
        o.eval_gates(g, slab, nthreads=cores)
        dt = time.perf_counter() - t
        if it >= args.warmup:
            times.append(dt)
    dec = o.decrypt(slab[64:])
    a, b = bits[g["in0"]], bits[g["in1"]]
    ok = bool(np.array_equal(dec, np.where(g["op"] == 0 + 3, 1 - (a & b), np.where(g["op"] == 1, a & b, a ^ b))))
    total = sum(times)
    val = boots * args.steps / total
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {"workload": "synthetic independent NAND/AND/XOR gates, STD128_OPT GINX (config 3 mix)",
                       "gates_per_step": sample, "bootstraps_per_step": boots, "xor": "composite, 3 bootstraps", "decrypt_ok": ok},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": "%d gates (%d bootstraps) per step, one OpenMP task per gate, oracle/bfhe_oracle.c" % (sample, boots)},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    B = bfhe_loader.load_package()
    ctx = B.Context(B.STD128_OPT, B.GINX, local)
    ctx.keygen(1)   # keys from seed 1 (config 3), replicated on every GPU
    ctx.btkeygen(2)
    stream = torch.cuda.Stream()  # every kernel of the engine is launched on this stream; the events below sit on it too
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)

    # default batch: a multiple of 3 (gate mix) x 4 (gates per CTA) x SM count, so every launch is whole waves
    sms = torch.cuda.get_device_properties(local).multi_processor_count
    count = args.gates or 3 * 4 * sms * 14
    n_in = 2 * count
    gates, boots = make_workload(B, count, n_in)
    bits = np.random.default_rng(42 + rank).integers(0, 2, n_in)
    st = ctx.stride
    host_in = torch.empty((n_in, st), dtype=torch.int32).pin_memory()
    host_in.numpy().view(np.uint32)[:] = ctx.encrypt(bits, seed=1000 + rank)
    host_out = torch.empty((count, st), dtype=torch.int32).pin_memory()
    slab = torch.zeros((n_in + count, st), dtype=torch.int32, device="cuda")
    slab[:n_in].copy_(host_in)
    torch.cuda.synchronize()
    slab_ptr = slab.data_ptr()

    # integer-pipe peak, measured live (the roofline denominator of this integer-bound path)
    imad_peak = max(ctx.microbench_int(0) for _ in range(3)) * 1e9

    # ---------------- device-resident throughput (value) ----------------
    for _ in range(args.warmup):
        ctx.eval_bingate_batch(slab_ptr, gates)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ctx.profile_enable(True)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record(stream)
    for _ in range(args.steps):
        ctx.eval_bingate_batch(slab_ptr, gates)
    ev1.record(stream)
    barrier()
    dev_ms = max_over_ranks(ev0.elapsed_time(ev1))
    br_ms, br_n = ctx.profile_read(0)
    ks_ms, ks_n = ctx.profile_read(1)
    nt_ms, nt_n = ctx.profile_read(2)
    ctx.profile_enable(False)
    clocks = sampler.stop() if rank == 0 else None
    value = world * boots * args.steps / (dev_ms * 1e-3)

    # correctness of what was just timed: every output decrypts to the truth table
    out = slab[n_in:].cpu().numpy().view(np.uint32)
    ok = bool(np.array_equal(ctx.decrypt(out), expected_bits(B, gates, bits)))

    # ---------------- end to end through the host-buffer API (e2e) ----------------
    hin, hout = host_in.numpy().view(np.uint32), host_out.numpy().view(np.uint32)
    ctx.eval_bingate_host(gates, hin, count, hout)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ctx.eval_bingate_host(gates, hin, count, hout)
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    barrier()
    e2e_val = world * boots * args.steps / e2e_s
    ok = ok and bool(np.array_equal(ctx.decrypt(hout), expected_bits(B, gates, bits)))

    # ---------------- roofline of the dominant kernel (blind rotation) ----------------
    br_s_per_launch = br_ms * 1e-3 / max(br_n, 1)
    boots_per_launch = boots * args.steps / max(br_n, 1)
    achieved_imad = boots_per_launch * W_IMAD / br_s_per_launch
    tile = 4
    hbm_bytes_per_boot = BK_BYTES / tile + KS_ROW_BYTES + IO_BYTES
    roofline = {"bound": "int", "kernel": "blind_rotate_kernel<10,4,7,4,GINX>", "achieved": achieved_imad / 1e12,
                "peak": imad_peak / 1e12, "unit": "TIMAD/s", "frac": achieved_imad / imad_peak,
                "peak_source": "measured live: register-only IMAD microbenchmark (bfhe_microbench_int), 148 SMs",
                "work_per_unit": "%d IMAD-class instr per bootstrapped gate (3 x 45.25M modular multiplications, SURVEY 8(d))" % W_IMAD,
                "kernel_ms_per_launch": 1e3 * br_s_per_launch, "kernel_share_of_step": br_ms / (br_ms + ks_ms + nt_ms),
                "step_ms_by_kernel": {"blind_rotate": br_ms / args.steps, "keyswitch": ks_ms / args.steps, "eval_not": nt_ms / args.steps},
                # dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture of this kernel on a 592-gate
                # launch (profiles/r1_blind_rotate_ncu_summary.md): the 65.8 MB key is read from HBM once, then served by L2
                "traffic": 73.1e6, "traffic_note": "bytes per 592-gate launch (ncu); algorithmic key bytes reach DRAM once per launch",
                "hbm": {"achieved": boots_per_launch * hbm_bytes_per_boot / br_s_per_launch / 1e9, "peak": peak_hbm(),
                        "unit": "GB/s", "bytes_per_unit": hbm_bytes_per_boot,
                        "note": "key streaming is L2/HBM traffic shared by the %d gates of a CTA tile; the path is integer-bound" % tile}}
    roofline["hbm"]["frac"] = roofline["hbm"]["achieved"] / roofline["hbm"]["peak"]

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u32", "data": "synthetic",
            "config": {"workload": "synthetic independent NAND/AND/XOR gates, STD128_OPT GINX (BASELINE config 3 mix), "
                                   "%d gates/rank/step" % count,
                       "gates_per_rank_per_step": count, "bootstraps_per_rank_per_step": boots, "xor": "composite, 3 bootstraps",
                       "gates_per_s": world * count * args.steps / (dev_ms * 1e-3),
                       "l2": "inputs (%d MB) + keys (330 MB) exceed the 126 MB L2" % (n_in * st * 4 // 2**20),
                       "parallelism": "replicated keys, gates sharded by rank, no data-path collective", "decrypt_ok": ok},
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": int(n_in * st * 4 + count * 32),
                    "d2h_bytes_per_step": int(count * st * 4)},
            "gpu_launches": int(br_n + ks_n + nt_n),
            "roofline": roofline, "clocks": clocks}
    if rank == 0 and world == 1:
        cores = os.cpu_count() or 1
        sample = args.cpu_sample or 16 * cores  # about 20-30 core-seconds of CPU work
        o, g, s2, b2, cores, _ = cpu_baseline(sample)
        t = time.perf_counter()
        o.eval_gates(g, s2, nthreads=cores)
        dt = time.perf_counter() - t
        line["cpu_baseline"] = {"value": b2 / dt, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": "%d gates (%d bootstraps) of the same mix, one OpenMP task per gate, %.1f s" % (sample, b2, dt)}
    if not args.no_aux:
        line["aux"] = aux_circuits(B, ctx, rank, world)
    if not ok:
        line["error"] = "decrypted outputs do not match the truth table"
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def peak_hbm():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6650.0  # fallback stated in B200_PROFILING.md


def aux_circuits(B, ctx, rank, world):
    """AES-128 and SHA-256 whole-circuit wall time (BASELINE config 5), waves sharded over the ranks with ncclAllGather.
    AES: second of two evaluations (the first builds the CUDA graph); SHA-256 (13 s of depth-bound work): one evaluation."""
    import torch
    import torch.distributed as dist
    vectors = json.load(open(os.path.join(ROOT, "tests", "golden", "vectors.json")))
    res = {}
    for name, key, reps in (("AES-non-expanded", "aes128", 2), ("sha256", "sha256", 1)):
        vec = vectors[name]["vectors"][1]
        c = B.Circuit(ctx)
        c.load_npz(os.path.join(ROOT, "tests", "golden", "circuits", name + ".npz"))
        if world > 1:
            uid = torch.from_numpy(B.nccl_unique_id() if rank == 0 else np.zeros(128, dtype=np.uint8)).cuda()
            dist.broadcast(uid, 0)
            c.set_sharding(rank, world, uid.cpu().numpy())
        for rep in range(reps):
            c.Reset()
            c.setEncrypted(True)
            c.SetInput(vec["inputs"], seed=7)
            if world > 1:
                dist.barrier()
            t = time.perf_counter()
            out = c.Clock()[0]
            wall = time.perf_counter() - t
            res.update({key + "_wall_ms": 1e3 * wall, key + "_device_ms": c.stats()["device_ms"], key + "_kat_ok": out == vec["golden"],
                        key + "_bootstraps": c.info()["bootstraps"], key + "_levels": c.info()["levels"],
                        key + "_waves": c.plan_misc()["n_levels"] - 1})
        c.close()
    return res


if __name__ == "__main__":
    main()
