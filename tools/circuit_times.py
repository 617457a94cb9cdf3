"""Whole-circuit wall time (BASELINE configs 4 and 5), encrypted, KAT-checked; one JSON line per circuit."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bfhe_loader

B = bfhe_loader.load_package()
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
V = json.load(open(os.path.join(ROOT, "tests", "golden", "vectors.json")))
names = sys.argv[1:] or ["comparator_32bit_signed_lt", "adder_32bit", "mult_32x32", "AES-expanded", "AES-non-expanded", "md5", "sha256"]
ctx = B.Context(B.STD128_OPT, B.GINX, 0)
ctx.keygen(1)
ctx.btkeygen(2)
if os.environ.get("GPC"):
    ctx.dbg_set_gates_per_cta(int(os.environ["GPC"]))
for name in names:
    c = B.Circuit(ctx)
    c.load_npz(os.path.join(ROOT, "tests", "golden", "circuits", name + ".npz"))
    if os.environ.get("WAVE_CAP") is not None:
        c.set_wave_capacity(int(os.environ["WAVE_CAP"]))
    info = c.info()
    waves = c.plan_misc()["n_levels"] - 1
    best = None
    for rep, v in enumerate(V[name]["vectors"][:2]):
        c.Reset()
        c.setEncrypted(True)
        t0 = time.perf_counter()
        c.SetInput(v["inputs"], seed=rep)
        t1 = time.perf_counter()
        out = c.Clock()[0]
        t2 = time.perf_counter()
        ok = out == v["golden"]
        if not ok:
            print("MISMATCH", name, rep, sum(int(a != b) for a, b in zip(out, v["golden"])), "of", len(out), file=sys.stderr, flush=True)
        r = dict(circuit=name, bootstraps=info["bootstraps"], levels=info["levels"], waves=waves, max_width=info["max_width"],
                 set_input_ms=1e3 * (t1 - t0), clock_wall_ms=1e3 * (t2 - t1), device_ms=c.stats()["device_ms"], kat_ok=ok,
                 bootstraps_per_s=info["bootstraps"] / (t2 - t1))
        if best is None or r["clock_wall_ms"] < best["clock_wall_ms"] or not ok:
            best = r
    print(json.dumps(best), flush=True)
    c.close()
