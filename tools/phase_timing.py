"""Per-phase cycle breakdown of the latency kernel (debug build with -DBFHE_PHASE_TIMING)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bfhe_loader
B = bfhe_loader.load_package()
if os.environ.get('BFHE_LIB'):
    B.LIB_PATH = os.path.join(os.path.dirname(B.LIB_PATH), os.environ['BFHE_LIB'])
ctx = B.Context(B.STD128_OPT, B.GINX, 0)
ctx.keygen(1); ctx.btkeygen(2)
for count in [int(a) for a in os.environ.get('COUNTS', '1,148').split(',')]:
    bits = np.random.default_rng(0).integers(0, 2, 2 * count)
    slab = ctx.slab(3 * count); slab.upload(ctx.encrypt(bits, seed=1))
    g = np.zeros(count, dtype=B.GATE_DTYPE)
    g["op"] = B.NAND; g["in0"] = 2 * np.arange(count); g["in1"] = 2 * np.arange(count) + 1; g["out"] = 2 * count + np.arange(count)
    ctx.dbg_set_gates_per_cta(int(os.environ.get('GPC', '8')))
    acc = ctx.dbg_blind_rotate(slab, g)
    if os.environ.get('GPC') in ('1', '2', '4'):  # first-generation throughput kernel: [gate][component][32 + phase]
        t = acc[:, :, 32:37].astype(np.float64)
        names = ["intt+decompose+4 ntt", "keyload+barrier1", "mac", "barrier2", "-"]
        for w in (0, 1):
            print(count, "gates, comp", w, {n: round(float(v), 1) for n, v in zip(names, t[:, w].mean(0))}, "total kcyc", round(float(t[:, w].sum(1).mean()), 1))
        continue
    if os.environ.get('GPC') == '128':  # slot-sliced cluster kernel: [gate][component 1][256 rank + 32 + 16 * who + phase], who = main warp 0 / 1, push warp 0 / 1
        names = ["group barrier", "A lut", "B sub-ntt", "dct barrier", "key wait", "C mac", "C inverse stages+push", "recv wait", "E cross+acc", "E sub-intt+barrier"]
        pnames = ["wait", "push", "key"] + ["-"] * 7
        for who, label in enumerate(["main warp 0 (group 0)", "main warp 4 (group 1)"]):
            t = np.stack([acc[:, 1, 256 * k + 32 + 16 * who:256 * k + 42 + 16 * who] for k in range(4)], 1).astype(np.float64)
            nm = names if who < 2 else pnames
            print(count, "gates,", label, {n: round(float(v), 1) for n, v in zip(nm, t.mean((0, 1))) if n != "-"}, "total kcyc", round(float(t.sum(2).mean()), 1))
        # one step's timeline (step 200, gate 0, CTA 0): clock of every mark relative to main warp 0 leaving the group barrier
        st = [acc[0, 0, 32 + 16 * who:42 + 16 * who].astype(np.int64) for who in range(2)]
        t0 = int(st[0][0])
        for who, label in enumerate(["main 0", "main 4"]):
            nm = names if who < 2 else pnames
            ev = sorted((int((v - t0) & 0xffffffff) if ((v - t0) & 0xffffffff) < 2**31 else int((v - t0) & 0xffffffff) - 2**32, n) for n, v in zip(nm, st[who]) if n != "-" and v)
            print("  timeline", label, " ".join("%s@%d" % (n.replace(" ", "_"), t) for t, n in ev))
        continue
    if os.environ.get('GPC') == '32':  # cluster kernel: [gate][rank][32 + 8*(warp==7) + phase]
        for h in (0, 1):
            t = acc[:, :, 32 + 8 * h:37 + 8 * h].astype(np.float64)
            names = ["intt(w0-1)+sync", "fwd split", "keywait+cluster sync", "mac", "cluster sync"]
            print(count, "gates, warp", 7 * h, {n: round(float(v), 1) for n, v in zip(names, t.mean((0, 1)))}, "total kcyc", round(float(t.sum(2).mean()), 1))
        continue
    if os.environ.get('GPC') == '16':  # kernels_v2.cu: [gate][component][32 + 8*h + phase]
        for h in (0, 1):
            t = acc[:, :, 32 + 8 * h:37 + 8 * h].astype(np.float64)
            names = ["intt+publish+polybar", "fwd ntt x2", "keyload+pairbar1", "mac", "pairbar2"]
            print(count, "gates, h", h, {n: round(float(v), 1) for n, v in zip(names, t.mean((0, 1)))}, "total kcyc", round(float(t.sum(2).mean()), 1))
        continue
    t = acc[:, :, 32:37].astype(np.float64)  # kilo-cycles per phase, warps 0 and 1
    names = ["intt+decompose", "barrier1", "ntt", "barrier2+keywait", "mac"]
    for w in (0, 1):
        print(count, "gates, warp", w, {n: round(float(v), 1) for n, v in zip(names, t[:, w].mean(0))}, "total kcyc", round(float(t[:, w].sum(1).mean()), 1))
