"""GPU: the reference's circuit-level configurations through the Circuit mirror (C ABI underneath).
Decrypted outputs must equal the harness goldens; every wire ciphertext must equal the oracle's on small circuits."""
import numpy as np
import pytest

from conftest import shared_keys
from helpers import VECTORS, load_circuit, oracle_run_plan

pytestmark = pytest.mark.gpu


def _run_encrypted(c, v, seed=0, verify=True):
    c.Reset()
    c.setEncrypted(True)
    c.setVerify(verify)
    c.SetInput(v["inputs"], seed=seed)
    return c.Clock()[0]


@pytest.mark.parametrize("ps,m", [("STD128_OPT", "GINX"), ("TOY", "GINX"), ("TOY", "AP")])
def test_config1_adder_2bit(bfhe, orc, ps, m):
    """TB_adder_2bit (BASELINE config 1): 10 seeded vectors, encrypted + gate-by-gate verify (src/test_adder.cpp:265-295)."""
    ctx = shared_keys(bfhe, getattr(bfhe, ps), getattr(bfhe, m), 0)
    c = load_circuit(bfhe, ctx, "adder_2bit")
    for t, v in enumerate(VECTORS["adder_2bit"]["vectors"]):
        assert _run_encrypted(c, v, seed=t) == v["golden"], v["src"]
        assert c.plain_out[0] == v["golden"]
        assert c.stats()["verify_mismatches"] == 0
    # wire-level: whole slab bit-identical to the oracle executing the same plan on the same fresh encryptions
    o = orc.Oracle(getattr(orc, ps), getattr(orc, m))
    o.import_keys(ctx.export_keys())
    v = VECTORS["adder_2bit"]["vectors"][3]
    _run_encrypted(c, v, seed=42)
    slab = c.download_slab()
    fresh = ctx.encrypt(np.concatenate(v["inputs"]), seed=42)
    out, ref = oracle_run_plan(c, o, v["inputs"], fresh=fresh)
    assert out == v["golden"]
    w = ctx.p.ct_words
    assert np.array_equal(slab[:, :w], ref[:, :w])


@pytest.mark.parametrize("ps,m", [("STD128_OPT", "GINX"), ("STD128_OPT", "AP"), ("TOY", "AP")])
def test_config2_parity_chained(bfhe, orc, ps, m):
    """TB_parity (BASELINE config 2): generate, feed 'even' back into bit 8, check -> (0,1); GINX and AP."""
    ctx = shared_keys(bfhe, getattr(bfhe, ps), getattr(bfhe, m), 0)
    c = load_circuit(bfhe, ctx, "parity")
    nvec = 20 if m == "GINX" else 8
    for t, v in enumerate(VECTORS["parity"]["vectors"][:nvec]):
        assert _run_encrypted(c, v, seed=100 + t) == v["golden"], v["src"]
        assert c.stats()["verify_mismatches"] == 0
    if ps == "STD128_OPT" and m == "AP":  # one AP gate, bit-exact against the oracle (config 2 names AP explicitly)
        o = orc.Oracle(orc.STD128_OPT, orc.AP)
        o.import_keys(ctx.export_keys())
        cts = ctx.encrypt([1, 1], seed=8)
        out = ctx.EvalBinGate(bfhe.AND, cts[0], cts[1])
        w = ctx.p.ct_words
        assert np.array_equal(out[:w], o.eval_bingate(orc.AND, cts[0], cts[1])[:w])
        assert ctx.decrypt(out)[0] == 1


@pytest.mark.parametrize("name", ["comparator_32bit_signed_lt", "comparator_32bit_signed_lteq",
                                  "comparator_32bit_unsigned_lt", "comparator_32bit_unsigned_lteq", "adder_32bit"])
def test_config4_comparators_and_adder(bfhe, name):
    ctx = shared_keys(bfhe, bfhe.STD128_OPT, bfhe.GINX, 0)
    c = load_circuit(bfhe, ctx, name)
    for t, v in enumerate(VECTORS[name]["vectors"][:4]):  # vector 0 is the forced-equal case
        assert _run_encrypted(c, v, seed=t) == v["golden"], v["src"]
        assert c.stats()["verify_mismatches"] == 0


def test_config4_multiplier(bfhe):
    ctx = shared_keys(bfhe, bfhe.STD128_OPT, bfhe.GINX, 0)
    c = load_circuit(bfhe, ctx, "mult_32x32")
    for t in (0, 1):
        v = VECTORS["mult_32x32"]["vectors"][t]
        assert _run_encrypted(c, v, seed=t, verify=(t == 0)) == v["golden"], v["src"]
        assert c.stats()["verify_mismatches"] == 0


def test_config5_aes128(bfhe):
    """old-bristol AES-128 (non-expanded), both reference KATs, graph and per-level launch paths agree bit for bit."""
    ctx = shared_keys(bfhe, bfhe.STD128_OPT, bfhe.GINX, 0)
    c = load_circuit(bfhe, ctx, "AES-non-expanded")
    v0, v1 = VECTORS["AES-non-expanded"]["vectors"]
    assert _run_encrypted(c, v0, seed=1, verify=True) == v0["golden"]
    assert c.stats()["verify_mismatches"] == 0
    assert _run_encrypted(c, v1, seed=2, verify=False) == v1["golden"]
    slab_graph = c.download_slab()
    c.use_graph(False)
    assert _run_encrypted(c, v1, seed=2, verify=False) == v1["golden"]
    assert np.array_equal(slab_graph, c.download_slab())


def test_out_file_path_on_gpu(bfhe, tmp_path):
    """ReadFile('.out') -> encrypted Clock, the reference's own entry path (src/test_adder.cpp:155-156)."""
    ctx = shared_keys(bfhe, bfhe.TOY, bfhe.GINX, 0)
    c0 = load_circuit(bfhe, ctx, "adder_2bit")
    p = tmp_path / "adder_2bit.out"
    c0.write_out(p)
    c = bfhe.Circuit(ctx)
    c.ReadFile(p)
    for t, v in enumerate(VECTORS["adder_2bit"]["vectors"][:5]):
        assert _run_encrypted(c, v, seed=t) == v["golden"]


def test_cpp_host_mirror_tb_adder(bfhe, tmp_path):
    """The C++ host layer (host/circuit.h + host/binfhecontext.h, the reference's class / method names) drives the same
    engine: examples/tb_circuit.cpp is TB_adder_2bit rewritten against it; exit code 0 = plaintext, encrypted+verify and
    the single-gate EvalBinGate / EvalNOT / throw-on-alias behaviour all pass."""
    import os
    import subprocess
    ctx = shared_keys(bfhe, bfhe.TOY, bfhe.GINX, 0)
    c0 = load_circuit(bfhe, ctx, "adder_2bit")
    p = tmp_path / "adder_2bit.out"
    c0.write_out(p)
    exe = os.path.join(os.path.dirname(bfhe.LIB_PATH), "examples", "tb_circuit")
    assert os.path.exists(exe), "run build() first"
    r = subprocess.run([exe, str(p), "-s", "STD128_OPT", "-m", "GINX", "-n", "4"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "ALL PASSED" in r.stdout


def test_kernel_forms_agree_on_every_wire(bfhe):
    """Race / determinism stress: the 32x32 multiplier (9 133 bootstraps, waves of 1..148 gates) evaluated with each
    blind-rotation form forced -- one gate per CTA (8), four per CTA (4), second-generation (16), one gate on a 2-CTA
    cluster (32), on a 4-CTA cluster (64) -- and twice with the cost model's own choice: every wire ciphertext must be bit-identical across all
    runs (a stale shared-memory read in the cluster form once produced valid-but-different ciphertexts in 1 of ~400k
    bootstraps; tools/race_hunt.py is the long-running version of this test)."""
    ctx = shared_keys(bfhe, bfhe.STD128_OPT, bfhe.GINX, 0)
    c = load_circuit(bfhe, ctx, "mult_32x32")
    v = VECTORS["mult_32x32"]["vectors"][0]
    ref = None
    try:
        for form in (8, 0, 32, 64, 32, 64, 16, 4, 0):
            ctx.dbg_set_gates_per_cta(form)
            assert _run_encrypted(c, v, seed=11, verify=False) == v["golden"], form
            slab = c.download_slab()
            if ref is None:
                ref = slab
            else:
                bad = np.nonzero((slab != ref).any(axis=1))[0]
                assert bad.size == 0, "form %d: %d wire ciphertexts differ, first row %d" % (form, bad.size, bad[0])
    finally:
        ctx.dbg_set_gates_per_cta(0)
