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


def test_config5_aes128_expanded(bfhe):
    """old-bristol AES-128 with the expanded key as input 2 (src/test_aes.cpp:187-202): both KATs, encrypted."""
    ctx = shared_keys(bfhe, bfhe.STD128_OPT, bfhe.GINX, 0)
    c = load_circuit(bfhe, ctx, "AES-expanded")
    for t, v in enumerate(VECTORS["AES-expanded"]["vectors"]):
        assert _run_encrypted(c, v, seed=20 + t, verify=False) == v["golden"], v["src"]


def test_config5_md5(bfhe):
    """old-bristol MD5, the four KATs of src/test_md5.cpp:203-228, encrypted (71 534 bootstraps in 3 852 levels each)."""
    ctx = shared_keys(bfhe, bfhe.STD128_OPT, bfhe.GINX, 0)
    c = load_circuit(bfhe, ctx, "md5")
    for t, v in enumerate(VECTORS["md5"]["vectors"]):
        assert _run_encrypted(c, v, seed=30 + t, verify=(t == 0)) == v["golden"], v["src"]
        assert c.stats()["verify_mismatches"] == 0


def test_config5_sha256(bfhe):
    """SHA-256 (new-bristol sha256.txt, IV as input 2), the four KATs of src/test_sha256.cpp:205-238, encrypted
    (354 505 bootstraps in 9 055 levels each: the slow test of the suite)."""
    ctx = shared_keys(bfhe, bfhe.STD128_OPT, bfhe.GINX, 0)
    c = load_circuit(bfhe, ctx, "sha256")
    for t, v in enumerate(VECTORS["sha256"]["vectors"]):
        assert _run_encrypted(c, v, seed=40 + t, verify=False) == v["golden"], v["src"]


EXT_OUT = """R0 = LOAD(In1,0)
R1 = LOAD(In1,1)
R2 = LOAD(In1,2)
R3 = NAND(R0, R1)
R4 = NOR(R1, R2)
R5 = XNOR(R3, R4)
R6 = XOR_FAST(R5, R0)
R7 = XNOR_FAST(R6, R2)
R8 = LUT3(R0, R1, R2, 0xE8)
R9 = NOT(R8)
R10 = NOT(R9)
R11 = NOT(R10)
R12 = AND(R10, R7)
Out0 = STORE(R3)
Out1 = STORE(R5)
Out2 = STORE(R7)
Out3 = STORE(R8)
Out4 = STORE(R9)
Out5 = STORE(R10)
Out6 = STORE(R11)
Out7 = STORE(R12)
"""


@pytest.mark.parametrize("ps,m", [("STD128_OPT", "GINX"), ("TOY", "AP")])
def test_extended_gate_kinds_and_not_chains(bfhe, orc, tmp_path, ps, m):
    """Rows f-4 + ADVICE: native NAND / NOR / XOR_FAST / XNOR_FAST, composite XNOR, a lowered LUT3 and a NOT-NOT-NOT chain feeding
    outputs, encrypted with gate-by-gate verify; every wire ciphertext equals the oracle executing the same plan."""
    import itertools
    ctx = shared_keys(bfhe, getattr(bfhe, ps), getattr(bfhe, m), 0)
    p = tmp_path / "ext.out"
    p.write_text(EXT_OUT)
    c = bfhe.Circuit(ctx)
    c.ReadFile(p)
    for t, (a, b, d) in enumerate(itertools.product((0, 1), repeat=3)):
        n3, n4 = 1 - (a & b), 1 - (b | d)
        r5 = 1 - (n3 ^ n4)
        r7 = 1 - ((r5 ^ a) ^ d)
        maj = int(a + b + d >= 2)
        want = [n3, r5, r7, maj, 1 - maj, maj, 1 - maj, maj & r7]
        for verify in (True, False):
            assert _run_encrypted(c, {"inputs": [[a, b, d]]}, seed=50 + t, verify=verify) == want, (a, b, d, verify)
            assert c.stats()["verify_mismatches"] == 0
    o = orc.Oracle(getattr(orc, ps), getattr(orc, m))
    o.import_keys(ctx.export_keys())
    for verify in (True, False):
        _run_encrypted(c, {"inputs": [[1, 0, 1]]}, seed=77, verify=verify)
        slab = c.download_slab()
        fresh = ctx.encrypt([1, 0, 1], seed=77)
        _, ref = oracle_run_plan(c, o, [[1, 0, 1]], fresh=fresh)
        w = ctx.p.ct_words
        assert np.array_equal(slab[:, :w], ref[:, :w]), verify


COUNTER = """R0 = LOAD(In1,0)
R1 = DFF(R5)
R2 = DFF(R6)
R5 = XOR(R1, R0)
R3 = AND(R1, R0)
R6 = XOR(R2, R3)
Out0 = STORE(R1)
Out1 = STORE(R2)
"""


@pytest.mark.parametrize("ps,graph", [("TOY", True), ("STD128_OPT", True), ("TOY", False)])
def test_dff_counter_encrypted(bfhe, tmp_path, ps, graph):
    """Clocked circuit on the GPU: a 2-bit counter with enable, six Clock() calls on one SetInput, then a hold, then Reset; verify mode
    checks every wire (the Q rows included) on every clock."""
    ctx = shared_keys(bfhe, getattr(bfhe, ps), bfhe.GINX, 0)
    p = tmp_path / "counter.out"
    p.write_text(COUNTER)
    c = bfhe.Circuit(ctx)
    c.ReadFile(p)
    c.use_graph(graph)
    for verify in (True, False):
        c.Reset(); c.setEncrypted(True); c.setVerify(verify)
        c.SetInput([[1]], seed=5)
        seen = [c.Clock()[0] for _ in range(6)]
        assert seen == [[0, 0], [1, 0], [0, 1], [1, 1], [0, 0], [1, 0]], (verify, seen)
        c.SetInput([[0]], seed=6)
        assert c.Clock()[0] == [0, 1] and c.Clock()[0] == [0, 1]
        assert c.stats()["verify_mismatches"] == 0


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
    blind-rotation form forced -- one gate per CTA (8), four per CTA (4), one gate on a 2-CTA cluster (32), on a slot-sliced 4-CTA
    cluster (128) -- and twice with the cost model's own choice: every wire ciphertext must be bit-identical across all
    runs (a stale shared-memory read in the cluster form once produced valid-but-different ciphertexts in 1 of ~400k
    bootstraps; tools/race_hunt.py is the long-running version of this test)."""
    ctx = shared_keys(bfhe, bfhe.STD128_OPT, bfhe.GINX, 0)
    c = load_circuit(bfhe, ctx, "mult_32x32")
    v = VECTORS["mult_32x32"]["vectors"][0]
    ref = None
    try:
        for form in (8, 0, 32, 128, 32, 128, 8, 128, 4, 0):
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
