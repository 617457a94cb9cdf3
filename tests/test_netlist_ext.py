"""Host logic added in round 2 (no GPU): the Bristol front-end on committed files, the gate types beyond the
reference's six -- NAND / NOR / XNOR / *_FAST as single bootstraps, DFF with real multi-Clock semantics, LUT3 / LUT4 lowered to 2-input
gates (row f-4; src/gate.h:51, stubs at src/gate.cpp:217-225) -- dumpNetList / dumpGates (src/circuit.cpp:844-865), and the planner
fixes (NOT chains, wave capacity 1)."""
import itertools
import os

import numpy as np
import pytest

from helpers import GOLDEN, VECTORS, load_circuit, oracle_run_plan

BRISTOL = os.path.join(GOLDEN, "bristol")


@pytest.fixture(scope="module")
def hctx(bfhe):
    return bfhe.Context(bfhe.TOY, bfhe.GINX, device=-1)


def _plain(c, inputs):
    c.Reset()
    c.setPlaintext(True)
    c.SetInput(inputs)
    return c.Clock()[0]


def test_read_bristol_old_header(bfhe, hctx):
    """old Bristol header 'ngates nwires / n1 n2 n3' (src/analyze.cpp:129-150): the parsed netlist equals the hand-written arrays."""
    c = bfhe.Circuit(hctx)
    c.ReadBristol(os.path.join(BRISTOL, "old_tiny.txt"), new_format=False)
    nl = c.get_netlist()
    K = bfhe
    assert nl["kind"].tolist() == [K.K_INPUT] * 4 + [K.K_XOR, K.K_AND, K.K_XOR, K.K_OUTPUT]
    assert nl["in0"].tolist() == [0, 0, 1, 1, 0, 1, 4, 6] and nl["in1"].tolist()[:7] == [0, 1, 0, 1, 2, 3, 5]
    assert nl["out"].tolist() == [0, 1, 2, 3, 4, 5, 6, 0] and nl["n_wires"] == 7
    assert nl["in_bits"].tolist() == [2, 2] and nl["out_bits"] == 1
    for a in itertools.product((0, 1), repeat=4):
        assert _plain(c, [list(a[:2]), list(a[2:])]) == [(a[0] ^ a[2]) ^ (a[1] & a[3])]
    assert c.info()["bootstraps"] == 7 and c.info()["levels"] == 4  # XOR = 3 bootstraps in 2 levels; AND || first XOR


def test_read_bristol_new_header_with_eqw(bfhe, hctx):
    """'Bristol fashion' header (niv n1 .. / nov m1 ..) and an EQW wire copy (src/analyze.cpp:152-180,273-280: the reference counts EQW
    but its assembler cannot emit it, App. D): EQW = NOT(NOT(x)), which the planner folds into operand flags."""
    c = bfhe.Circuit(hctx)
    c.ReadBristol(os.path.join(BRISTOL, "new_tiny_eqw.txt"), new_format=True)
    assert c.dumpGateCount() == dict(input=4, output=2, **{"and": 1, "or": 0, "xor": 1, "not": 3})
    for a in itertools.product((0, 1), repeat=4):
        assert _plain(c, [list(a[:2]), list(a[2:])]) == [a[0] & a[2], (1 - a[1]) ^ a[3]]
    # the NOT-NOT pair of EQW feeds an output: it must alias the AND's row, not become two chained EvalNOTs in one batch
    misc = c.plan_misc()
    for L in range(misc["n_levels"]):
        assert len(c.plan_misc(L)["nots"]) == 0
    assert not (int(misc["out_rows"][0]) >> 31)


def test_bristol_fixture_matches_npz(bfhe, hctx, tmp_path):
    """ReadBristol on a file written back from a golden netlist reproduces the committed .npz arrays (old format, 32-bit comparator)."""
    name = "comparator_32bit_unsigned_lt"
    d = np.load(os.path.join(GOLDEN, "circuits", name + ".npz"))
    kind, in0, in1, out = d["kind"], d["in0"], d["in1"], d["out"]
    ops = {bfhe.K_AND: "AND", bfhe.K_XOR: "XOR", bfhe.K_OR: "OR"}
    body = []
    for k, a, b, o in zip(kind, in0, in1, out):
        if k in ops:
            body.append("2 1 %d %d %d %s" % (a, b, o, ops[k]))
        elif k == bfhe.K_NOT:
            body.append("1 1 %d %d INV" % (a, o))
    p = tmp_path / "cmp.txt"
    p.write_text("%d %d\n%d %d %d\n\n%s\n" % (len(body), int(d["n_wires"]), d["in_bits"][0], d["in_bits"][1], int(d["out_bits"]), "\n".join(body)))
    c = bfhe.Circuit(hctx)
    c.ReadBristol(p)
    nl = c.get_netlist()
    for key in ("kind", "in0", "in1", "out"):
        assert np.array_equal(nl[key], d[key]), key
    assert c.info() == VECTORS[name]["info"]


OUT_EXT = """R0 = LOAD(In1,0)
R1 = LOAD(In1,1)
R2 = NAND(R0, R1)
R3 = NOR(R0, R1)
R4 = XNOR(R0, R1)
R5 = XOR_FAST(R0, R1)
R6 = XNOR_FAST(R0, R1)
Out0 = STORE(R2)
Out1 = STORE(R3)
Out2 = STORE(R4)
Out3 = STORE(R5)
Out4 = STORE(R6)
"""


def test_native_gate_kinds(bfhe, hctx, tmp_path):
    """NAND / NOR / XOR_FAST / XNOR_FAST are ONE bootstrap each with OpenFHE's gate constants; XNOR is the composite XOR read through a
    free NOT (what EvalBinGate(XNOR) computes)."""
    p = tmp_path / "ext.out"
    p.write_text(OUT_EXT)
    c = bfhe.Circuit(hctx)
    c.ReadFile(p)
    assert c.dumpGateCountEx() == dict(dff=0, lut3=0, lut4=0, nand=1, nor=1, xnor=1, xor_fast=1, xnor_fast=1)
    assert c.info()["bootstraps"] == 4 + 3 and c.info()["gates"] == 5
    for a, b in itertools.product((0, 1), repeat=2):
        assert _plain(c, [[a, b]]) == [1 - (a & b), 1 - (a | b), 1 - (a ^ b), a ^ b, 1 - (a ^ b)]
    ops = sorted(int(g["op"]) & 0xff for g in c.level_plan(1, 0, 1)[0])
    assert ops == sorted([bfhe.NAND, bfhe.NOR, bfhe.XOR_FAST, bfhe.XNOR_FAST, bfhe.AND, bfhe.AND])
    # emitted text parses back to the same function
    q = tmp_path / "ext2.out"
    c.write_out(q)
    c2 = bfhe.Circuit(hctx)
    c2.ReadFile(q)
    for a, b in itertools.product((0, 1), repeat=2):
        assert _plain(c2, [[a, b]]) == _plain(c, [[a, b]])


COUNTER = """# 2-bit counter with enable: Q0' = Q0 xor en, Q1' = Q1 xor (Q0 and en)
R0 = LOAD(In1,0)
R1 = DFF(R5)
R2 = DFF(R6)
R5 = XOR(R1, R0)
R3 = AND(R1, R0)
R6 = XOR(R2, R3)
Out0 = STORE(R1)
Out1 = STORE(R2)
"""


def test_dff_counter_plaintext(bfhe, hctx, tmp_path):
    """DFF: Q shows the state latched by the previous Clock(); Reset() returns the flip-flops to 0; a circuit with flip-flops may be
    clocked repeatedly, a combinational one only once (src/circuit.cpp:538-541)."""
    p = tmp_path / "counter.out"
    p.write_text(COUNTER)
    c = bfhe.Circuit(hctx)
    c.ReadFile(p)
    assert c.dumpGateCountEx()["dff"] == 2
    c.Reset(); c.setPlaintext(True); c.SetInput([[1]])
    seen = [c.Clock()[0] for _ in range(6)]
    assert seen == [[0, 0], [1, 0], [0, 1], [1, 1], [0, 0], [1, 0]]
    c.SetInput([[0]])  # enable low: holds
    assert c.Clock()[0] == [0, 1] and c.Clock()[0] == [0, 1]
    c.Reset(); c.setPlaintext(True); c.SetInput([[1]])
    assert c.Clock()[0] == [0, 0]
    quads = c.dff_plan()
    assert quads.shape == (2, 4) and len(set(quads[:, 1].tolist()) | set(quads[:, 2].tolist()) | set(quads[:, 3].tolist())) == 6


def test_dff_shift_register_reads_old_state(bfhe, hctx, tmp_path):
    p = tmp_path / "shift.out"
    p.write_text("R0 = LOAD(In1,0)\nR1 = DFF(R0)\nR2 = DFF(R1)\nR3 = DFF(R2)\nOut0 = STORE(R1)\nOut1 = STORE(R2)\nOut2 = STORE(R3)\n")
    c = bfhe.Circuit(hctx)
    c.ReadFile(p)
    c.Reset(); c.setPlaintext(True)
    outs = []
    for bit in (1, 0, 1, 1, 0):
        c.SetInput([[bit]])
        outs.append(c.Clock()[0])
    assert outs == [[0, 0, 0], [1, 0, 0], [0, 1, 0], [1, 0, 1], [1, 1, 0]]


def test_dff_counter_encrypted_via_oracle(bfhe, orc, hctx, tmp_path):
    """The product's clocked plan executed by the oracle (CPU stand-in for the kernels): 5 clocks of the counter decrypt to 0,1,2,3,0."""
    p = tmp_path / "counter.out"
    p.write_text(COUNTER)
    ctx = bfhe.Context(bfhe.TOY, bfhe.GINX, device=-1)
    ctx.keygen(31); ctx.btkeygen(32)
    o = orc.Oracle(orc.TOY, orc.GINX)
    o.import_keys(ctx.export_keys())
    c = bfhe.Circuit(ctx)
    c.ReadFile(p)
    state = None
    seen = []
    for clk in range(5):
        out, slab = oracle_run_plan(c, o, [[1]], seed=40 + clk, state=state)
        state = slab
        seen.append(out)
    assert seen == [[0, 0], [1, 0], [0, 1], [1, 1], [0, 0]]


def _lut_text(k, tt):
    ins = "".join("R%d = LOAD(In1,%d)\n" % (i, i) for i in range(k))
    args = ", ".join("R%d" % i for i in range(k))
    return ins + "R9 = LUT%d(%s, 0x%X)\nOut0 = STORE(R9)\n" % (k, args, tt)


def test_lut3_all_tables_and_lut4_sample(bfhe, hctx, tmp_path):
    """LUT3 over all 254 non-constant tables and LUT4 over a sample: the lowered netlist computes the table; at most 5 / 13 bootstraps...
    (XOR cofactors count 3 each, so the bound in bootstraps is looser than the bound in gates)."""
    p = tmp_path / "lut.out"
    c = bfhe.Circuit(hctx)
    for tt in range(1, 255):
        p.write_text(_lut_text(3, tt))
        c.ReadFile(p)
        assert c.dumpGateCountEx()["lut3"] == 1
        for a in itertools.product((0, 1), repeat=3):
            assert _plain(c, [list(a)]) == [(tt >> (a[0] | a[1] << 1 | a[2] << 2)) & 1], (tt, a)
        assert c.info()["bootstraps"] <= 9
    rng = np.random.default_rng(1)
    for tt in [0x6996, 0x8000, 0xFFFE, 0x0001, 0xAAAA, 0x5555, 0xCCCC, 0xF0F0, 0x00FF, 0xCAFE] + rng.integers(1, 0xFFFF, 40).tolist():
        p.write_text(_lut_text(4, int(tt)))
        c.ReadFile(p)
        for a in itertools.product((0, 1), repeat=4):
            assert _plain(c, [list(a)]) == [(int(tt) >> (a[0] | a[1] << 1 | a[2] << 2 | a[3] << 3)) & 1], (tt, a)
    for k, tt in ((3, 0), (3, 0xFF), (4, 0xFFFF)):
        p.write_text(_lut_text(k, tt))
        with pytest.raises(bfhe.BfheError):
            c.ReadFile(p)


def test_lut_from_arrays(bfhe, hctx):
    K = bfhe
    kind = [K.K_INPUT, K.K_INPUT, K.K_INPUT, K.K_LUT3, K.K_OUTPUT]
    c = bfhe.Circuit(hctx)
    c.load_netlist_ex(kind, [0, 0, 0, 0, 3], [0, 1, 2, 1, 0], [0, 0, 0, 2, 0], [0] * 5, [0, 0, 0, 0xE8, 0], [0, 1, 2, 3, 0], 4, [3], 1)
    for a in itertools.product((0, 1), repeat=3):
        assert _plain(c, [list(a)]) == [int(sum(a) >= 2)]  # 0xE8 = majority


def test_not_chain_is_not_a_chain_of_launches(bfhe, hctx, tmp_path):
    """w1 = NOT(w0), w2 = NOT(w1) materialised in ONE batched EvalNOT launch raced.  Every materialised NOT now reads a
    bootstrapped (or input) row, never another NOT's row -- in verify mode too, where every NOT output owns a ciphertext."""
    p = tmp_path / "chain.out"
    p.write_text("R0 = LOAD(In1,0)\nR1 = LOAD(In1,1)\nR2 = AND(R0, R1)\nR3 = NOT(R2)\nR4 = NOT(R3)\nR5 = NOT(R4)\nR6 = OR(R4, R0)\n"
                 "Out0 = STORE(R3)\nOut1 = STORE(R4)\nOut2 = STORE(R5)\nOut3 = STORE(R6)\n")
    for verify in (False, True):
        c = bfhe.Circuit(hctx)
        c.ReadFile(p)
        c.Reset(); c.setPlaintext(True)
        if verify:  # the verify plan is built by SetInput, which then fails at the upload: there is no device here
            c.setVerify(True)
            with pytest.raises(bfhe.BfheError):
                c.SetInput([[1, 1]])
        else:
            c.SetInput([[1, 1]])
        misc = c.plan_misc()
        produced = set()
        for L in range(misc["n_levels"]):
            g, first, rpr = c.level_plan(L, 0, 1)
            boot_rows = set(g["out"].tolist())
            produced |= boot_rows
            nots = c.plan_misc(L)["nots"]
            for a, b in nots:
                assert int(a) in produced and int(b) not in produced  # input of every EvalNOT is a bootstrapped row of this or an earlier level
            assert len(set(nots[:, 1].tolist())) == len(nots)
        rows = [int(r) & 0x7fffffff for r in misc["out_rows"]]
        assert rows[0] == rows[2] and rows[1] != rows[0]  # NOT^3 shares NOT^1's row; NOT^2 aliases the AND's row
        if not verify:
            assert c.Clock()[0] == [0, 1, 0, 1]


def test_wave_capacity_one_terminates(bfhe, hctx):
    """cap = 1 with an XOR (unit weight 2) used to spin forever; the capacity is clamped to the heaviest unit."""
    c = load_circuit(bfhe, hctx, "adder_2bit")
    c.set_wave_capacity(1)
    misc = c.plan_misc()
    widths = [len(c.level_plan(L, 0, 1)[0]) for L in range(1, misc["n_levels"])]
    assert sum(widths) == c.info()["bootstraps"] and max(widths) <= 2
    v = VECTORS["adder_2bit"]["vectors"][0]
    assert _plain(c, v["inputs"]) == v["golden"]


def test_dump_netlist_and_gates_text(bfhe, hctx, tmp_path):
    """Circuit::dumpNetList / dumpGates (src/circuit.cpp:844-865): wire -> reader gate names in std::map order; gate names are
    '<KIND>:<gate number>' with the numbering of Circuit::ReadFile (src/circuit.cpp:154,179,209,...)."""
    p = tmp_path / "t.out"
    p.write_text("R0 = LOAD(In1,0)\nR1 = LOAD(In1,1)\nR10 = AND(R0, R1)\nR2 = NOT(R10)\nR3 = XOR(R2, R0)\nOut0 = STORE(R3)\n")
    c = bfhe.Circuit(hctx)
    c.ReadFile(p)
    assert c.dumpNetList() == ("Netlist \nBIT:0\nOUT:0\nR:0 AND:2 XOR:4\nR:1 AND:2\nR:10 NOT:3\nR:2 XOR:4\nR:3 OUTPUT:5\n")
    assert c.dumpGates() == "Inputlist \nINPUT:0\nINPUT:1\nAlllist \nAND:2\nNOT:3\nXOR:4\nOUTPUT:5\n"


def test_seed_contract(bfhe):
    """seed 0 = OS entropy (two key generations differ), non-zero seed = reproducible (tests / parity only)."""
    a, b = bfhe.Context(bfhe.TOY, bfhe.GINX, device=-1), bfhe.Context(bfhe.TOY, bfhe.GINX, device=-1)
    a.keygen(0); b.keygen(0)
    a.btkeygen(0); b.btkeygen(0)
    ka, kb = a.export_keys(), b.export_keys()
    assert not np.array_equal(ka, kb)
    a.keygen(5); b.keygen(5); a.btkeygen(6); b.btkeygen(6)
    assert np.array_equal(a.export_keys(), b.export_keys())
    c1, c2 = a.encrypt([1, 0, 1], seed=0), a.encrypt([1, 0, 1], seed=0)
    assert not np.array_equal(c1, c2) and a.decrypt(c1).tolist() == [1, 0, 1] and a.decrypt(c2).tolist() == [1, 0, 1]
    assert np.array_equal(a.encrypt([1, 0], seed=9), a.encrypt([1, 0], seed=9))
    # the ternary secret and the fresh masks are unbiased draws
    blob_sk = np.frombuffer(ka.tobytes()[128:128 + 4 * 64], dtype=np.int32)
    assert set(blob_sk.tolist()) <= {-1, 0, 1}
    cts = a.encrypt(np.zeros(400, dtype=np.uint8), seed=0)[:, :64].ravel()
    assert cts.max() < 512 and abs(cts.mean() - 255.5) < 6
