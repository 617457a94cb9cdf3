"""Host logic of the circuit layer: parsing, levelisation, plaintext evaluation against every golden vector."""
import numpy as np
import pytest

from helpers import VECTORS, circuit_path, load_circuit


@pytest.fixture(scope="module")
def hctx(bfhe):
    return bfhe.Context(bfhe.STD128_OPT, bfhe.GINX, device=-1)


@pytest.mark.parametrize("name", sorted(VECTORS))
def test_plaintext_matches_reference_goldens(bfhe, hctx, name):
    """README.md:13-26 three-way check, leg (2): plaintext circuit evaluation == golden."""
    c = load_circuit(bfhe, hctx, name)
    for v in VECTORS[name]["vectors"]:
        c.Reset()
        c.setPlaintext(True)
        c.SetInput(v["inputs"])
        assert c.Clock()[0] == v["golden"], v["src"]
    assert c.info() == VECTORS[name]["info"]
    assert c.dumpGateCount() == VECTORS[name]["gate_count"]


def test_levelisation_matches_survey_table(bfhe, hctx):
    """SURVEY App. A [MEASURED]: bootstraps / bootstrap levels / max width under the XOR = 2 AND + OR rule."""
    exp = {"adder_2bit": (13, 4, 6), "parity": (24, 8, 8), "mult_32x32": (9133, 132, 1039),
           "AES-non-expanded": (82172, 420, 430), "AES-expanded": (66415, 416, 376), "md5": (71534, 3852, 45),
           "sha256": (354505, 9055, 1796), "comparator_32bit_signed_lt": (150, 22, 42)}
    for name, (b, l, w) in exp.items():
        i = load_circuit(bfhe, hctx, name).info()
        assert (i["bootstraps"], i["levels"], i["max_width"]) == (b, l, w), name


@pytest.mark.parametrize("name", ["adder_2bit", "parity", "comparator_32bit_signed_lteq"])
def test_out_format_roundtrip(bfhe, hctx, name, tmp_path):
    """Circuit::ReadFile grammar: emit '.out' text, parse it back, same netlist semantics."""
    c = load_circuit(bfhe, hctx, name)
    p = tmp_path / (name + "_FHE.out")
    c.write_out(p)
    txt = p.read_text()
    assert "LOAD(In1,0)" in txt and "STORE(" in txt and txt.startswith("# Max depth")
    c2 = bfhe.Circuit(hctx)
    c2.ReadFile(p)
    assert c2.info() == c.info()
    for v in VECTORS[name]["vectors"][:4]:
        c2.Reset(); c2.setPlaintext(True); c2.SetInput(v["inputs"])
        assert c2.Clock()[0] == v["golden"]


def test_out_parser_details(bfhe, hctx, tmp_path):
    p = tmp_path / "t.out"
    p.write_text("# number input1 bits 2\n# comment\nR0 = LOAD(In1,0)\nR1 = LOAD(In1,1)\nR7 = NOT(R0) !depth = 3\n"
                 "R2 = OR(R7, R1)\nR3 = XOR(R2, R0)  !depth = 1\nBOOT something\nOut0 = STORE(R3) ! depth = 0\nOut1 = STORE(R7)\n")
    c = bfhe.Circuit(hctx)
    c.ReadFile(p)
    assert c.dumpGateCount() == dict(input=2, output=2, **{"and": 0, "or": 1, "xor": 1, "not": 1})
    for a in (0, 1):
        for b in (0, 1):
            c.Reset(); c.setPlaintext(True); c.SetInput([[a, b]])
            assert c.Clock()[0] == [(((1 - a) | b) ^ a), 1 - a]
    # the reference exits on a done circuit clocked twice (src/circuit.cpp:538-541)
    with pytest.raises(bfhe.BfheError):
        c.Clock()


@pytest.mark.parametrize("text,code", [
    ("R0 = LOAD(In1,0)\nR1 = AND(R0, R0)\nOut0 = STORE(R1)\n", "ERR_ALIAS"),        # EvalBinGate(ct, ct)
    ("R0 = LOAD(In1,0)\nR1 = AND(R0, R5)\nOut0 = STORE(R1)\n", "ERR_FORMAT"),       # undriven wire
    ("R0 = LOAD(In1,0)\nR0 = NOT(R0)\nOut0 = STORE(R0)\n", "ERR_FORMAT"),           # double assignment
    ("R0 = LOAD(In1,0)\nR1 = AND(R0 R0)\n", "ERR_FORMAT"),                          # parse error
    ("R0 = LOAD(In1,0)\nR1 = AND(R0, R2)\nR2 = NOT(R1)\nOut0 = STORE(R2)\n", "ERR_FORMAT"),  # loop
])
def test_malformed_circuits_are_rejected_at_load(bfhe, hctx, tmp_path, text, code):
    """The reference hangs or exits on these; here they fail at load with an error code."""
    p = tmp_path / "bad.out"
    p.write_text(text)
    c = bfhe.Circuit(hctx)
    with pytest.raises(bfhe.BfheError) as e:
        c.ReadFile(p)
    assert e.value.code == getattr(bfhe, code)


def test_and_of_wire_with_its_own_not_is_allowed(bfhe, hctx, tmp_path):
    """a and NOT(a) are distinct ciphertexts for OpenFHE (only pointer-equal inputs throw)."""
    p = tmp_path / "ok.out"
    p.write_text("R0 = LOAD(In1,0)\nR1 = NOT(R0)\nR2 = AND(R0, R1)\nOut0 = STORE(R2)\n")
    c = bfhe.Circuit(hctx)
    c.ReadFile(p)
    g, _, _ = c.level_plan(1, 0, 1)
    assert len(g) == 1 and g[0]["in0"] == g[0]["in1"] and (int(g[0]["op"]) & (bfhe.NEG0 | bfhe.NEG1)) == bfhe.NEG1


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_level_partition_covers_every_gate_once(bfhe, hctx, world):
    c = load_circuit(bfhe, hctx, "comparator_32bit_unsigned_lt")
    c.set_sharding(0, world)
    misc = c.plan_misc()
    seen_rows = set()
    total = 0
    for L in range(misc["n_levels"]):
        first = None
        for r in range(world):
            g, f, rpr = c.level_plan(L, r, world)
            first = f
            assert len(g) <= rpr
            if len(g):  # each rank's outputs are one contiguous run inside its own chunk of the level block
                assert np.array_equal(g["out"], f + r * rpr + np.arange(len(g)))
            for row in g["out"]:
                assert row not in seen_rows
                seen_rows.add(int(row))
            if L > 0:
                total += len(g)
            assert all(int(x) < f for x in g["in0"]) or L == 0
    assert total == c.info()["bootstraps"]
    assert misc["total_rows"] > max(seen_rows)


@pytest.mark.parametrize("name,cap", [("adder_2bit", 2), ("parity", 3), ("comparator_32bit_signed_lt", 5), ("mult_32x32", 148),
                                      ("AES-non-expanded", 148)])
def test_wave_packing_is_a_valid_schedule(bfhe, hctx, name, cap):
    """bfhe_circuit_set_wave_capacity: every wave holds at most `cap` bootstraps (whole throughput waves of 4*cap when at least
    8*cap gates are ready), every operand row is produced by an earlier wave, every bootstrap is placed exactly once, and the
    ASAP statistics the reference-shaped info() reports do not change."""
    c = load_circuit(bfhe, hctx, name)
    info0, asap_levels = c.info(), c.plan_misc()["n_levels"] - 1
    c.set_wave_capacity(cap)
    assert c.info() == info0
    misc = c.plan_misc()
    total, produced = 0, set()
    for L in range(misc["n_levels"]):
        g, first, rpr = c.level_plan(L, 0, 1)
        if L > 0:
            assert 0 < len(g) and (len(g) <= cap or (len(g) % (4 * cap) == 0)), (L, len(g))
            total += len(g)
            for x in g:
                assert int(x["in0"]) in produced and int(x["in1"]) in produced, (L, x)
        nots = c.plan_misc(L)["nots"]
        for row in g["out"]:
            produced.add(int(row))
        for a, b in nots:
            assert int(a) in produced
            produced.add(int(b))
    assert total == info0["bootstraps"]
    waves = misc["n_levels"] - 1
    assert waves >= asap_levels and waves >= -(-info0["bootstraps"] // max(cap, 1)) // 4
    if name == "AES-non-expanded":  # 420 ASAP levels cost ~790 one-gate-per-SM launches; packed: close to bootstraps / 148 = 556
        assert waves <= 600, waves
    c.set_wave_capacity(0)
    assert c.plan_misc()["n_levels"] - 1 == asap_levels


@pytest.mark.parametrize("name,cap", [("adder_2bit", 2), ("parity", 2), ("comparator_32bit_signed_lteq", 7)])
def test_wave_packing_keeps_ciphertexts(bfhe, orc, hctx, name, cap):
    """Same keys, same fresh encryptions: the packed schedule produces bit-identical output ciphertexts (oracle as executor)."""
    from helpers import oracle_run_plan
    o = orc.Oracle(orc.TOY, orc.GINX)
    o.keygen(5)
    v = VECTORS[name]["vectors"][0]
    c = load_circuit(bfhe, bfhe.Context(bfhe.TOY, bfhe.GINX, device=-1), name)
    outs0, slab0 = oracle_run_plan(c, o, v["inputs"], seed=3)
    rows0 = [int(r) & 0x7fffffff for r in c.plan_misc()["out_rows"]]
    c.set_wave_capacity(cap)
    outs1, slab1 = oracle_run_plan(c, o, v["inputs"], seed=3)
    rows1 = [int(r) & 0x7fffffff for r in c.plan_misc()["out_rows"]]
    assert outs0 == outs1 == v["golden"]
    for a, b in zip(rows0, rows1):
        assert np.array_equal(slab0[a], slab1[b])
