"""The C-ABI library loads and exports every symbol include/bfhe.h declares (no compute without a GPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "bfhe.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(bfhe_[a-z0-9_]+)\s*\(", txt)))


def test_header_and_library_agree(bfhe):
    syms = header_symbols()
    assert len(syms) >= 40
    L = ctypes.CDLL(bfhe.LIB_PATH)
    missing = [s for s in syms if not hasattr(L, s)]
    assert not missing, missing
    assert sorted(bfhe.ABI_SYMBOLS) == syms


def test_no_cpu_fallback(bfhe):
    """Without a device the host-side calls work and every Eval* entry point fails loudly."""
    import numpy as np
    ctx = bfhe.Context(bfhe.TOY, bfhe.GINX, device=-1)
    ctx.keygen(1)
    cts = ctx.encrypt([0, 1, 1])
    assert ctx.decrypt(cts).tolist() == [0, 1, 1]
    g = np.array([(bfhe.AND, 0, 1, 2)], dtype=bfhe.GATE_DTYPE)
    for call in (lambda: ctx.eval_bingate_batch(0, g), lambda: ctx.bootstrap_batch(0, [0], [1]),
                 lambda: ctx.eval_not_batch(0, [0], [1]), lambda: ctx.eval_bingate_host(g, cts, 1), lambda: ctx.slab(4)):
        with pytest.raises(bfhe.BfheError) as e:
            call()
        assert e.value.code == bfhe.ERR_CUDA
    c = bfhe.Circuit(ctx)
    c.load_npz(os.path.join(ROOT, "tests", "golden", "circuits", "adder_2bit.npz"))
    c.setEncrypted(True)
    with pytest.raises(bfhe.BfheError) as e:
        c.SetInput([[0, 1], [1, 0]])
    assert e.value.code in (bfhe.ERR_CUDA, bfhe.ERR_STATE)


def test_rejects_unsupported_parameter_sets(bfhe):
    """Circuit's ctor accepts only TOY / STD128_OPT and AP / GINX (src/circuit.cpp:69-86)."""
    for ps, m in ((1, bfhe.GINX), (4, bfhe.GINX), (bfhe.TOY, 2)):
        with pytest.raises(bfhe.BfheError):
            bfhe.Context(ps, m, device=-1)


def test_params_match_reference_table(bfhe):
    """GenerateBinFHEContext parameter table."""
    p = bfhe.Context(bfhe.STD128_OPT, bfhe.GINX, device=-1).p
    assert (p.n, p.N, p.q, p.Q, p.qKS, p.baseKS, p.dKS, p.baseG, p.dG, p.baseR, p.dR) == \
        (502, 1024, 1024, 134215681, 1 << 14, 128, 2, 128, 4, 32, 2)
    p = bfhe.Context(bfhe.TOY, bfhe.AP, device=-1).p
    assert (p.n, p.N, p.q, p.Q, p.qKS, p.baseKS, p.dKS, p.baseG, p.dG, p.baseR, p.dR) == \
        (64, 512, 512, 134215681, 134215681, 25, 6, 512, 3, 23, 2)


def test_key_blob_roundtrip_and_validation(bfhe, tmp_path):
    ctx = bfhe.Context(bfhe.TOY, bfhe.GINX, device=-1)
    ctx.keygen(3)
    ctx.btkeygen(4)
    blob = ctx.export_keys()
    p = str(tmp_path / "k.bin")
    ctx.save_keys(p)
    c2 = bfhe.Context(bfhe.TOY, bfhe.GINX, device=-1)
    c2.load_keys(p)
    assert (c2.export_keys() == blob).all()
    cts = ctx.encrypt([1, 0, 1], seed=9)
    assert c2.decrypt(cts).tolist() == [1, 0, 1]
    c3 = bfhe.Context(bfhe.TOY, bfhe.AP, device=-1)  # wrong method
    with pytest.raises(bfhe.BfheError) as e:
        c3.import_keys(blob)
    assert e.value.code == bfhe.ERR_FORMAT
    with pytest.raises(bfhe.BfheError):
        c2.import_keys(blob[:1000])
    # public blob (no secret key): decrypt must refuse
    c4 = bfhe.Context(bfhe.TOY, bfhe.GINX, device=-1)
    c4.import_keys(ctx.export_keys(include_sk=False))
    with pytest.raises(bfhe.BfheError):
        c4.decrypt(cts)
