"""Row f-3: OpenFHE 1.0.x cereal-JSON key / ciphertext exchange (host/openfhe_json.cpp).

What these tests pin: the reader and the writer agree with each other and with the documented layout, EVALUATION-format refresh
keys are converted with the engine's OpenFHE-convention host transform, and malformed or mismatching files are refused.  What they
cannot pin: agreement with a file written by a real OpenFHE 1.0.x build (not available here; INTEGRATION.md "Closing the parity gap").
"""
import json
import os

import numpy as np
import pytest


def _ctx(B, paramset, method):
    c = B.Context(paramset, method, -1)
    c.keygen(11)
    c.btkeygen(12)
    return c


@pytest.mark.parametrize("method", ["GINX"])  # a TOY AP refresh key is 17 M coefficients of JSON: see test_ap_order_small
def test_key_roundtrip_toy(B, tmp_path, method):
    a = _ctx(B, B.TOY, getattr(B, method))
    paths = {w: str(tmp_path / ("k%d.json" % w)) for w in (B.OFHE_SECRET_KEY, B.OFHE_REFRESH_KEY, B.OFHE_SWITCH_KEY)}
    for w, p in paths.items():
        a.export_openfhe_json(w, p)
    b = B.Context(B.TOY, getattr(B, method), -1)
    for w, p in paths.items():
        b.import_openfhe_json(w, p)
    assert np.array_equal(a.export_keys(), b.export_keys())
    # the imported context is usable: host encrypt / decrypt agree across the two
    bits = np.array([0, 1, 1, 0, 1], dtype=np.uint8)
    assert b.decrypt(a.encrypt(bits, seed=5)).tolist() == bits.tolist()


def test_refresh_key_layout_and_eval_form(B, tmp_path):
    """the file holds [1][2][n] RGSW keys of [2 dG][2] polynomials in EVALUATION format: check one polynomial against a direct O(N^2)
    evaluation at psi^(2 bitrev(j) + 1), psi the smallest primitive 2N-th root -- OpenFHE's ChineseRemainderTransformFTT order"""
    a = _ctx(B, B.TOY, B.GINX)
    path = str(tmp_path / "bk.json")
    a.export_openfhe_json(B.OFHE_REFRESH_KEY, path)
    doc = json.load(open(path))
    k = doc["value0"]["ptr_wrapper"]["data"]["k"]
    p = a.p
    assert len(k) == 1 and len(k[0]) == 2 and len(k[0][0]) == p.n
    rg = k[0][1][3]["ptr_wrapper"]["data"]["elements"]  # the -1 key of secret coefficient 3
    assert len(rg) == 2 * p.dG and len(rg[0]) == 2
    poly = rg[2][1]
    assert poly["f"] == 0 and poly["v"]["ptr_wrapper"]["data"]["m"] == p.Q
    ev = np.array(poly["v"]["ptr_wrapper"]["data"]["v"], dtype=object)
    # coefficient form of the same polynomial from the BFHEKEY1 blob: header, sk, then BK [i][sign][row][col][N]
    blob = a.export_keys()
    N, Q = p.N, int(p.Q)
    words = np.frombuffer(blob.tobytes(), dtype=np.uint32)
    per_key = 2 * p.dG * 2
    # locate the BK region: it ends where the KSK starts; use sizes from the header-independent formula
    bk_words = p.n * 2 * per_key * N
    ksk_bytes = N * p.baseKS * p.dKS * (p.n + 1) * (2 if p.qKS <= 65536 else 4)
    bk0 = (blob.size - ksk_bytes) // 4 - bk_words
    idx = ((3 * 2 + 1) * per_key + 2 * 2 + 1) * N
    coef = [int(x) for x in words[bk0 + idx: bk0 + idx + N]]
    # smallest primitive 2N-th root of unity mod Q
    def is_prim(g):
        return pow(g, N, Q) == Q - 1
    gen = next(g for g in range(2, 2000) if pow(g, (Q - 1) // 2, Q) == Q - 1 and all(pow(g, (Q - 1) // f, Q) != 1 for f in _factors(Q - 1)))
    root = pow(gen, (Q - 1) // (2 * N), Q)
    roots = sorted(pow(root, e, Q) for e in range(1, 2 * N, 2))
    psi = roots[0]
    assert is_prim(psi)
    bits = N.bit_length() - 1
    for j in (0, 1, 2, 5, N // 2, N - 1):
        e = 2 * int(format(j, "0%db" % bits)[::-1], 2) + 1
        x = pow(psi, e, Q)
        val = 0
        for cdeg in reversed(coef):
            val = (val * x + cdeg) % Q
        assert int(ev[j]) == val, j


def _factors(m):
    out, d = [], 2
    while d * d <= m:
        if m % d == 0:
            out.append(d)
            while m % d == 0:
                m //= d
        d += 1
    if m > 1:
        out.append(m)
    return out


def test_coefficient_format_and_string_integers(B, tmp_path):
    """polynomials flagged COEFFICIENT ("f": 1) are taken as they are; integers may be decimal strings or {"v": n} objects"""
    a = _ctx(B, B.TOY, B.GINX)
    src = str(tmp_path / "bk.json")
    a.export_openfhe_json(B.OFHE_REFRESH_KEY, src)
    doc = json.load(open(src))
    ref = np.frombuffer(a.export_keys().tobytes(), dtype=np.uint8)
    # rewrite polynomial [0][0][0] elements[0][0] in coefficient form with string integers
    p = a.p
    blob = a.export_keys()
    words = np.frombuffer(blob.tobytes(), dtype=np.uint32)
    per_key = 2 * p.dG * 2
    ksk_bytes = p.N * p.baseKS * p.dKS * (p.n + 1) * (2 if p.qKS <= 65536 else 4)
    bk0 = (blob.size - ksk_bytes) // 4 - p.n * 2 * per_key * p.N
    poly = doc["value0"]["ptr_wrapper"]["data"]["k"][0][0][0]["ptr_wrapper"]["data"]["elements"][0][0]
    poly["f"] = 1
    poly["v"]["ptr_wrapper"]["data"]["v"] = [str(int(x)) for x in words[bk0: bk0 + p.N]]
    poly["v"]["ptr_wrapper"]["data"]["m"] = {"v": int(p.Q)}
    dst = str(tmp_path / "bk2.json")
    json.dump(doc, open(dst, "w"))
    b = B.Context(B.TOY, B.GINX, -1)
    b.import_openfhe_json(B.OFHE_SECRET_KEY, _export(a, B.OFHE_SECRET_KEY, tmp_path))
    b.import_openfhe_json(B.OFHE_REFRESH_KEY, dst)
    b.import_openfhe_json(B.OFHE_SWITCH_KEY, _export(a, B.OFHE_SWITCH_KEY, tmp_path))
    assert np.array_equal(np.frombuffer(b.export_keys().tobytes(), dtype=np.uint8), ref)


def _export(ctx, what, tmp_path):
    p = str(tmp_path / ("x%d.json" % what))
    ctx.export_openfhe_json(what, p)
    return p


def test_switch_key_split_layout(B, tmp_path):
    """later 1.0.x LWESwitchingKeyImpl: {"a": [N][baseKS][dKS] vectors, "b": [N][baseKS][dKS] integers}"""
    a = _ctx(B, B.TOY, B.GINX)
    doc = json.load(open(_export(a, B.OFHE_SWITCH_KEY, tmp_path)))
    k = doc["value0"]["ptr_wrapper"]["data"]["k"]
    split = {"cereal_class_version": 1,
             "a": [[[e["a"] for e in row] for row in blk] for blk in k],
             "b": [[[e["b"] for e in row] for row in blk] for blk in k]}
    doc["value0"]["ptr_wrapper"]["data"] = split
    dst = str(tmp_path / "ks_split.json")
    json.dump(doc, open(dst, "w"))
    b = B.Context(B.TOY, B.GINX, -1)
    b.import_openfhe_json(B.OFHE_SECRET_KEY, _export(a, B.OFHE_SECRET_KEY, tmp_path))
    b.import_openfhe_json(B.OFHE_REFRESH_KEY, _export(a, B.OFHE_REFRESH_KEY, tmp_path))
    b.import_openfhe_json(B.OFHE_SWITCH_KEY, dst)
    assert np.array_equal(a.export_keys(), b.export_keys())


def test_ciphertext_roundtrip_and_std128_secret_key(B, tmp_path):
    a = B.Context(B.STD128_OPT, B.GINX, -1)
    a.keygen(3)
    cts = a.encrypt([1, 0, 1], seed=4)
    path = str(tmp_path / "ct.json")
    for row in cts:
        a.export_openfhe_ct_json(row, path)
        doc = json.load(open(path))["value0"]["ptr_wrapper"]["data"]
        assert len(doc["a"]["v"]) == a.p.n and doc["a"]["m"] == a.p.q and doc["b"] == int(row[a.p.n])
        assert np.array_equal(a.import_openfhe_ct_json(path), row)
    sk = str(tmp_path / "sk.json")
    a.export_openfhe_json(B.OFHE_SECRET_KEY, sk)
    vals = set(json.load(open(sk))["value0"]["ptr_wrapper"]["data"]["s"]["v"])
    assert vals <= {0, 1, int(a.p.qKS) - 1}  # ternary key stored modulo qKS
    b = B.Context(B.STD128_OPT, B.GINX, -1)
    b.import_openfhe_json(B.OFHE_SECRET_KEY, sk)
    assert b.decrypt(cts).tolist() == [1, 0, 1]


def test_ap_order_small(B, tmp_path):
    """AP refresh key: file order [i][j][k] with the unused j = 0 slot a null pointer.  A hand-made document holding only the first secret
    coefficient must be refused with a count error that shows the null slots were skipped, not counted"""
    a = B.Context(B.TOY, B.AP, -1)
    src = str(tmp_path / "bk_ap.json")
    p = a.p
    poly = {"cereal_class_version": 1, "v": {"ptr_wrapper": {"valid": 1, "data": {"v": [0] * p.N, "m": int(p.Q)}}}, "f": 1, "p": {"ptr_wrapper": {"id": 1}}}
    rg = {"ptr_wrapper": {"id": 2, "data": {"elements": [[poly, poly] for _ in range(2 * p.dG)]}}}
    doc = {"value0": {"ptr_wrapper": {"id": 1, "data": {"k": [[[{"ptr_wrapper": {"id": 0}} if j == 0 else rg for _ in range(p.dR)] for j in range(p.baseR)]]}}}}
    json.dump(doc, open(src, "w"))
    b = B.Context(B.TOY, B.AP, -1)
    with pytest.raises(B.BfheError) as ei:
        b.import_openfhe_json(B.OFHE_REFRESH_KEY, src)
    assert ei.value.code == B.ERR_FORMAT and "expected %d" % (p.n * (p.baseR - 1) * p.dR) in str(ei.value)
    assert "%d RGSW" % ((p.baseR - 1) * p.dR) in str(ei.value)  # the null j = 0 slots were not counted


def test_refusals(B, tmp_path):
    a = _ctx(B, B.TOY, B.GINX)
    b = B.Context(B.TOY, B.GINX, -1)
    bad = str(tmp_path / "bad.json")
    open(bad, "w").write('{"value0": {"ptr_wrapper": {"id": 1, "data": {"s": {"v": [0, 1, 5], "m": 512}}}}}')
    with pytest.raises(B.BfheError) as ei:
        b.import_openfhe_json(B.OFHE_SECRET_KEY, bad)
    assert ei.value.code == B.ERR_FORMAT
    open(bad, "w").write('{"value0": [1, 2')
    with pytest.raises(B.BfheError) as ei:
        b.import_openfhe_json(B.OFHE_SECRET_KEY, bad)
    assert ei.value.code == B.ERR_FORMAT
    with pytest.raises(B.BfheError) as ei:
        b.import_openfhe_json(B.OFHE_SECRET_KEY, str(tmp_path / "missing.json"))
    assert ei.value.code == B.ERR_IO
    # a refresh key over another ring modulus (OpenFHE's other FirstPrime outcome) is refused with a precise message
    doc = json.load(open(_export(a, B.OFHE_REFRESH_KEY, tmp_path)))
    doc["value0"]["ptr_wrapper"]["data"]["k"][0][0][0]["ptr_wrapper"]["data"]["elements"][0][0]["v"]["ptr_wrapper"]["data"]["m"] = 134246401
    json.dump(doc, open(bad, "w"))
    with pytest.raises(B.BfheError) as ei:
        b.import_openfhe_json(B.OFHE_REFRESH_KEY, bad)
    assert ei.value.code == B.ERR_FORMAT and "134246401" in str(ei.value)
    # a STD128_OPT file offered to a TOY context
    c = B.Context(B.STD128_OPT, B.GINX, -1)
    c.keygen(1)
    with pytest.raises(B.BfheError):
        b.import_openfhe_json(B.OFHE_SECRET_KEY, _export(c, B.OFHE_SECRET_KEY, tmp_path))


@pytest.fixture
def B(bfhe):
    return bfhe
