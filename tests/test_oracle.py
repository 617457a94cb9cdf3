"""Pins the oracle: algebraic properties of every stage, truth tables, and the reference's own decrypted
known-answer vectors for configs 1 and 2 (TB_adder_2bit, TB_parity) evaluated gate by gate on encrypted data.
(No ciphertext-level fixtures exist upstream -- SURVEY 8(c): "parity unpinned" -- so decrypted bits are the pin.)"""
import numpy as np
import pytest

from helpers import VECTORS, load_circuit, oracle_run_plan

_ORC = {}


def get_oracle(orc, ps, m, seed=5):
    k = (ps, m)
    if k not in _ORC:
        o = orc.Oracle(ps, m)
        o.keygen(seed)
        _ORC[k] = o
    return _ORC[k]


TT = {"OR": lambda a, b: a | b, "AND": lambda a, b: a & b, "NOR": lambda a, b: 1 - (a | b), "NAND": lambda a, b: 1 - (a & b),
      "XOR_FAST": lambda a, b: a ^ b, "XNOR_FAST": lambda a, b: 1 - (a ^ b), "XOR": lambda a, b: a ^ b, "XNOR": lambda a, b: 1 - (a ^ b)}


def test_modulus_and_parameters(orc):
    assert orc.lib().orc_modulus_Q(1024) == 134215681 == (1 << 27) - 2047  # PreviousPrime(FirstPrime(27, 2N), 2N)
    assert orc.lib().orc_modulus_Q(512) == 134215681
    p = orc.Oracle(orc.STD128_OPT, orc.GINX).p
    assert (p.n, p.N, p.q, p.qKS, p.dG, p.dKS, p.dR) == (502, 1024, 1024, 1 << 14, 4, 2, 2)


@pytest.mark.parametrize("ps", ["TOY", "STD128_OPT"])
def test_ntt_is_a_negacyclic_ring_isomorphism(orc, ps):
    o = orc.Oracle(getattr(orc, ps), orc.GINX)
    N, Q = o.p.N, o.p.Q
    rng = np.random.default_rng(1)
    a = rng.integers(0, Q, N).astype(np.uint32)
    b = rng.integers(0, Q, N).astype(np.uint32)
    assert np.array_equal(o.ntt_inv(o.ntt_fwd(a)), a)
    # X * a(X) mod X^N + 1 == rotate with sign flip
    x = np.zeros(N, dtype=np.uint32); x[1] = 1
    prod = o.ntt_inv((o.ntt_fwd(a).astype(np.uint64) * o.ntt_fwd(x) % Q).astype(np.uint32))
    ref = np.roll(a, 1).astype(np.int64); ref[0] = (Q - ref[0]) % Q
    assert np.array_equal(prod, ref.astype(np.uint32))
    # linearity
    s = ((a.astype(np.uint64) + b) % Q).astype(np.uint32)
    assert np.array_equal(o.ntt_fwd(s), ((o.ntt_fwd(a).astype(np.uint64) + o.ntt_fwd(b)) % Q).astype(np.uint32))
    # schoolbook negacyclic product on a short prefix
    aa, bb = a.astype(object), b.astype(object)
    full = o.ntt_inv((o.ntt_fwd(a).astype(np.uint64) * o.ntt_fwd(b) % Q).astype(np.uint32))
    for k in (0, 1, N - 1):
        acc = 0
        for i in range(N):
            j = (k - i) % N
            acc += (aa[i] * bb[j]) * (1 if i <= k else -1)
        assert full[k] == acc % Q


@pytest.mark.parametrize("ps", ["TOY", "STD128_OPT"])
def test_signed_digit_decompose(orc, ps):
    """a12: digits in [-B/2, B/2), digit l of component j in row j+2l, recomposition exact, centring rule t < Q>>1."""
    o = orc.Oracle(getattr(orc, ps), orc.GINX)
    N, Q, B, dG = o.p.N, o.p.Q, o.p.baseG, o.p.dG
    rng = np.random.default_rng(2)
    two = rng.integers(0, Q, 2 * N).astype(np.uint32)
    two[:6] = [0, 1, Q - 1, Q >> 1, (Q >> 1) - 1, (Q >> 1) + 1]
    d = o.decompose(two).astype(np.int64)
    signed = np.where(d > Q // 2, d - Q, d)
    assert signed.min() >= -B // 2 and signed.max() < B // 2
    for j in range(2):
        rec = sum(signed[j + 2 * l] * (B ** l) for l in range(dG)) % Q
        diff = (rec - two[j * N:(j + 1) * N].astype(np.int64)) % Q
        if B ** dG // 2 * (1 - 1 / B) > Q / 2:  # STD128_OPT: dG signed digits cover the whole centred range -> exact
            assert not diff.any()
        else:
            # TOY: 3 signed base-2^9 digits stop at 255*(1+2^9+2^18) < Q/2; like OpenFHE the top digit is simply
            # sign-truncated, i.e. the centred value wraps mod 2^27 for the ~0.2% of residues nearest +-Q/2
            assert set(np.unique(diff)) <= {0, (1 << 27) % Q, (-(1 << 27)) % Q}
            assert (diff != 0).mean() < 0.01


@pytest.mark.parametrize("ps,m", [("TOY", "GINX"), ("TOY", "AP")])
def test_truth_tables_toy(orc, ps, m):
    o = get_oracle(orc, getattr(orc, ps), getattr(orc, m))
    for name, f in TT.items():
        for a in (0, 1):
            for b in (0, 1):
                ca, cb = o.encrypt([a], seed=10 + a)[0], o.encrypt([b], seed=20 + b)[0]
                assert o.decrypt(o.eval_bingate(getattr(orc, name), ca, cb)) == f(a, b), (name, a, b)
    c1 = o.encrypt([1], seed=1)[0]
    assert o.decrypt(o.eval_not(c1)) == 0 and o.decrypt(o.bootstrap(c1)) == 1
    with pytest.raises(RuntimeError):  # EvalBinGate(ct, ct) throws in OpenFHE
        o.eval_bingate(orc.AND, c1, c1)


def test_truth_tables_std128_ginx_and_noise(orc):
    o = get_oracle(orc, orc.STD128_OPT, orc.GINX)
    q, n = o.p.q, o.p.n
    for name in ("NAND", "AND", "OR"):
        for a, b in ((0, 0), (0, 1), (1, 1)):
            ca, cb = o.encrypt([a], seed=3 + a)[0], o.encrypt([b], seed=7 + b)[0]
            out = o.eval_bingate(getattr(orc, name), ca, cb)
            assert o.decrypt(out) == TT[name](a, b)
            assert out[:n + 1].max() < q


def test_keyswitch_modswitch_pipeline_is_consistent(orc):
    """a14-a17 chained by hand equal the one-call gate."""
    o = get_oracle(orc, orc.TOY, orc.GINX)
    ca, cb = o.encrypt([1], seed=1)[0], o.encrypt([0], seed=2)[0]
    prep = o.prep(orc.NAND, ca, cb)
    acc = o.blind_rotate(orc.NAND, prep)
    out = o.keyswitch_modswitch(o.extract_modswitch(acc))
    assert np.array_equal(out, o.eval_bingate(orc.NAND, ca, cb))


@pytest.mark.parametrize("name,ps,m,nvec", [("adder_2bit", "TOY", "GINX", 10), ("adder_2bit", "TOY", "AP", 4),
                                            ("parity", "TOY", "GINX", 20), ("parity", "TOY", "AP", 6),
                                            ("adder_2bit", "STD128_OPT", "GINX", 2)])
def test_reference_configs_1_and_2_encrypted_on_oracle(bfhe, orc, name, ps, m, nvec):
    """TB_adder_2bit / TB_parity: encrypted evaluation (oracle as executor of the product's level plan) decrypts to
    the harness goldens (src/test_adder.cpp:283-295, src/test_parity.cpp:277-289,357-369)."""
    o = get_oracle(orc, getattr(orc, ps), getattr(orc, m))
    ctx = bfhe.Context(getattr(bfhe, ps), getattr(bfhe, m), device=-1)
    circ = load_circuit(bfhe, ctx, name)
    for v in VECTORS[name]["vectors"][:nvec]:
        out, _ = oracle_run_plan(circ, o, v["inputs"], seed=4)
        assert out == v["golden"], v["src"]
