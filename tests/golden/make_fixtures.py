"""Generate the committed golden fixtures from the reference's own data (run in the build container, where
/root/reference exists; the GPU box only ever sees the outputs).

  tests/golden/circuits/<name>.npz   netlists of the reference's example circuits as integer arrays
                                     (parsed by the product's O(G) front-end; no reference source text is copied)
  tests/golden/vectors.json          input / expected-output bit vectors exactly as the reference's harnesses build
                                     them: seeded glibc rand() vectors (src/test_adder.cpp:180-217,
                                     src/test_parity.cpp:176-206, src/test_multiplier.cpp:183-224,
                                     src/test_comparator.cpp:184-269) and the hard-coded known-answer vectors
                                     (src/test_aes.cpp:185-229, src/test_md5.cpp:203-228, src/test_sha256.cpp:204-239)
"""
import ctypes
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
import bfhe_loader  # noqa: E402

B = bfhe_loader.load_package()
REF = "/root/reference/examples"
libc = ctypes.CDLL("libc.so.6")


def hex2bits(h):  # HexStr2UintVec, src/utils.cpp:49-71: last hex digit first, each nibble LSB first
    out = []
    for ch in reversed(h):
        n = int(ch, 16)
        out += [(n >> b) & 1 for b in range(4)]
    return out


def bin2bits(s):  # BinStr2UintVec, src/utils.cpp:73-89
    return [int(ch) for ch in reversed(s)]


def rand_pairs(seed, nbits):  # srand(test_ix); in1[i] = rand()%2; in2[i] = rand()%2 interleaved
    libc.srand(seed)
    a, b = [], []
    for _ in range(nbits):
        a.append(libc.rand() % 2)
        b.append(libc.rand() % 2)
    return a, b


def to_int(bits):
    return sum(int(b) << i for i, b in enumerate(bits))


CIRCUITS = {
    # name: (path, kind)  kind: out | old | new
    "adder_2bit": ("simple_ckts/adder_2bit/adder_2bit.out", "out"),
    "parity": ("simple_ckts/parity/parity.out", "out"),
    "adder_32bit": ("old_bristol_ckts/arith/adder_32bit.txt", "old"),
    "comparator_32bit_signed_lt": ("old_bristol_ckts/arith/comparator_32bit_signed_lt.txt", "old"),
    "comparator_32bit_signed_lteq": ("old_bristol_ckts/arith/comparator_32bit_signed_lteq.txt", "old"),
    "comparator_32bit_unsigned_lt": ("old_bristol_ckts/arith/comparator_32bit_unsigned_lt.txt", "old"),
    "comparator_32bit_unsigned_lteq": ("old_bristol_ckts/arith/comparator_32bit_unsigned_lteq.txt", "old"),
    "mult_32x32": ("old_bristol_ckts/arith/mult_32x32.txt", "old"),
    "AES-non-expanded": ("old_bristol_ckts/crypto/AES-non-expanded.txt", "old"),
    "AES-expanded": ("old_bristol_ckts/crypto/AES-expanded.txt", "old"),
    "md5": ("old_bristol_ckts/crypto/md5.txt", "old"),
    "sha256": ("new_bristol_ckts/crypto/sha256.txt", "new"),  # old sha-256.txt is absent upstream (.MISSING_LARGE_BLOBS:5)
}

MSGS = ["00" * 64,
        "000102030405060708090a0b0c0d0e0f101112131415161718191a1b1c1d1e1f202122232425262728292a2b2c2d2e2f"
        "303132333435363738393a3b3c3d3e3f",
        "ff" * 64,
        "243f6a8885a308d313198a2e03707344a4093822299f31d0082efa98ec4e6c89452821e638d01377be5466cf34e90c6c"
        "c0ac29b7c97c50dd3f84d5b5b5470917"]
SHA256 = ["da5698be17b9b46962335799779fbeca8ce5d491c0d26243bafef9ea1837a9d8",
          "fc99a2df88f42a7a7bb9d18033cdc6a20256755f9d5b9a5044a9cc315abe84a7",
          "ef0c748df4da50a8d6c43c013edc3ce76c9d9fa9a1458ade56eb86c0a64492d2",
          "cf0ae4eb67d38ffeb94068984b22abde4e92bc548d14585e48dca8882d7b09ce"]
MD5 = ["ac1d1f03d08ea56eb767ab1f91773174", "cad94491c9e401d9385bfc721ef55f62", "b487195651913e494b55c6bddf405c01",
       "3715f568f422db75cc8d65e11764ff01"]
SHA256_IV = "6a09e667bb67ae853c6ef372a54ff53a510e527f9b05688c1f83d9ab5be0cd19"
AES = {
    "AES-non-expanded": [
        ("0" * 32, "0" * 32, "01110100110101000010110001010011100110100101111100110010000100011101110000110100"
                             "010100011111011100101011110100101001011101100110"),
        ("f" * 32, "f" * 32, "10011110100111010101110010011000010010100000111010001010010011010000110011110011"
                             "000000010100110100111110100001001111110100111101")],
    "AES-expanded": [
        ("0" * 32, "0" * 352, "0110110001101100011011000110110001101100011011000110110001101100011011000110110001101100"
                              "0110110001101100011011000110110001101100"),
        ("f" * 32, "f" * 352, "0011001000110010001100100011001000110010001100100011001000110010001100100011001000110010"
                              "0011001000110010001100100011001000110010")],
}


def vectors(name, info):
    vs = []
    if name in ("adder_2bit", "adder_32bit"):
        n = info["input_bits"][0]
        for t in range(10):
            a, b = rand_pairs(t, n)
            s = to_int(a) + to_int(b)
            vs.append(dict(inputs=[a, b], golden=[(s >> i) & 1 for i in range(n + 1)], src="src/test_adder.cpp:180-217 seed %d" % t))
    elif name == "parity":
        for t in range(10):  # two chained runs per vector (src/test_parity.cpp:293-297)
            libc.srand(t)
            bits = [libc.rand() % 2 for _ in range(8)] + [0]
            odd = bin(to_int(bits)).count("1") & 1
            even = 1 - odd
            vs.append(dict(inputs=[bits], golden=[even, odd], src="src/test_parity.cpp:176-206 seed %d generate" % t))
            vs.append(dict(inputs=[bits[:8] + [even]], golden=[0, 1], src="src/test_parity.cpp:293-297 seed %d check" % t))
    elif name.startswith("comparator"):
        for t in range(10):
            a, b = rand_pairs(t, 32)
            if t == 0:
                b = list(a)
            ia, ib = to_int(a), to_int(b)
            if "unsigned" not in name:
                ia = ia - (1 << 32) if ia >> 31 else ia
                ib = ib - (1 << 32) if ib >> 31 else ib
            out = int(ib >= ia) if "lteq" in name else int(ib > ia)
            vs.append(dict(inputs=[a, b], golden=[out], src="src/test_comparator.cpp:184-269 seed %d" % t))
    elif name == "mult_32x32":
        for t in range(10):
            a, b = rand_pairs(t, 32)
            c = to_int(a) * to_int(b)
            vs.append(dict(inputs=[a, b], golden=[(c >> i) & 1 for i in range(64)], src="src/test_multiplier.cpp:183-224 seed %d" % t))
    elif name in AES:
        for i, (h1, h2, ob) in enumerate(AES[name]):  # no reversal (src/test_aes.cpp:263-267 is commented out)
            vs.append(dict(inputs=[hex2bits(h1), hex2bits(h2)], golden=bin2bits(ob), src="src/test_aes.cpp:185-229 subtest %d" % i))
    elif name == "md5":
        for i, (m, d) in enumerate(zip(MSGS, MD5)):  # std::reverse of input and golden (src/test_md5.cpp:253-254)
            vs.append(dict(inputs=[hex2bits(m)[::-1]], golden=hex2bits(d)[::-1], src="src/test_md5.cpp:203-228 subtest %d" % i))
    elif name == "sha256":
        # new-format circuit = compression function with the chaining value as input 2; no reversal
        for i, (m, d) in enumerate(zip(MSGS, SHA256)):
            vs.append(dict(inputs=[hex2bits(m), hex2bits(SHA256_IV)], golden=hex2bits(d), src="src/test_sha256.cpp:204-239 subtest %d" % i))
    return vs


def main():
    ctx = B.Context(B.TOY, B.GINX, device=-1)
    allv = {}
    for name, (rel, kind) in CIRCUITS.items():
        c = B.Circuit(ctx)
        path = os.path.join(REF, rel)
        if kind == "out":
            c.ReadFile(path)
        else:
            c.ReadBristol(path, new_format=(kind == "new"))
        nl = c.get_netlist()
        info = c.info()
        np.savez_compressed(os.path.join(HERE, "circuits", name + ".npz"), **nl)
        vs = vectors(name, info)
        for v in vs:  # the plaintext circuit must reproduce every golden vector before it is committed
            c.Reset()
            c.setPlaintext(True)
            c.SetInput(v["inputs"])
            out = c.Clock()[0]
            assert out == v["golden"], (name, v["src"])
        allv[name] = dict(info=info, gate_count=c.dumpGateCount(), source=rel, vectors=vs)
        print(name, info, len(vs), "vectors ok")
    json.dump(allv, open(os.path.join(HERE, "vectors.json"), "w"))


if __name__ == "__main__":
    main()
