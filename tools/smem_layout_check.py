"""Checks the shared-memory row layout of kernels_v2.cu: bijection, bank-conflict freedom of the three tile shapes, and the
closed-form address expressions the kernel uses (base XOR compile-time constant) against phys()."""


def phys(p):
    T3, c = p >> 4, (p >> 2) & 3
    return 16 * (T3 ^ ((T3 >> 3) & 1)) + 4 * (c ^ ((T3 >> 1) & 3)) + (p & 3)


def wavefronts(addrs, width):
    grp = {1: 32, 2: 16, 4: 8}[width]
    tot = 0
    for g0 in range(0, len(addrs), grp):
        banks = {}
        for a in addrs[g0:g0 + grp]:
            for wd in range(width):
                banks.setdefault((a + wd) % 32, set()).add(a + wd)
        tot += max(len(s) for s in banks.values())
    return tot


def main():
    assert sorted(phys(p) for p in range(1024)) == list(range(1024))
    for it in range(2):
        for lane in range(32):
            T = 32 * it + lane
            # column shape: positions T + 64k
            base1 = 16 * (T >> 4) + 4 * (((T >> 2) & 3) ^ (T >> 5)) + (T & 3)
            for k in range(16):
                i = ((k >> 1) & 1) * 2 + (k & 1)
                assert (base1 ^ (16 * (i >> 1) + 8 * (i & 1))) + 64 * k == phys(T + 64 * k)
            # pair shape: positions 64u + 8r + 2w (+1)
            u, w = T >> 2, T & 3
            base2 = 64 * u + 16 * ((u >> 1) & 1) + 8 * (u & 1) + 4 * (w >> 1) + 2 * (w & 1)
            for r in range(8):
                K = 32 * (r >> 2) + 16 * ((r >> 1) & 1) + 8 * (r & 1) + 4 * (r >> 2)
                assert base2 ^ K == phys(64 * u + 8 * r + 2 * w), (T, r)
                assert phys(64 * u + 8 * r + 2 * w + 1) == (base2 ^ K) + 1
            # row shape: positions 16*T + 4c
            base3 = 16 * (T ^ ((T >> 3) & 1)) + 4 * ((T >> 1) & 3)
            for c in range(4):
                assert base3 ^ (4 * c) == phys(16 * T + 4 * c)
    for it in range(2):
        for k in range(16):
            assert wavefronts([phys(64 * k + 32 * it + l) for l in range(32)], 1) == 1
        for r in range(8):
            assert wavefronts([phys(64 * ((32 * it + l) >> 2) + 8 * r + 2 * (l & 3)) for l in range(32)], 2) == 2
        for c in range(4):
            assert wavefronts([phys(16 * (32 * it + l) + 4 * c) for l in range(32)], 4) == 4
    print("layout ok: bijective, conflict-free in all three shapes, closed-form addresses agree")


if __name__ == "__main__":
    main()
