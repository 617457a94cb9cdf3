"""Numpy model of the slot-sliced cluster kernel's transform network (csrc/kernels_cl.cu): checks, without a GPU, that
(two cross-block stages) o (per-CTA 256-point sub-transform: pass A, transpose, pass B, two shuffle stages) with the twiddle indices of
engine.cu's table generator equals the plain negacyclic NTT in Cooley-Tukey in-place order (slot P = evaluation at psi^(2 bitrev(P) + 1)),
and that the inverse network inverts it (unscaled by N, like the kernels)."""
import numpy as np

Q = (1 << 27) - (1 << 11) + 1
N, LOGN, R, NB = 1024, 10, 4, 256


def bitrev(x, bits):
    return int(format(x, "0%db" % bits)[::-1], 2)


def min_root():
    phi = Q - 1
    fac = []
    t = phi
    f = 2
    while f * f <= t:
        if t % f == 0:
            fac.append(f)
            while t % f == 0:
                t //= f
        f += 1
    if t > 1:
        fac.append(t)
    g = 2
    while any(pow(g, phi // p, Q) == 1 for p in fac):
        g += 1
    w = pow(g, phi // (2 * N), Q)
    best, cur, w2 = w, w, w * w % Q
    for _ in range(1, 2 * N, 2):
        best = min(best, cur)
        cur = cur * w2 % Q
    return best


PSI = min_root()
TW = [pow(PSI, bitrev(k, LOGN), Q) for k in range(N)]
ITW = [pow(pow(PSI, Q - 2, Q), bitrev(k, LOGN), Q) for k in range(N)]


def ref_fwd(a):
    a = list(a)
    t, m = N, 1
    while m < N:
        t >>= 1
        for i in range(m):
            w = TW[m + i]
            for j in range(2 * i * t, 2 * i * t + t):
                u, v = a[j], w * a[j + t] % Q
                a[j], a[j + t] = (u + v) % Q, (u - v) % Q
        m <<= 1
    return a


def ref_inv_unscaled(a):
    a = list(a)
    t, m = 1, N
    while m > 1:
        h, j1 = m >> 1, 0
        for i in range(h):
            w = ITW[h + i]
            for j in range(j1, j1 + t):
                u, v = a[j], a[j + t]
                a[j], a[j + t] = (u + v) % Q, (u - v) * w % Q
            j1 += 2 * t
        t <<= 1
        m >>= 1
    return a


def tables_fwd(k):
    T = TW
    A = [0] * 8
    A[1] = T[4 + k]
    for g in range(2):
        A[2 + g] = T[8 + 2 * k + g]
    for g in range(4):
        A[4 + g] = T[16 + 4 * k + g]
    B = np.zeros((32, 8), dtype=object)
    C = np.zeros((32, 8), dtype=object)
    for lane in range(32):
        blk = lane >> 2
        B[lane][1] = T[32 + 8 * k + blk]
        for g in range(2):
            B[lane][2 + g] = T[64 + 16 * k + 2 * blk + g]
            C[lane][2 + g] = T[256 + 64 * k + 2 * lane + g]
        for g in range(4):
            B[lane][4 + g] = T[128 + 32 * k + 4 * blk + g]
            C[lane][4 + g] = T[512 + 128 * k + 4 * lane + g]
    return A, B, C


def tables_inv(k):
    T = ITW
    A = [0] * 8
    A[1] = T[4 + k]
    for g in range(2):
        A[2 + g] = T[8 + 2 * k + g]
    for g in range(4):
        A[4 + g] = T[16 + 4 * k + g]
    S = [[T[(512 >> s) + ((256 * k + t) >> (s + 1))] for t in range(NB)] for s in range(5)]
    return A, S


def ct8(x, w, T):
    for i in range(4):
        g = i // T
        a = g * 2 * T + (i % T)
        b = a + T
        t = x[b] * w[4 // T + g] % Q
        x[a], x[b] = (x[a] + t) % Q, (x[a] - t) % Q


def gs8(x, w, T):
    for i in range(4):
        g = i // T
        a = g * 2 * T + (i % T)
        b = a + T
        x[a], x[b] = (x[a] + x[b]) % Q, (x[a] - x[b]) * w[4 // T + g] % Q


def sub_fwd(k, y):
    """y[256]: block k after the two cross-block stages.  Returns slots[t] = value at in-place position 256 k + t (the MAC order)."""
    A, B, C = tables_fwd(k)
    row = list(y)
    for lane in range(32):  # pass A
        x = [row[lane + 32 * m] for m in range(8)]
        for T in (4, 2, 1):
            ct8(x, A, T)
        for m in range(8):
            row[lane + 32 * m] = x[m]
    for lane in range(32):  # pass B
        blk, q = lane >> 2, lane & 3
        x = [row[32 * blk + q + 4 * m] for m in range(8)]
        for T in (4, 2, 1):
            ct8(x, B[lane], T)
        for m in range(8):
            row[32 * blk + q + 4 * m] = x[m]
    for lane in range(32):  # pass C
        x = [row[8 * lane + j] for j in range(8)]
        for T in (2, 1):
            ct8(x, C[lane], T)
        for j in range(8):
            row[8 * lane + j] = x[j]
    return row


def sub_inv(k, slots):
    A, S = tables_inv(k)
    v = list(slots)
    for s in range(5):  # five stages across the lanes of a warp, one value per thread t
        mask = 1 << s
        nv = [0] * NB
        for t in range(NB):
            upper = bool(t & mask)
            x, o = v[t], v[t ^ mask]
            nv[t] = (o - x) * S[s][t] % Q if upper else (x + o) % Q
        v = nv
    row = v
    for lane in range(32):  # the three widest stages, pass-A layout
        x = [row[lane + 32 * m] for m in range(8)]
        for T in (1, 2, 4):
            gs8(x, A, T)
        for m in range(8):
            row[lane + 32 * m] = x[m]
    return row


def slot_position(k, t):
    return NB * k + t


def main():
    rng = np.random.default_rng(1)
    a = [int(v) for v in rng.integers(0, Q, N)]
    ref = ref_fwd(a)
    w1, w2, w3 = TW[1], TW[2], TW[3]
    allslots = {}
    for k in range(R):
        wb = w3 if k & 2 else w2
        c1 = (Q - w1) if k & 2 else w1
        c2 = (Q - wb) if k & 1 else wb
        c3 = wb * w1 % Q
        if k in (1, 2):
            c3 = Q - c3
        y = [(a[j] + c1 * a[j + 512] + c2 * a[j + 256] + c3 * a[j + 768]) % Q for j in range(NB)]
        slots = sub_fwd(k, y)
        for t in range(NB):
            P = slot_position(k, t)
            assert slots[t] == ref[P], ("fwd", k, t)
            e = 2 * bitrev(P, LOGN) + 1
            assert slots[t] == sum(a[i] * pow(PSI, e * i, Q) for i in range(N)) % Q if t < 2 else True
        allslots[k] = slots
    print("forward network == reference NTT in in-place order: ok")
    # inverse: partial values per block, then the two cross-block stages
    part = [sub_inv(k, allslots[k]) for k in range(R)]
    iw1, iwb, iwc = ITW[1], ITW[2], ITW[3]
    out = [0] * N
    for j in range(NB):
        p0, p1, p2, p3 = (part[k][j] for k in range(R))
        u0, u1 = (p0 + p1) % Q, (p0 - p1) * iwb % Q
        u2, u3 = (p2 + p3) % Q, (p2 - p3) * iwc % Q
        out[j], out[j + 512] = (u0 + u2) % Q, (u0 - u2) * iw1 % Q
        out[j + 256], out[j + 768] = (u1 + u3) % Q, (u1 - u3) * iw1 % Q
    assert out == ref_inv_unscaled(ref) == [v * N % Q for v in a]
    print("inverse network == N * identity: ok")


if __name__ == "__main__":
    main()
