// hostmath.hpp -- host-side number theory, RNG and a plain CPU NTT.
// Used ONLY for parameter/table generation and key generation (one-off, host side by design:
// SURVEY.md 8(a) rows a1-a3).  Nothing on the gate-evaluation path runs here.
#pragma once
#include "common.hpp"
#include <cmath>
#include <cstdio>
#include <sys/random.h>
#include <vector>

namespace bfhe {

typedef unsigned __int128 u128;

inline u64 mulmod64(u64 a, u64 b, u64 m) { return (u64)((u128)a * b % m); }
inline u64 powmod64(u64 a, u64 e, u64 m) {
  u64 r = 1;
  a %= m;
  while (e) {
    if (e & 1) r = mulmod64(r, a, m);
    a = mulmod64(a, a, m);
    e >>= 1;
  }
  return r;
}
inline bool is_prime64(u64 n) {
  static const u64 bases[] = {2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37};
  if (n < 2) return false;
  for (u64 b : bases)
    if (n % b == 0) return n == b;
  u64 d = n - 1;
  int s = 0;
  while (!(d & 1)) d >>= 1, s++;
  for (u64 b : bases) {
    u64 x = powmod64(b, d, n);
    if (x == 1 || x == n - 1) continue;
    bool comp = true;
    for (int r = 1; r < s; r++) {
      x = mulmod64(x, x, n);
      if (x == n - 1) { comp = false; break; }
    }
    if (comp) return false;
  }
  return true;
}
// OpenFHE FirstPrime / PreviousPrime (nbtheory.cpp) as used by GenerateBinFHEContext
inline u64 first_prime(u32 nbits, u64 m) {
  u64 r = powmod64(2, nbits, m);
  u64 q = ((u64)1 << nbits) + (m - r) + 1;
  while (!is_prime64(q)) q += m;
  return q;
}
inline u64 previous_prime(u64 q, u64 m) {
  q -= m;
  while (!is_prime64(q)) q -= m;
  return q;
}
inline u64 min_primitive_root(u64 M, u64 Q) { // smallest primitive M-th root of unity mod Q
  u64 phi = Q - 1, t = phi;
  std::vector<u64> fac;
  for (u64 f = 2; f * f <= t; f++)
    if (t % f == 0) {
      fac.push_back(f);
      while (t % f == 0) t /= f;
    }
  if (t > 1) fac.push_back(t);
  u64 g = 2;
  for (;; g++) {
    bool ok = true;
    for (u64 f : fac)
      if (powmod64(g, phi / f, Q) == 1) { ok = false; break; }
    if (ok) break;
  }
  u64 w = powmod64(g, phi / M, Q), best = w, cur = w, w2 = mulmod64(w, w, Q);
  for (u64 k = 1; k < M; k += 2) {
    if (cur < best) best = cur;
    cur = mulmod64(cur, w2, Q);
  }
  return best;
}
inline u32 bitrev32(u32 x, int bits) {
  u32 r = 0;
  for (int i = 0; i < bits; i++) r |= ((x >> i) & 1) << (bits - 1 - i);
  return r;
}
inline int ilog2_ceil(u64 x) {
  int l = 0;
  while (((u64)1 << l) < x) l++;
  return l;
}

// ---- randomness ------------------------------------------------------------------------------------------------------
// Secret keys, RGSW / key-switching keys and fresh-encryption masks and noise come from ChaCha20 (RFC 8439 block function)
// keyed with a 256-bit SeedKey; every key row / ciphertext gets its own 64-bit stream id (the ChaCha nonce), so generation is
// deterministic for a SeedKey regardless of the OpenMP thread count.
//   seed == 0  -> the SeedKey is 256 bits of OS entropy (getrandom / /dev/urandom): the default, and the only secure choice
//   seed != 0  -> the SeedKey is derived from the 64-bit seed alone: reproducible keys and ciphertexts for tests, oracle parity
//                 and benchmarks; NOT confidential (2^64 search space, and the seeds in the tests are public)
// (the reference draws from OpenFHE's PRNG, which seeds itself from the OS; include/bfhe.h states this contract)
struct SeedKey {
  u32 k[8];
  static bool os_entropy(void *buf, size_t len) {
    u8 *p = (u8 *)buf;
    size_t got = 0;
    while (got < len) {
      ssize_t r = getrandom(p + got, len - got, 0);
      if (r <= 0) break;
      got += (size_t)r;
    }
    if (got == len) return true;
    FILE *f = std::fopen("/dev/urandom", "rb");
    if (!f) return false;
    const size_t rd = std::fread(p, 1, len, f);
    std::fclose(f);
    return rd == len;
  }
  static bool make(u64 seed, SeedKey &out) {
    if (seed == 0) return os_entropy(out.k, sizeof out.k);
    u64 x = seed; // splitmix64 expansion of the test seed into the key words (domain separation only, not security)
    for (int i = 0; i < 4; i++) {
      u64 z = (x += 0x9E3779B97F4A7C15ull);
      z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
      z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
      z ^= z >> 31;
      out.k[2 * i] = (u32)z; out.k[2 * i + 1] = (u32)(z >> 32);
    }
    return true;
  }
};

struct Rng { // ChaCha20 keystream: key = SeedKey, nonce = stream id, 64-bit block counter
  u32 st[16], buf[16];
  int pos = 16;
  Rng(const SeedKey &key, u64 stream) {
    st[0] = 0x61707865u; st[1] = 0x3320646eu; st[2] = 0x79622d32u; st[3] = 0x6b206574u;
    for (int i = 0; i < 8; i++) st[4 + i] = key.k[i];
    st[12] = 0; st[13] = 0;
    st[14] = (u32)stream; st[15] = (u32)(stream >> 32);
  }
  static u32 rotl(u32 x, int k) { return (x << k) | (x >> (32 - k)); }
  void refill() {
    u32 x[16];
    for (int i = 0; i < 16; i++) x[i] = st[i];
#define BFHE_QR(a, b, c, d)                                                                                            \
  x[a] += x[b]; x[d] = rotl(x[d] ^ x[a], 16); x[c] += x[d]; x[b] = rotl(x[b] ^ x[c], 12);                              \
  x[a] += x[b]; x[d] = rotl(x[d] ^ x[a], 8);  x[c] += x[d]; x[b] = rotl(x[b] ^ x[c], 7);
    for (int r = 0; r < 10; r++) {
      BFHE_QR(0, 4, 8, 12) BFHE_QR(1, 5, 9, 13) BFHE_QR(2, 6, 10, 14) BFHE_QR(3, 7, 11, 15)
      BFHE_QR(0, 5, 10, 15) BFHE_QR(1, 6, 11, 12) BFHE_QR(2, 7, 8, 13) BFHE_QR(3, 4, 9, 14)
    }
#undef BFHE_QR
    for (int i = 0; i < 16; i++) buf[i] = x[i] + st[i];
    if (++st[12] == 0) ++st[13];
    pos = 0;
  }
  u32 next32() {
    if (pos >= 16) refill();
    return buf[pos++];
  }
  u64 next() { const u64 lo = next32(); return lo | ((u64)next32() << 32); }
  u64 uniform(u64 m) { // unbiased: rejection sampling on the smallest covering power of two
    if (m <= 1) return 0;
    const int bits = ilog2_ceil(m);
    const u64 mask = bits >= 64 ? ~0ull : (((u64)1 << bits) - 1);
    for (;;) {
      const u64 v = (bits <= 32 ? (u64)next32() : next()) & mask;
      if (v < m) return v;
    }
  }
  int ternary() { return (int)uniform(3) - 1; }
};

// discrete Gaussian by inversion of a cumulative table (what OpenFHE's DiscreteGaussianGenerator does for small sigma):
// P(|x| = k) proportional to exp(-k^2 / (2 sigma^2)), tail cut at 12 sigma (mass < 2^-100), 64-bit uniform draw
struct GaussTable {
  std::vector<u64> cdf; // cdf[k] = floor(2^64 * P(|x| <= k)) for k >= 0, saturating
  explicit GaussTable(double sigma) {
    const int kmax = (int)std::ceil(12 * sigma);
    std::vector<long double> w(kmax + 1);
    long double tot = 0;
    for (int k = 0; k <= kmax; k++) { w[k] = std::exp(-(long double)k * k / (2.0L * sigma * sigma)) * (k ? 2.0L : 1.0L); tot += w[k]; }
    long double acc = 0;
    cdf.resize(kmax + 1);
    for (int k = 0; k <= kmax; k++) {
      acc += w[k] / tot;
      const long double v = acc * 18446744073709551616.0L;
      cdf[k] = v >= 18446744073709551615.0L ? ~0ull : (u64)v;
    }
    cdf[kmax] = ~0ull;
  }
  i64 sample(Rng &r) const {
    const u64 u = r.next();
    size_t k = 0;
    while (k + 1 < cdf.size() && u > cdf[k]) k++;
    if (k == 0) return 0;
    return (r.next32() & 1) ? (i64)k : -(i64)k;
  }
};

// plain negacyclic NTT over Z_Q[X]/(X^N+1) for key generation (a*z products)
struct HostNtt {
  u32 N = 0, Q = 0;
  int logN = 0;
  std::vector<u32> tw, itw; // psi^bitrev(k), psi^-bitrev(k)
  u32 ninv = 0;
  void init(u32 N_, u32 Q_, u64 psi) {
    N = N_; Q = Q_; logN = ilog2_ceil(N);
    tw.resize(N); itw.resize(N);
    u64 psi_inv = powmod64(psi, Q - 2, Q);
    for (u32 k = 0; k < N; k++) {
      tw[k] = (u32)powmod64(psi, bitrev32(k, logN), Q);
      itw[k] = (u32)powmod64(psi_inv, bitrev32(k, logN), Q);
    }
    ninv = (u32)powmod64(N, Q - 2, Q);
  }
  void fwd(u32 *a) const {
    u32 t = N;
    for (u32 m = 1; m < N; m <<= 1) {
      t >>= 1;
      for (u32 i = 0; i < m; i++) {
        const u64 w = tw[m + i];
        const u32 j1 = 2 * i * t;
        for (u32 j = j1; j < j1 + t; j++) {
          u32 u = a[j], v = (u32)(w * a[j + t] % Q);
          u32 s = u + v;
          a[j] = s >= Q ? s - Q : s;
          a[j + t] = u >= v ? u - v : u + Q - v;
        }
      }
    }
  }
  void inv(u32 *a) const {
    u32 t = 1;
    for (u32 m = N; m > 1; m >>= 1) {
      u32 h = m >> 1, j1 = 0;
      for (u32 i = 0; i < h; i++) {
        const u64 w = itw[h + i];
        for (u32 j = j1; j < j1 + t; j++) {
          u32 u = a[j], v = a[j + t];
          u32 s = u + v;
          a[j] = s >= Q ? s - Q : s;
          u32 d = u >= v ? u - v : u + Q - v;
          a[j + t] = (u32)(w * d % Q);
        }
        j1 += 2 * t;
      }
      t <<= 1;
    }
    for (u32 j = 0; j < N; j++) a[j] = (u32)((u64)a[j] * ninv % Q);
  }
};

} // namespace bfhe
