// hostmath.hpp -- host-side number theory, RNG and a plain CPU NTT.
// Used ONLY for parameter/table generation and key generation (one-off, host side by design:
// SURVEY.md 8(a) rows a1-a3).  Nothing on the gate-evaluation path runs here.
#pragma once
#include "common.hpp"
#include <cmath>
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

// counter-seeded xoshiro256**: every key row gets its own stream, so key generation is
// deterministic for a seed regardless of the OpenMP thread count.
struct Rng {
  u64 s[4];
  static u64 splitmix(u64 &x) {
    u64 z = (x += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
  }
  Rng(u64 seed, u64 stream) {
    u64 x = seed ^ (stream * 0xD1342543DE82EF95ull + 0x632BE59BD9B4E019ull);
    for (auto &v : s) v = splitmix(x);
  }
  static u64 rotl(u64 x, int k) { return (x << k) | (x >> (64 - k)); }
  u64 next() {
    u64 res = rotl(s[1] * 5, 7) * 9, t = s[1] << 17;
    s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45);
    return res;
  }
  u64 uniform(u64 m) { return next() % m; }
  int ternary() { return (int)(next() % 3) - 1; }
  i64 gauss(double sigma) {
    double u1 = ((next() >> 11) + 1.0) / 9007199254740993.0;
    double u2 = (next() >> 11) / 9007199254740992.0;
    return llround(sigma * std::sqrt(-2.0 * std::log(u1)) * std::cos(6.283185307179586 * u2));
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
