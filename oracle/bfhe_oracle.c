/*
 * bfhe_oracle.c -- CPU ORACLE (TEST INFRASTRUCTURE ONLY).  See bfhe_oracle.h.
 *
 * Restates OpenFHE binfhe 1.0.x (not vendored in /root/reference; algorithm spec in
 * SURVEY.md App. C).  PARITY UNPINNED at ciphertext level (no fixtures exist upstream).
 * Every function names the OpenFHE routine it restates and the reference call site
 * that reaches it.  64-bit arithmetic, evaluation-form accumulator, textbook loops:
 * deliberately a different formulation from the CUDA path so that agreement between
 * the two is evidence, not tautology.
 */
#include "bfhe_oracle.h"
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef uint64_t u64;
typedef uint32_t u32;
typedef int64_t i64;
typedef unsigned __int128 u128;

/* ---- key blob format (documented in include/bfhe.h; re-declared here on purpose) ---- */
typedef struct {
  char magic[8]; /* "BFHEKEY1" */
  u32 version, paramset, method, n, N, q;
  u32 baseKS, dKS, baseG, dG, baseR, dR;
  u32 has_sk, ksk_elem_bytes;
  u64 Q, qKS, bk_words, ksk_elems;
  u64 reserved[4];
} keyblob_header;

struct orc_ctx {
  orc_params p;
  u32 logBG, factor; /* factor = 2N/q */
  u64 psi, psi_inv, n_inv;
  u32 *tw, *tw_sh, *itw, *itw_sh; /* psi powers in bit-reversed order + Shoup companions */
  int32_t *sk;                    /* n, values in {-1,0,1} */
  int32_t *z;                     /* N, RLWE key (only after orc_keygen) */
  u32 *bk_coef;                   /* canonical coefficient-form bootstrapping key */
  u32 *bk_eval;                   /* same, evaluation form (oracle's own slot order) */
  u64 bk_words;
  void *ksk;                      /* [N][baseKS][dKS][n+1], u16 or u32 */
  u32 ksk_elem_bytes;
  u64 ksk_elems;
  u32 *mono;                      /* GINX: NTT(X^m - 1), m in [0,2N)  (RingGSWCryptoParams::m_monomials) */
  u32 gate_const[6];
  u64 Gpow[8];
  u64 mu; /* floor(2^64 / Q) for Barrett reduction of 64-bit sums */
  int has_keys;
};

/* ---------------- number theory (OpenFHE core/math/nbtheory.cpp) ---------------- */
static u64 mulmod(u64 a, u64 b, u64 m) { return (u64)((u128)a * b % m); }
static u64 powmod(u64 a, u64 e, u64 m) {
  u64 r = 1;
  a %= m;
  while (e) {
    if (e & 1) r = mulmod(r, a, m);
    a = mulmod(a, a, m);
    e >>= 1;
  }
  return r;
}
static int is_prime(u64 n) {
  static const u64 bases[] = {2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37};
  if (n < 2) return 0;
  for (int i = 0; i < 12; i++)
    if (n % bases[i] == 0) return n == bases[i];
  u64 d = n - 1;
  int s = 0;
  while (!(d & 1)) d >>= 1, s++;
  for (int i = 0; i < 12; i++) {
    u64 x = powmod(bases[i], d, n);
    if (x == 1 || x == n - 1) continue;
    int comp = 1;
    for (int r = 1; r < s; r++) {
      x = mulmod(x, x, n);
      if (x == n - 1) { comp = 0; break; }
    }
    if (comp) return 0;
  }
  return 1;
}
/* FirstPrime(nBits, m): first prime = 1 mod m at or above 2^nBits + m + 1 (r = 2^nBits mod m = 0 here) */
static u64 first_prime(u32 nbits, u64 m) {
  u64 r = powmod(2, nbits, m);
  u64 q = ((u64)1 << nbits) + (m - r) + 1;
  while (!is_prime(q)) q += m;
  return q;
}
static u64 previous_prime(u64 q, u64 m) {
  q -= m;
  while (!is_prime(q)) q -= m;
  return q;
}
/* GenerateBinFHEContext: Q = PreviousPrime(FirstPrime(27, 2N), 2N)  (binfhecontext.cpp, reached from src/circuit.cpp:88) */
uint64_t orc_modulus_Q(int N) { return previous_prime(first_prime(27, 2 * (u64)N), 2 * (u64)N); }

/* smallest primitive 2N-th root of unity mod Q (any primitive root gives identical ciphertexts, App. C.6) */
static u64 min_root_of_unity(u64 M, u64 Q) {
  u64 phi = Q - 1, fac[16];
  int nf = 0;
  u64 t = phi;
  for (u64 f = 2; f * f <= t; f++)
    if (t % f == 0) {
      fac[nf++] = f;
      while (t % f == 0) t /= f;
    }
  if (t > 1) fac[nf++] = t;
  u64 g = 2;
  for (;; g++) {
    int ok = 1;
    for (int i = 0; i < nf; i++)
      if (powmod(g, phi / fac[i], Q) == 1) { ok = 0; break; }
    if (ok) break;
  }
  u64 w = powmod(g, phi / M, Q), best = w, cur = w, w2 = mulmod(w, w, Q);
  for (u64 k = 1; k < M; k += 2) { /* odd powers are exactly the primitive M-th roots */
    if (cur < best) best = cur;
    cur = mulmod(cur, w2, Q);
  }
  return best;
}
static u32 bitrev(u32 x, int bits) {
  u32 r = 0;
  for (int i = 0; i < bits; i++) r |= ((x >> i) & 1) << (bits - 1 - i);
  return r;
}
static int ilog2(u64 x) {
  int l = 0;
  while (((u64)1 << l) < x) l++;
  return l;
}

/* ---------------- RNG (not parity relevant; OpenFHE's PRNG is unseeded) ---------------- */
typedef struct { u64 s[4]; } rng_t;
static u64 splitmix(u64 *x) {
  u64 z = (*x += 0x9E3779B97F4A7C15ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
static void rng_seed(rng_t *r, u64 seed) {
  for (int i = 0; i < 4; i++) r->s[i] = splitmix(&seed);
}
static u64 rotl(u64 x, int k) { return (x << k) | (x >> (64 - k)); }
static u64 rng_next(rng_t *r) {
  u64 *s = r->s, res = rotl(s[1] * 5, 7) * 9, t = s[1] << 17;
  s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45);
  return res;
}
static u64 rng_uniform(rng_t *r, u64 m) { return rng_next(r) % m; }
static int rng_ternary(rng_t *r) { return (int)(rng_next(r) % 3) - 1; }
static i64 rng_gauss(rng_t *r, double sigma) {
  double u1 = ((rng_next(r) >> 11) + 1.0) / 9007199254740993.0;
  double u2 = (rng_next(r) >> 11) / 9007199254740992.0;
  return llround(sigma * sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * u2));
}
#define SIGMA 3.19

/* ---------------- context (BinFHEContext::GenerateBinFHEContext, src/circuit.cpp:88) ---------------- */
static u32 shoup(u64 w, u64 Q) { return (u32)((w << 32) / Q); }

orc_ctx *orc_create(int paramset, int method) {
  orc_ctx *c = (orc_ctx *)calloc(1, sizeof(*c));
  orc_params *p = &c->p;
  p->paramset = paramset;
  p->method = method;
  if (paramset == ORC_TOY) { /* LWECryptoParams(64, 512, 512, Q, Q, 3.19, 25); RingGSW(1<<9, 23) */
    p->n = 64; p->N = 512; p->q = 512; p->Q = orc_modulus_Q(512); p->qKS = p->Q;
    p->baseKS = 25; p->baseG = 1u << 9; p->baseR = 23;
  } else if (paramset == ORC_STD128_OPT) { /* LWECryptoParams(502, 1024, 1024, Q, 1<<14, 3.19, 1<<7); RingGSW(1<<7, 32) */
    p->n = 502; p->N = 1024; p->q = 1024; p->Q = orc_modulus_Q(1024); p->qKS = 1u << 14;
    p->baseKS = 1u << 7; p->baseG = 1u << 7; p->baseR = 32;
  } else { free(c); return NULL; }
  if (method != ORC_AP && method != ORC_GINX) { free(c); return NULL; }
  p->dKS = (u32)ceil(log((double)p->qKS) / log((double)p->baseKS));
  p->dG = (u32)ceil(log((double)p->Q) / log((double)p->baseG));
  p->dR = (u32)ceil(log((double)p->q) / log((double)p->baseR));
  p->ct_words = p->n + 1;
  p->ct_stride = (p->ct_words + 3) & ~3u;
  c->logBG = ilog2(p->baseG);
  c->factor = 2 * p->N / p->q;
  u32 N = p->N;
  u64 Q = p->Q;
  int logN = ilog2(N);
  c->psi = min_root_of_unity(2 * (u64)N, Q);
  c->psi_inv = powmod(c->psi, Q - 2, Q);
  c->n_inv = powmod(N, Q - 2, Q);
  c->mu = (u64)((((u128)1) << 64) / Q);
  c->tw = malloc(4 * N); c->tw_sh = malloc(4 * N); c->itw = malloc(4 * N); c->itw_sh = malloc(4 * N);
  for (u32 k = 0; k < N; k++) {
    u64 w = powmod(c->psi, bitrev(k, logN), Q), wi = powmod(c->psi_inv, bitrev(k, logN), Q);
    c->tw[k] = (u32)w; c->tw_sh[k] = shoup(w, Q);
    c->itw[k] = (u32)wi; c->itw_sh[k] = shoup(wi, Q);
  }
  /* gate constants (RingGSWCryptoParams::m_gateConst): OR 5q/8, AND 7q/8, NOR q/8, NAND 3q/8, XOR_FAST 5q/8, XNOR_FAST q/8 */
  u32 q = p->q;
  c->gate_const[ORC_OR] = 5 * (q >> 3); c->gate_const[ORC_AND] = 7 * (q >> 3);
  c->gate_const[ORC_NOR] = 1 * (q >> 3); c->gate_const[ORC_NAND] = 3 * (q >> 3);
  c->gate_const[ORC_XOR_FAST] = 5 * (q >> 3); c->gate_const[ORC_XNOR_FAST] = 1 * (q >> 3);
  u64 g = 1;
  for (u32 i = 0; i < p->dG; i++) { c->Gpow[i] = g; g *= p->baseG; }
  c->bk_words = (method == ORC_GINX) ? (u64)p->n * 2 * (2 * p->dG) * 2 * N
                                     : (u64)p->n * (p->baseR - 1) * p->dR * (2 * p->dG) * 2 * N;
  c->ksk_elem_bytes = (p->qKS <= 65536) ? 2 : 4;
  c->ksk_elems = (u64)N * p->baseKS * p->dKS * (p->n + 1);
  if (method == ORC_GINX) { /* monomials X^m - 1 in evaluation form */
    c->mono = malloc((size_t)2 * N * N * 4);
    for (u32 m = 0; m < 2 * N; m++) {
      u32 *poly = c->mono + (size_t)m * N;
      memset(poly, 0, 4 * N);
      poly[0] = (u32)(Q - 1);
      if (m < N) poly[m] = (u32)((poly[m] + 1) % Q);
      else poly[m - N] = (u32)((poly[m - N] + Q - 1) % Q);
      orc_ntt_fwd(c, poly);
    }
  }
  return c;
}
void orc_destroy(orc_ctx *c) {
  if (!c) return;
  free(c->tw); free(c->tw_sh); free(c->itw); free(c->itw_sh); free(c->sk); free(c->z);
  free(c->bk_coef); free(c->bk_eval); free(c->ksk); free(c->mono); free(c);
}
void orc_get_params(const orc_ctx *c, orc_params *out) { *out = c->p; }

/* ---------------- negacyclic NTT (NativePoly::SetFormat; transformnat-impl.h) ---------------- */
static inline u32 mul_shoup(u32 x, u32 w, u32 wsh, u32 Q) {
  u32 hi = (u32)(((u64)x * wsh) >> 32);
  u32 r = x * w - hi * Q;
  return r - (Q & -(u32)(r >= Q));
}
/* Cooley-Tukey, natural in, bit-reversed out: slot j = a(psi^(2*br(j)+1)) */
void orc_ntt_fwd(const orc_ctx *c, u32 *restrict a) {
  const u32 N = c->p.N, Q = (u32)c->p.Q;
  const u32 *restrict tw = c->tw, *restrict tws = c->tw_sh;
  u32 t = N;
  for (u32 m = 1; m < N; m <<= 1) {
    t >>= 1;
    for (u32 i = 0; i < m; i++) {
      const u32 w = tw[m + i], ws = tws[m + i];
      u32 *restrict lo = a + 2 * i * t, *restrict hi = lo + t;
#pragma omp simd
      for (u32 j = 0; j < t; j++) {
        const u32 u = lo[j], v = mul_shoup(hi[j], w, ws, Q);
        const u32 s = u + v;
        lo[j] = s - (Q & -(u32)(s >= Q));
        hi[j] = u - v + (Q & -(u32)(u < v));
      }
    }
  }
}
/* Gentleman-Sande, bit-reversed in, natural out, times N^-1 */
void orc_ntt_inv(const orc_ctx *c, u32 *restrict a) {
  const u32 N = c->p.N, Q = (u32)c->p.Q;
  const u32 *restrict tw = c->itw, *restrict tws = c->itw_sh;
  u32 t = 1;
  for (u32 m = N; m > 1; m >>= 1) {
    const u32 h = m >> 1;
    for (u32 i = 0; i < h; i++) {
      const u32 w = tw[h + i], ws = tws[h + i];
      u32 *restrict lo = a + 2 * i * t, *restrict hi = lo + t;
#pragma omp simd
      for (u32 j = 0; j < t; j++) {
        const u32 u = lo[j], v = hi[j];
        const u32 s = u + v;
        lo[j] = s - (Q & -(u32)(s >= Q));
        const u32 d = u - v + (Q & -(u32)(u < v));
        hi[j] = mul_shoup(d, w, ws, Q);
      }
    }
    t <<= 1;
  }
  const u32 ninv = (u32)c->n_inv, nsh = shoup(c->n_inv, Q);
#pragma omp simd
  for (u32 j = 0; j < N; j++) a[j] = mul_shoup(a[j], ninv, nsh, Q);
}
static inline u32 mulmod32(u32 a, u32 b, u32 Q) { return (u32)((u64)a * b % Q); }
/* s mod Q for any 64-bit s (Barrett with mu = floor(2^64/Q)) */
static inline u32 barrett64(u64 s, u64 Q, u64 mu) {
  u64 qh = (u64)(((u128)s * mu) >> 64);
  u64 r = s - qh * Q;
  while (r >= Q) r -= Q;
  return (u32)r;
}

/* ---------------- key generation (BinFHEContext::KeyGen / BTKeyGen, src/circuit.cpp:90-91) ---------------- */
/* RGSW encryption of (sign * X^mm) under z; rows (a_r, a_r*z + e_r) + m*G, COEFFICIENT form out
 * (RingGSWAccumulatorCGGI::KeyGenCGGI / RingGSWAccumulatorDM::KeyGenDM). msg_present = 0 encrypts 0. */
static void rgsw_encrypt(const orc_ctx *c, const u32 *z_eval, int msg_present, u32 mm, int sign, rng_t *r, u32 *out) {
  u32 N = c->p.N, rows = 2 * c->p.dG;
  u32 Q = (u32)c->p.Q;
  u32 *tmp = malloc(4 * N);
  for (u32 row = 0; row < rows; row++) {
    u32 *a = out + ((size_t)row * 2 + 0) * N, *b = out + ((size_t)row * 2 + 1) * N;
    for (u32 j = 0; j < N; j++) a[j] = (u32)rng_uniform(r, Q);
    memcpy(tmp, a, 4 * N);
    orc_ntt_fwd(c, tmp);
    for (u32 j = 0; j < N; j++) tmp[j] = mulmod32(tmp[j], z_eval[j], Q);
    orc_ntt_inv(c, tmp);
    for (u32 j = 0; j < N; j++) {
      i64 e = rng_gauss(r, SIGMA);
      b[j] = (u32)(((i64)tmp[j] + e + Q) % Q);
    }
    if (msg_present) {
      u64 g = c->Gpow[row >> 1] % Q;
      u32 *dst = (row & 1) ? b : a; /* row 2i: component 0; row 2i+1: component 1 */
      dst[mm] = (u32)(sign > 0 ? (dst[mm] + g) % Q : (dst[mm] + Q - g) % Q);
    }
  }
  free(tmp);
}

static void ksk_store(const orc_ctx *c, u64 idx, u32 v) {
  if (c->ksk_elem_bytes == 2) ((uint16_t *)c->ksk)[idx] = (uint16_t)v;
  else ((u32 *)c->ksk)[idx] = v;
}
static inline u32 ksk_load(const orc_ctx *c, u64 idx) {
  return c->ksk_elem_bytes == 2 ? ((const uint16_t *)c->ksk)[idx] : ((const u32 *)c->ksk)[idx];
}

static void build_bk_eval(orc_ctx *c) {
  free(c->bk_eval);
  c->bk_eval = malloc(c->bk_words * 4);
  u32 N = c->p.N;
  u64 npoly = c->bk_words / N;
#pragma omp parallel for schedule(static)
  for (i64 k = 0; k < (i64)npoly; k++) {
    memcpy(c->bk_eval + k * N, c->bk_coef + k * N, 4 * N);
    orc_ntt_fwd(c, c->bk_eval + k * N);
  }
}

int orc_keygen(orc_ctx *c, uint64_t seed) {
  const orc_params *p = &c->p;
  u32 n = p->n, N = p->N;
  u32 Q = (u32)p->Q;
  u64 qKS = p->qKS;
  rng_t r;
  rng_seed(&r, seed);
  free(c->sk); free(c->z); free(c->bk_coef); free(c->ksk);
  c->sk = malloc(4 * n); c->z = malloc(4 * N);
  for (u32 i = 0; i < n; i++) c->sk[i] = rng_ternary(&r); /* LWEEncryptionScheme::KeyGen: uniform ternary */
  for (u32 i = 0; i < N; i++) c->z[i] = rng_ternary(&r);  /* KeyGenN */
  u32 *z_eval = malloc(4 * N);
  for (u32 i = 0; i < N; i++) z_eval[i] = c->z[i] < 0 ? Q - 1 : (u32)c->z[i];
  orc_ntt_fwd(c, z_eval);
  /* key switching key (LWEEncryptionScheme::KeySwitchGen): KSK[i][j][k] = (a, <a,s> + e + z_i*j*B^k) mod qKS */
  c->ksk = malloc(c->ksk_elems * c->ksk_elem_bytes);
  u64 Bk = 1;
  u64 *Bpow = malloc(8 * p->dKS);
  for (u32 k = 0; k < p->dKS; k++) { Bpow[k] = Bk; Bk *= p->baseKS; }
  for (u32 i = 0; i < N; i++)
    for (u32 j = 0; j < p->baseKS; j++)
      for (u32 k = 0; k < p->dKS; k++) {
        u64 base = (((u64)i * p->baseKS + j) * p->dKS + k) * (n + 1);
        i64 b = rng_gauss(&r, SIGMA) + (i64)c->z[i] * (i64)((u64)j * Bpow[k] % qKS);
        for (u32 t = 0; t < n; t++) {
          u64 a = rng_uniform(&r, qKS);
          ksk_store(c, base + t, (u32)a);
          b += (i64)a * c->sk[t];
        }
        b %= (i64)qKS;
        if (b < 0) b += qKS;
        ksk_store(c, base + n, (u32)b);
      }
  free(Bpow);
  /* bootstrapping key */
  c->bk_coef = malloc(c->bk_words * 4);
  size_t rgsw_words = (size_t)(2 * p->dG) * 2 * N;
  if (p->method == ORC_GINX) { /* RingGSWAccumulatorCGGI::KeyGenAcc: ek[0][0][i] = E(s_i==1), ek[0][1][i] = E(s_i==-1) */
    for (u32 i = 0; i < n; i++) {
      rgsw_encrypt(c, z_eval, c->sk[i] == 1, 0, +1, &r, c->bk_coef + ((size_t)i * 2 + 0) * rgsw_words);
      rgsw_encrypt(c, z_eval, c->sk[i] == -1, 0, +1, &r, c->bk_coef + ((size_t)i * 2 + 1) * rgsw_words);
    }
  } else { /* RingGSWAccumulatorDM::KeyGenAcc: ek[i][j][k] = E(X^{s_i * j * Br^k}) , j in [1,Br) */
    u64 Br = 1;
    for (u32 k = 0; k < p->dR; k++, Br *= p->baseR)
      for (u32 i = 0; i < n; i++)
        for (u32 j = 1; j < p->baseR; j++) {
          i64 m = (i64)c->sk[i] * (i64)j * (i64)Br;
          i64 mm = ((m % (i64)p->q) + p->q) % p->q * c->factor;
          int sign = 1;
          if (mm >= (i64)N) { mm -= N; sign = -1; }
          size_t idx = (((size_t)i * (p->baseR - 1) + (j - 1)) * p->dR + k);
          rgsw_encrypt(c, z_eval, 1, (u32)mm, sign, &r, c->bk_coef + idx * rgsw_words);
        }
  }
  free(z_eval);
  build_bk_eval(c);
  c->has_keys = 1;
  return 0;
}

size_t orc_keyblob_size(const orc_ctx *c) {
  size_t s = sizeof(keyblob_header);
  s += ((size_t)c->p.n * 4 + 7) & ~(size_t)7;
  s += (c->bk_words * 4 + 7) & ~(size_t)7;
  s += (c->ksk_elems * c->ksk_elem_bytes + 7) & ~(size_t)7;
  return s;
}
int orc_export_keys(const orc_ctx *c, void *buf, size_t cap) {
  if (!c->has_keys || cap < orc_keyblob_size(c)) return -1;
  keyblob_header h;
  memset(&h, 0, sizeof h);
  memcpy(h.magic, "BFHEKEY1", 8);
  const orc_params *p = &c->p;
  h.version = 1; h.paramset = p->paramset; h.method = p->method; h.n = p->n; h.N = p->N; h.q = p->q;
  h.baseKS = p->baseKS; h.dKS = p->dKS; h.baseG = p->baseG; h.dG = p->dG; h.baseR = p->baseR; h.dR = p->dR;
  h.has_sk = c->sk != NULL; h.ksk_elem_bytes = c->ksk_elem_bytes;
  h.Q = p->Q; h.qKS = p->qKS; h.bk_words = c->bk_words; h.ksk_elems = c->ksk_elems;
  char *o = buf;
  memcpy(o, &h, sizeof h); o += sizeof h;
  memset(o, 0, ((size_t)p->n * 4 + 7) & ~(size_t)7);
  if (c->sk) memcpy(o, c->sk, (size_t)p->n * 4);
  o += ((size_t)p->n * 4 + 7) & ~(size_t)7;
  memcpy(o, c->bk_coef, c->bk_words * 4); o += (c->bk_words * 4 + 7) & ~(size_t)7;
  memcpy(o, c->ksk, c->ksk_elems * c->ksk_elem_bytes);
  return 0;
}
int orc_import_keys(orc_ctx *c, const void *buf, size_t len) {
  const orc_params *p = &c->p;
  keyblob_header h;
  if (len < sizeof h) return -1;
  memcpy(&h, buf, sizeof h);
  if (memcmp(h.magic, "BFHEKEY1", 8) || h.version != 1) return -2;
  if (h.paramset != p->paramset || h.method != p->method || h.n != p->n || h.N != p->N || h.q != p->q ||
      h.Q != p->Q || h.qKS != p->qKS || h.baseKS != p->baseKS || h.dKS != p->dKS || h.baseG != p->baseG ||
      h.dG != p->dG || h.baseR != p->baseR || h.dR != p->dR || h.bk_words != c->bk_words ||
      h.ksk_elems != c->ksk_elems || h.ksk_elem_bytes != c->ksk_elem_bytes)
    return -3;
  if (len < orc_keyblob_size(c)) return -4;
  const char *in = (const char *)buf + sizeof h;
  free(c->sk); free(c->bk_coef); free(c->ksk); free(c->z);
  c->z = NULL;
  c->sk = NULL;
  if (h.has_sk) { c->sk = malloc(4 * p->n); memcpy(c->sk, in, 4 * p->n); }
  in += ((size_t)p->n * 4 + 7) & ~(size_t)7;
  c->bk_coef = malloc(c->bk_words * 4);
  memcpy(c->bk_coef, in, c->bk_words * 4);
  in += (c->bk_words * 4 + 7) & ~(size_t)7;
  c->ksk = malloc(c->ksk_elems * c->ksk_elem_bytes);
  memcpy(c->ksk, in, c->ksk_elems * c->ksk_elem_bytes);
  build_bk_eval(c);
  c->has_keys = 1;
  return 0;
}

/* ---------------- LWE (lwe-pke.cpp) ---------------- */
/* LWEEncryptionScheme::Encrypt with output = FRESH (src/circuit.cpp:506 uses the BOOTSTRAPPED default:
 * orc_encrypt_fresh followed by orc_bootstrap) */
void orc_encrypt_fresh(const orc_ctx *c, int bit, uint64_t seed, uint32_t *ct) {
  const orc_params *p = &c->p;
  rng_t r;
  rng_seed(&r, seed);
  u32 q = p->q;
  i64 b = (i64)(bit % 4) * (q / 4) + rng_gauss(&r, SIGMA);
  for (u32 i = 0; i < p->n; i++) {
    ct[i] = (u32)rng_uniform(&r, q);
    b += (i64)ct[i] * c->sk[i];
  }
  b %= (i64)q;
  if (b < 0) b += q;
  ct[p->n] = (u32)b;
}
/* LWEEncryptionScheme::Decrypt (src/circuit.cpp:800, src/gate.cpp:72..): r = b - <a,s> + q/8; floor(4r/q) */
int orc_decrypt(const orc_ctx *c, const uint32_t *ct) {
  const orc_params *p = &c->p;
  i64 q = p->q, r = ct[p->n];
  for (u32 i = 0; i < p->n; i++) r -= (i64)ct[i] * c->sk[i];
  r = ((r % q) + q) % q;
  r = (r + q / 8) % q;
  return (int)(4 * r / q);
}
/* BinFHEScheme::EvalNOT (src/gate.cpp:112,198,199): (-a, q/4 - b) */
void orc_eval_not(const orc_ctx *c, const uint32_t *in, uint32_t *out) {
  u32 q = c->p.q, n = c->p.n;
  for (u32 i = 0; i < n; i++) out[i] = (q - in[i]) % q;
  out[n] = (q / 4 + q - in[n]) % q;
}
/* LWE linear prep of EvalBinGate: ct1 + ct2, or 2*(ct1 - ct2) for XOR_FAST/XNOR_FAST; Bootstrap: (a, b + q/4) */
void orc_prep(const orc_ctx *c, uint32_t op, const uint32_t *ct1, const uint32_t *ct2, uint32_t *prep) {
  u32 q = c->p.q, n = c->p.n, w = n + 1;
  u32 gate = op & 0xff;
  u32 *x = malloc(4 * w), *y = malloc(4 * w);
  if (op & ORC_NEG0) orc_eval_not(c, ct1, x); else memcpy(x, ct1, 4 * w);
  if (gate == ORC_BOOTSTRAP) {
    for (u32 i = 0; i < n; i++) prep[i] = x[i];
    prep[n] = (x[n] + q / 4) % q;
  } else {
    if (op & ORC_NEG1) orc_eval_not(c, ct2, y); else memcpy(y, ct2, 4 * w);
    for (u32 i = 0; i < w; i++)
      prep[i] = (gate == ORC_XOR_FAST || gate == ORC_XNOR_FAST) ? (2 * (x[i] + q - y[i])) % q : (x[i] + y[i]) % q;
  }
  free(x); free(y);
}

/* RingGSWAccumulator::SignedDigitDecompose (rgsw-acc.cpp): in = 2 polys (coef), out = 2*dG polys, digit l of poly j -> j+2l */
void orc_signed_digit_decompose(const orc_ctx *c, const uint32_t *in, uint32_t *out) {
  const u32 N = c->p.N, dG = c->p.dG; /* dG is 3 (TOY) or 4 (STD128_OPT) */
  /* OpenFHE does this in signed 64-bit; |d| < 2^27 so signed 32-bit gives the same digits */
  const int32_t Q = (int32_t)c->p.Q, QHalf = Q >> 1;
  const int gBits = c->logBG, shift = 32 - gBits;
  for (u32 j = 0; j < 2; j++) {
    const u32 *restrict src = in + (size_t)j * N;
    u32 *restrict o0 = out + (size_t)(j + 0) * N, *restrict o1 = out + (size_t)(j + 2) * N;
    u32 *restrict o2 = out + (size_t)(j + 4) * N, *restrict o3 = dG > 3 ? out + (size_t)(j + 6) * N : o2;
#pragma omp simd
    for (u32 k = 0; k < N; k++) {
      const int32_t t = (int32_t)src[k];
      int32_t d = (t < QHalf) ? t : t - Q;
      int32_t r;
      r = (int32_t)((u32)d << shift) >> shift; d -= r; d >>= gBits; /* signed remainder, then exact division */
      o0[k] = (u32)(r + (Q & (r >> 31)));
      r = (int32_t)((u32)d << shift) >> shift; d -= r; d >>= gBits;
      o1[k] = (u32)(r + (Q & (r >> 31)));
      r = (int32_t)((u32)d << shift) >> shift; d -= r; d >>= gBits;
      const u32 v2 = (u32)(r + (Q & (r >> 31)));
      r = (int32_t)((u32)d << shift) >> shift;
      const u32 v3 = (u32)(r + (Q & (r >> 31)));
      if (dG > 3) { o2[k] = v2; o3[k] = v3; } else o2[k] = v2;
    }
  }
}

/* acc += / = sum_l dct[l] * key[l][c]  helpers */
static void ext_product(const orc_ctx *c, const u32 *dct, const u32 *key, u32 *res /* 2N eval */) {
  u32 N = c->p.N, rows = 2 * c->p.dG;
  const u64 Q = c->p.Q, mu = c->mu;
  u64 *restrict sum = (u64 *)malloc((size_t)N * 8);
  for (u32 col = 0; col < 2; col++) {
    memset(sum, 0, (size_t)N * 8); /* 2*dG <= 8 products below 2^54 each: no overflow */
    for (u32 l = 0; l < rows; l++) {
      const u32 *restrict d = dct + (size_t)l * N, *restrict kk = key + ((size_t)l * 2 + col) * N;
#pragma omp simd
      for (u32 k = 0; k < N; k++) sum[k] += (u64)d[k] * kk[k];
    }
    for (u32 k = 0; k < N; k++) res[col * N + k] = barrett64(sum[k], Q, mu);
  }
  free(sum);
}

/* BootstrapGateCore + EvalAcc (GINX: RingGSWAccumulatorCGGI::EvalAcc/AddToAcc; AP: RingGSWAccumulatorDM) */
void orc_blind_rotate(const orc_ctx *c, int gate, const uint32_t *prep, uint32_t *acc_coef) {
  const orc_params *p = &c->p;
  u32 N = p->N, n = p->n, q = p->q, rows = 2 * p->dG;
  u32 Q = (u32)p->Q;
  u32 Q8 = Q / 8 + 1, Q8neg = Q - Q8;
  u32 q1 = c->gate_const[gate == ORC_BOOTSTRAP ? ORC_AND : gate], q2 = (q1 + q / 2) % q;
  u32 *acc = calloc(2 * N, 4), *ct = malloc(2 * N * 4), *dct = malloc((size_t)rows * N * 4), *prod = malloc(2 * N * 4);
  u32 b = prep[n];
  for (u32 j = 0; j < q / 2; j++) { /* test vector */
    u32 t = (b + q - j) % q;
    u32 v;
    if (q1 < q2) v = (t >= q1 && t < q2) ? Q8neg : Q8;
    else v = (t >= q2 && t < q1) ? Q8 : Q8neg;
    acc[N + j * c->factor] = v;
  }
  orc_ntt_fwd(c, acc);
  orc_ntt_fwd(c, acc + N);
  size_t rgsw_words = (size_t)rows * 2 * N;
  u32 nsteps = (p->method == ORC_GINX) ? n : n * p->dR;
  for (u32 step = 0; step < nsteps; step++) {
    u32 i = (p->method == ORC_GINX) ? step : step / p->dR;
    u32 aneg = (q - prep[i]) % q;
    const u32 *key1 = NULL, *key2 = NULL;
    u32 idx_pos = 0, idx_neg = 0;
    if (p->method == ORC_GINX) {
      idx_pos = aneg * c->factor;                     /* in [0, 2N) */
      idx_neg = (2 * N - idx_pos) % (2 * N);
      key1 = c->bk_eval + ((size_t)i * 2 + 0) * rgsw_words;
      key2 = c->bk_eval + ((size_t)i * 2 + 1) * rgsw_words;
    } else {
      u32 k = step % p->dR;
      u32 a0 = aneg;
      for (u32 t = 0; t < k; t++) a0 /= p->baseR;
      a0 %= p->baseR;
      if (a0 == 0) continue;
      key1 = c->bk_eval + (((size_t)i * (p->baseR - 1) + (a0 - 1)) * p->dR + k) * rgsw_words;
    }
    memcpy(ct, acc, 2 * N * 4);
    orc_ntt_inv(c, ct);
    orc_ntt_inv(c, ct + N);
    orc_signed_digit_decompose(c, ct, dct);
    for (u32 l = 0; l < rows; l++) orc_ntt_fwd(c, dct + (size_t)l * N);
    if (p->method == ORC_GINX) {
      const u32 *mp = c->mono + (size_t)idx_pos * N, *mn = c->mono + (size_t)idx_neg * N;
      ext_product(c, dct, key1, prod);
      for (u32 col = 0; col < 2; col++)
        for (u32 k = 0; k < N; k++)
          acc[col * N + k] = barrett64(acc[col * N + k] + (u64)prod[col * N + k] * mp[k], Q, c->mu);
      ext_product(c, dct, key2, prod);
      for (u32 col = 0; col < 2; col++)
        for (u32 k = 0; k < N; k++)
          acc[col * N + k] = barrett64(acc[col * N + k] + (u64)prod[col * N + k] * mn[k], Q, c->mu);
    } else {
      ext_product(c, dct, key1, acc); /* AP replaces the accumulator */
    }
  }
  orc_ntt_inv(c, acc);
  orc_ntt_inv(c, acc + N);
  memcpy(acc_coef, acc, 2 * N * 4);
  free(acc); free(ct); free(dct); free(prod);
}

/* RoundqQ: floor(0.5 + v*q/Q) mod q, evaluated in IEEE double as OpenFHE does (lwe-pke.cpp) */
static u32 round_qQ(u64 v, u64 to, u64 from) {
  return (u32)((u64)floor(0.5 + (double)v * (double)to / (double)from) % to);
}
/* sample extraction + ModSwitch(Q -> qKS)  (BinFHEScheme::EvalBinGate tail) */
void orc_extract_modswitch(const orc_ctx *c, const uint32_t *acc, uint32_t *ext) {
  u32 N = c->p.N;
  u64 Q = c->p.Q, qKS = c->p.qKS;
  /* Transpose: a'_0 = a_0, a'_k = -a_{N-k} */
  for (u32 k = 0; k < N; k++) {
    u64 v = (k == 0) ? acc[0] : (Q - acc[N - k]) % Q;
    ext[k] = round_qQ(v, qKS, Q);
  }
  u64 bq = (acc[N] + (Q >> 3) + 1) % Q;
  ext[N] = round_qQ(bq, qKS, Q);
}
/* LWEEncryptionScheme::KeySwitch then ModSwitch(qKS -> q) */
void orc_keyswitch_modswitch(const orc_ctx *c, const uint32_t *ext, uint32_t *out) {
  const orc_params *p = &c->p;
  u32 N = p->N, n = p->n;
  u64 qKS = p->qKS;
  u64 *a = calloc(n + 1, 8);
  a[n] = ext[N];
  for (u32 i = 0; i < N; i++) {
    u32 atmp = ext[i];
    for (u32 j = 0; j < p->dKS; j++, atmp /= p->baseKS) {
      u32 a0 = atmp % p->baseKS;
      u64 base = (((u64)i * p->baseKS + a0) * p->dKS + j) * (n + 1);
      for (u32 k = 0; k <= n; k++) {
        u64 v = ksk_load(c, base + k);
        a[k] = a[k] >= v ? a[k] - v : a[k] + qKS - v;
      }
    }
  }
  for (u32 k = 0; k <= n; k++) out[k] = round_qQ(a[k], p->q, qKS);
  free(a);
}

static void gate_core(const orc_ctx *c, uint32_t op, const uint32_t *ct1, const uint32_t *ct2, uint32_t *out) {
  u32 N = c->p.N, n = c->p.n;
  u32 *prep = malloc(4 * (n + 1)), *acc = malloc(8 * N), *ext = malloc(4 * (N + 1));
  orc_prep(c, op, ct1, ct2, prep);
  orc_blind_rotate(c, op & 0xff, prep, acc);
  orc_extract_modswitch(c, acc, ext);
  orc_keyswitch_modswitch(c, ext, out);
  free(prep); free(acc); free(ext);
}

/* BinFHEScheme::EvalBinGate (src/gate.cpp:133,172,200-202).  XOR/XNOR are the composite
 * OR(AND(a, NOT b), AND(NOT a, b)) of binfhe 1.0.x, identical to src/gate.cpp:198-202. */
int orc_eval_bingate(const orc_ctx *c, int op, const uint32_t *ct1, const uint32_t *ct2, uint32_t *out) {
  int gate = op & 0xff;
  if (ct1 == ct2 && gate != ORC_BOOTSTRAP) return -1; /* "Please only use independent ciphertexts" */
  if (gate == ORC_XOR || gate == ORC_XNOR) {
    u32 w = c->p.n + 1;
    u32 *t1 = malloc(4 * w), *t2 = malloc(4 * w);
    /* honour outer NEG flags by toggling the inner ones */
    u32 f0 = op & ORC_NEG0, f1 = op & ORC_NEG1;
    gate_core(c, ORC_AND | f0 | (f1 ^ ORC_NEG1), ct1, ct2, t1);
    gate_core(c, ORC_AND | (f0 ^ ORC_NEG0) | f1, ct1, ct2, t2);
    gate_core(c, ORC_OR, t1, t2, out);
    if (gate == ORC_XNOR) { memcpy(t1, out, 4 * w); orc_eval_not(c, t1, out); }
    free(t1); free(t2);
    return 0;
  }
  gate_core(c, op, ct1, ct2, out);
  return 0;
}
/* BinFHEScheme::Bootstrap (reached via Encrypt's BOOTSTRAPPED default, src/circuit.cpp:506) */
void orc_bootstrap(const orc_ctx *c, const uint32_t *in, uint32_t *out) { gate_core(c, ORC_BOOTSTRAP, in, in, out); }

int orc_eval_gates(const orc_ctx *c, const orc_gate *gates, int count, uint32_t *slab, int nthreads) {
  u32 stride = c->p.ct_stride;
  int err = 0;
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
  (void)nthreads;
#pragma omp parallel for schedule(dynamic, 1)
  for (int g = 0; g < count; g++) {
    const orc_gate *d = &gates[g];
    u32 *tmp = malloc(4 * stride);
    int rc = orc_eval_bingate(c, d->op, slab + (size_t)d->in0 * stride, slab + (size_t)d->in1 * stride, tmp);
    if (rc) {
#pragma omp atomic write
      err = rc;
    } else memcpy(slab + (size_t)d->out * stride, tmp, 4 * (c->p.n + 1));
    free(tmp);
  }
  return err;
}
