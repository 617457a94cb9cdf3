/*
 * bfhe_oracle.h -- CPU ORACLE (TEST INFRASTRUCTURE ONLY, NOT PRODUCT CODE).
 *
 * Plain-C restatement of the gate-bootstrapping path that the reference
 * (openfheorg/openfhe-boolean-circuit-evaluator) reaches through
 * lbcrypto::BinFHEContext at src/circuit.cpp:88-91,506,800 and
 * src/gate.cpp:72,112,133,172,198-202.  The arithmetic itself lives in the
 * third-party dependency openfhe-development (module src/binfhe), effective pin
 * v1.0.1 (Release_Notes.md:4), which is NOT vendored under /root/reference and is
 * not installable here; this file restates its published algorithm (Ducas-
 * Micciancio FHEW "AP", Chillotti et al. TFHE "GINX", Micciancio-Polyakov
 * ePrint 2020/086) as summarised in SURVEY.md App. C.
 *
 * PARITY UNPINNED at ciphertext level: the reference ships no key / ciphertext /
 * NTT fixtures; what IS pinned here are the reference's decrypted
 * known-answer vectors (tests/golden) and truth tables.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library.  The product (openfhe-boolean-circuit-evaluator_b200/)
 * never does.
 */
#ifndef BFHE_ORACLE_H
#define BFHE_ORACLE_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* enum values follow OpenFHE 1.0.x binfhe-constants.h ordering */
enum { ORC_TOY = 0, ORC_STD128_OPT = 5 };
enum { ORC_AP = 0, ORC_GINX = 1 };
enum { ORC_OR = 0, ORC_AND = 1, ORC_NOR = 2, ORC_NAND = 3, ORC_XOR_FAST = 4, ORC_XNOR_FAST = 5,
       ORC_XOR = 6, ORC_XNOR = 7, ORC_BOOTSTRAP = 8 };
#define ORC_NEG0 0x100u /* operand 0 goes through EvalNOT first */
#define ORC_NEG1 0x200u

typedef struct {
  uint32_t paramset, method;
  uint32_t n, N, q;
  uint64_t Q, qKS;
  uint32_t baseKS, dKS, baseG, dG, baseR, dR;
  uint32_t ct_words;  /* n + 1 */
  uint32_t ct_stride; /* ct_words rounded up to a multiple of 4 */
} orc_params;

typedef struct {
  uint32_t op; /* gate | ORC_NEG0 | ORC_NEG1 */
  uint32_t in0, in1, out; /* slab rows */
} orc_gate;

typedef struct orc_ctx orc_ctx;

orc_ctx *orc_create(int paramset, int method);
void orc_destroy(orc_ctx *);
void orc_get_params(const orc_ctx *, orc_params *);

int orc_keygen(orc_ctx *, uint64_t seed);
size_t orc_keyblob_size(const orc_ctx *);
int orc_export_keys(const orc_ctx *, void *buf, size_t cap);
int orc_import_keys(orc_ctx *, const void *buf, size_t len);

/* LWE */
void orc_encrypt_fresh(const orc_ctx *, int bit, uint64_t seed, uint32_t *ct);
int orc_decrypt(const orc_ctx *, const uint32_t *ct);
void orc_eval_not(const orc_ctx *, const uint32_t *in, uint32_t *out);
int orc_eval_bingate(const orc_ctx *, int gate, const uint32_t *ct1, const uint32_t *ct2, uint32_t *out);
void orc_bootstrap(const orc_ctx *, const uint32_t *in, uint32_t *out);
/* one OpenMP task per gate, like src/circuit.cpp:698-710; returns 0 on success */
int orc_eval_gates(const orc_ctx *, const orc_gate *gates, int count, uint32_t *slab, int nthreads);

/* stage-level entry points for kernel parity tests */
void orc_ntt_fwd(const orc_ctx *, uint32_t *poly);
void orc_ntt_inv(const orc_ctx *, uint32_t *poly);
void orc_signed_digit_decompose(const orc_ctx *, const uint32_t *in2N, uint32_t *out);
void orc_prep(const orc_ctx *, uint32_t op, const uint32_t *ct1, const uint32_t *ct2, uint32_t *prep);
/* blind rotation of a prepared ciphertext; acc returned in COEFFICIENT form, 2*N words */
void orc_blind_rotate(const orc_ctx *, int gate, const uint32_t *prep, uint32_t *acc_coef);
void orc_extract_modswitch(const orc_ctx *, const uint32_t *acc_coef, uint32_t *ext);
void orc_keyswitch_modswitch(const orc_ctx *, const uint32_t *ext, uint32_t *out);
uint64_t orc_modulus_Q(int N);

#ifdef __cplusplus
}
#endif
#endif
