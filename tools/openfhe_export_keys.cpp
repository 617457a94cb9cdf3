// openfhe_export_keys.cpp -- NOT built in this repository (OpenFHE is not available here).
//
// A maintainer with an OpenFHE v1.0.x build can compile this file against it to close the "parity unpinned" gap of
// it generates a BinFHE context exactly as the reference does (src/circuit.cpp:88-91 in
// openfhe-boolean-circuit-evaluator), dumps the keys in this engine's BFHEKEY1 exchange format (layout in INTEGRATION.md
// section 4; header struct = KeyBlobHeader in csrc/engine.cu), plus a few input / output ciphertexts of EvalBinGate as flat
// uint32 arrays.  tests/ can then load the blob with bfhe_load_keys(), run the same gates through the C ABI and compare the
// output ciphertexts byte for byte with OpenFHE's.
//
// The accessors below follow the v1.0.x headers (binfhecontext.h, rgsw-acckey.h, lwe-keyswitchkey.h) as recalled in
// SURVEY.md App. C; names may need small adjustments for a particular patch release.
//
//   g++ -std=c++17 openfhe_export_keys.cpp -I$OPENFHE/include/openfhe{,/core,/binfhe,/pke} -L$OPENFHE/lib \
//       -lOPENFHEbinfhe -lOPENFHEcore -o openfhe_export_keys && ./openfhe_export_keys STD128_OPT GINX keys.bfhe gates.bin
#include "binfhecontext.h"

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

using namespace lbcrypto;

struct KeyBlobHeader { // must match csrc/engine.cu
  char magic[8];
  uint32_t version, paramset, method, n, N, q;
  uint32_t baseKS, dKS, baseG, dG, baseR, dR;
  uint32_t has_sk, ksk_elem_bytes;
  uint64_t Q, qKS, bk_words, ksk_elems;
  uint64_t reserved[4];
};
static size_t pad8(size_t x) { return (x + 7) & ~(size_t)7; }
static void put(std::vector<uint8_t>& out, const void* p, size_t bytes, size_t padded) {
  const size_t at = out.size();
  out.resize(at + padded, 0);
  std::memcpy(out.data() + at, p, bytes);
}

// one RGSW ciphertext -> 2*dG rows x 2 polynomials, COEFFICIENT form, uint32 words (this engine applies its own NTT at load)
static void dump_rgsw(const RingGSWEvalKey& ek, std::vector<uint32_t>& out) {
  auto rows = ek->GetElements();  // std::vector<std::vector<NativePoly>>, [2*dG][2], EVALUATION form
  for (auto& row : rows)
    for (auto& poly : row) {
      NativePoly p(poly);
      p.SetFormat(Format::COEFFICIENT);
      for (usint i = 0; i < p.GetLength(); i++) out.push_back((uint32_t)p[i].ConvertToInt());
    }
}

int main(int argc, char** argv) {
  if (argc < 5) { std::fprintf(stderr, "usage: %s TOY|STD128_OPT AP|GINX keys.bfhe gates.bin\n", argv[0]); return 2; }
  const bool toy = !std::strcmp(argv[1], "TOY"), ap = !std::strcmp(argv[2], "AP");
  auto cc = BinFHEContext();
  cc.GenerateBinFHEContext(toy ? TOY : STD128_OPT, ap ? AP : GINX);
  auto sk = cc.KeyGen();
  cc.BTKeyGen(sk);

  const auto lwe = cc.GetParams()->GetLWEParams();
  const auto rgsw = cc.GetParams()->GetRingGSWParams();
  KeyBlobHeader h{};
  std::memcpy(h.magic, "BFHEKEY1", 8);
  h.version = 1;
  h.paramset = toy ? 0 : 5;  // BFHE_TOY / BFHE_STD128_OPT (include/bfhe.h)
  h.method = ap ? 0 : 1;     // BFHE_AP / BFHE_GINX
  h.n = lwe->Getn(); h.N = lwe->GetN(); h.q = (uint32_t)lwe->Getq().ConvertToInt();
  h.Q = lwe->GetQ().ConvertToInt(); h.qKS = lwe->GetqKS().ConvertToInt();
  h.baseKS = lwe->GetBaseKS(); h.baseG = rgsw->GetBaseG(); h.baseR = rgsw->GetBaseR();
  h.dG = rgsw->GetDigitsG(); h.dR = (uint32_t)rgsw->GetDigitsR().size();
  h.dKS = 0;
  for (uint64_t v = 1; v < h.qKS; v *= h.baseKS) h.dKS++;
  h.has_sk = 1;
  h.ksk_elem_bytes = h.qKS <= 65536 ? 2 : 4;

  // secret key: ternary, stored by OpenFHE mod qKS -> int32 in {-1, 0, 1}
  std::vector<int32_t> s(h.n);
  const auto& sv = sk->GetElement();
  for (uint32_t i = 0; i < h.n; i++) {
    const uint64_t v = sv[i].ConvertToInt(), m = sv.GetModulus().ConvertToInt();
    s[i] = v == 0 ? 0 : (v == 1 ? 1 : (v == m - 1 ? -1 : 2));
  }

  // bootstrapping key.  GINX: [i][sign: 0 = RGSW(s_i == +1), 1 = RGSW(s_i == -1)]; AP: [i][j - 1][k], j = 1 .. baseR - 1
  std::vector<uint32_t> bk;
  const auto& acc = cc.GetRefreshKey();
  if (!ap) {
    for (uint32_t i = 0; i < h.n; i++)
      for (uint32_t sign = 0; sign < 2; sign++) dump_rgsw((*acc)[0][sign][i], bk);
  } else {
    for (uint32_t i = 0; i < h.n; i++)
      for (uint32_t j = 1; j < h.baseR; j++)
        for (uint32_t k = 0; k < h.dR; k++) dump_rgsw((*acc)[i][j][k], bk);
  }
  h.bk_words = bk.size();

  // key-switching key [N][baseKS][dKS][n + 1]: a then b, mod qKS
  const auto& ks = cc.GetSwitchKey();
  const auto& A = ks->GetElementsA();
  const auto& B = ks->GetElementsB();
  std::vector<uint8_t> ksk;
  h.ksk_elems = (uint64_t)h.N * h.baseKS * h.dKS * (h.n + 1);
  for (uint32_t i = 0; i < h.N; i++)
    for (uint32_t j = 0; j < h.baseKS; j++)
      for (uint32_t k = 0; k < h.dKS; k++)
        for (uint32_t t = 0; t <= h.n; t++) {
          const uint64_t v = t < h.n ? A[i][j][k][t].ConvertToInt() : B[i][j][k].ConvertToInt();
          if (h.ksk_elem_bytes == 2) { uint16_t w = (uint16_t)v; put(ksk, &w, 2, 2); }
          else { uint32_t w = (uint32_t)v; put(ksk, &w, 4, 4); }
        }

  std::vector<uint8_t> blob;
  put(blob, &h, sizeof h, sizeof h);
  put(blob, s.data(), s.size() * 4, pad8(s.size() * 4));
  put(blob, bk.data(), bk.size() * 4, pad8(bk.size() * 4));
  put(blob, ksk.data(), ksk.size(), pad8(ksk.size()));
  FILE* f = std::fopen(argv[3], "wb");
  std::fwrite(blob.data(), 1, blob.size(), f);
  std::fclose(f);

  // golden gates: fresh (un-bootstrapped) encryptions of (a, b) and OpenFHE's outputs, (n + 1) uint32 words each: a[0..n), b
  auto dump_ct = [&](const LWECiphertext& ct, std::vector<uint32_t>& out) {
    for (uint32_t i = 0; i < h.n; i++) out.push_back((uint32_t)ct->GetA()[i].ConvertToInt());
    out.push_back((uint32_t)ct->GetB().ConvertToInt());
  };
  std::vector<uint32_t> g;
  const BINGATE gates[] = {OR, AND, NOR, NAND, XOR_FAST, XNOR_FAST};
  for (int a = 0; a < 2; a++)
    for (int b = 0; b < 2; b++) {
      auto ca = cc.Encrypt(sk, a, FRESH), cb = cc.Encrypt(sk, b, FRESH);
      dump_ct(ca, g); dump_ct(cb, g);
      for (BINGATE gt : gates) dump_ct(cc.EvalBinGate(gt, ca, cb), g);
      dump_ct(cc.EvalNOT(ca), g);
      dump_ct(cc.Bootstrap(ca), g);
    }
  f = std::fopen(argv[4], "wb");
  std::fwrite(g.data(), 4, g.size(), f);
  std::fclose(f);
  std::printf("wrote %zu key bytes, %zu ciphertext words (per (a,b): in0, in1, OR, AND, NOR, NAND, XOR_FAST, XNOR_FAST, NOT, Bootstrap)\n",
              blob.size(), g.size());
  return 0;
}
