// tb_circuit.cpp -- what the reference's TB_* mains do (src/TB_adder_2bit.cpp:64-104, src/test_adder.cpp:82-300),
// written against host/circuit.h exactly as against the reference's circuit.h: plaintext pass, then encrypted pass with
// gate-by-gate verify, both compared with a golden computed in plain C++.  Unlike the reference's mains the exit code
// reports failure.   usage: tb_circuit <adder.out> [-s TOY|STD128_OPT] [-m AP|GINX] [-n loops]
#include "../host/circuit.h"
#include <cstdio>
#include <cstring>

int main(int argc, char **argv) {
  if (argc < 2) { std::fprintf(stderr, "usage: %s <adder .out file> [-s set] [-m method] [-n loops]\n", argv[0]); return 2; }
  lbcrypto::BINFHE_PARAMSET set = lbcrypto::STD128_OPT;
  lbcrypto::BINFHE_METHOD method = lbcrypto::GINX;
  unsigned loops = 10;
  for (int i = 2; i + 1 < argc; i += 2) {
    if (!std::strcmp(argv[i], "-s")) set = !std::strcmp(argv[i + 1], "TOY") ? lbcrypto::TOY : lbcrypto::STD128_OPT;
    else if (!std::strcmp(argv[i], "-m")) method = !std::strcmp(argv[i + 1], "AP") ? lbcrypto::AP : lbcrypto::GINX;
    else if (!std::strcmp(argv[i], "-n")) loops = (unsigned)std::atoi(argv[i + 1]);
  }
  Circuit circ(set, method);
  if (!circ.ReadFile(argv[1])) return 1;
  uint32_t nin = 0, bits[8], nout = 0;
  bfhe_circuit_info(circ.raw(), &nin, bits, &nout, nullptr, nullptr, nullptr, nullptr);
  if (nin != 2 || bits[0] != bits[1] || nout != bits[0] + 1) { std::fprintf(stderr, "not an adder circuit\n"); return 2; }
  const unsigned n = bits[0];
  bool passed = true;
  for (unsigned t = 0; t < loops; t++) {
    srand(t); // src/test_adder.cpp:180-190
    Inputs inputs(2);
    std::vector<unsigned> in1(n), in2(n);
    for (unsigned ix = 0; ix < n; ix++) {
      in1[ix] = rand() % 2; inputs[0].push_back(in1[ix]);
      in2[ix] = rand() % 2; inputs[1].push_back(in2[ix]);
    }
    std::vector<unsigned> golden(n + 1, 0); // ripple full adder, src/test_adder.cpp:206-217
    unsigned c = 0;
    for (unsigned ix = 0; ix < n; ix++) {
      unsigned a = in1[ix], b = in2[ix], tmp = a ^ b;
      golden[ix] = c ^ tmp;
      c = (a & b) | (c & tmp);
    }
    golden[n] = c;
    circ.Reset(); circ.setPlaintext(true); circ.SetInput(inputs); // :228-233
    Outputs pout = circ.Clock();
    circ.Reset(); circ.setEncrypted(true); circ.setVerify(true); circ.SetInput(inputs); // :265-270
    Outputs eout = circ.Clock();
    bool ok = pout[0] == golden && eout[0] == golden && circ.verifyMismatches() == 0;
    std::printf("test %u %s\n", t, ok ? "passed" : "FAILED");
    passed = passed && ok;
  }
  // the single-gate API with the reference's names (src/gate.cpp:112,133,172)
  lbcrypto::BinFHEContext cc;
  cc.GenerateBinFHEContext(lbcrypto::TOY, lbcrypto::GINX);
  auto sk = cc.KeyGen();
  cc.BTKeyGen(sk);
  auto a = cc.Encrypt(sk, 1), b = cc.Encrypt(sk, 0);
  lbcrypto::LWEPlaintext r;
  cc.Decrypt(sk, cc.EvalBinGate(lbcrypto::OR, a, cc.EvalNOT(a)), &r);
  passed = passed && r == 1;
  cc.Decrypt(sk, cc.EvalBinGate(lbcrypto::AND, a, b), &r);
  passed = passed && r == 0;
  bool threw = false;
  try { cc.EvalBinGate(lbcrypto::AND, a, a); } catch (...) { threw = true; } // src/gate.cpp:134 relies on this
  passed = passed && threw;
  std::printf("%s\n", passed ? "ALL PASSED" : "SOME FAILED");
  return passed ? 0 : 1;
}
