// netlist.hpp -- circuit front-end: the reference's ".out" assembler format (Circuit::ReadFile,
// src/circuit.cpp:102-366; grammar in SURVEY.md App. B) and old/new Bristol netlists
// (analyze_bristol + assemble_bristol, src/analyze.cpp:56-394, src/assemble.cpp:46-429) parsed
// straight into an integer netlist in O(G) -- no "<stem>_FHE.out" text round trip, no O(G^2) fan-out scan.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace bfhe {

// GateEnum of src/gate.h:51 first (INPUT..LUT4: DFF / LUT3 / LUT4 are declared there but are "remember to write" stubs,
// src/gate.cpp:217-225); then the gate types OpenFHE's EvalBinGate offers natively as ONE bootstrap each
// (BINGATE NAND / NOR / XOR_FAST / XNOR_FAST) and the composite XNOR.
enum class GateKind : uint8_t { INPUT, OUTPUT, NOT, AND, OR, XOR, DFF, LUT3, LUT4, NAND, NOR, XNOR, XOR_FAST, XNOR_FAST, KIND_COUNT };

struct NetGate {
  GateKind kind;
  uint32_t in0 = 0, in1 = 0; // wire ids (INPUT: bus, bit; DFF: in0 = D)
  uint32_t out = 0;          // wire id   (OUTPUT: output bit index; DFF: Q)
  uint32_t in2 = 0, in3 = 0; // LUT3 / LUT4 only
  uint32_t table = 0;        // LUT truth table: bit (in0 | in1 << 1 | in2 << 2 | in3 << 3) is the output
};

struct Netlist {
  std::vector<NetGate> gates;      // file order
  uint32_t n_wires = 0;
  std::vector<uint32_t> in_bits;   // width of each input bus (In1, In2, ...)
  uint32_t out_bits = 0;           // single output bus "OUT:0" (src/circuit.cpp:183-185)
  uint32_t n_input = 0, n_output = 0, n_and = 0, n_or = 0, n_xor = 0, n_not = 0; // as read (before LUT lowering)
  uint32_t n_dff = 0, n_lut3 = 0, n_lut4 = 0, n_nand = 0, n_nor = 0, n_xnor = 0, n_xor_fast = 0, n_xnor_fast = 0;
  std::vector<uint32_t> wire_reg;  // wire id -> register number of the source file ("R:<n>" in the reference's wire names)
  void count(GateKind k);
};

// returns empty string on success, else an error message
std::string parse_out_file(const std::string &path, Netlist &nl);
std::string parse_bristol_file(const std::string &path, bool new_format, Netlist &nl);
// LUT3 / LUT4 -> NOT / AND / OR / XOR by Shannon expansion (at most 5 / 13 bootstraps; constants and literals folded); other
// gates are passed through.  The counters of `nl` keep describing the file; returns an error text for constant LUTs.
std::string lower_luts(Netlist &nl);
// the reference's Circuit::dumpNetList / dumpGates text (src/circuit.cpp:844-865): wire name -> fan-out gate names, in the
// order of the reference's std::map<std::string, ...>; input gate names then all gate names
std::string dump_netlist_text(const Netlist &nl);
std::string dump_gates_text(const Netlist &nl);

} // namespace bfhe
