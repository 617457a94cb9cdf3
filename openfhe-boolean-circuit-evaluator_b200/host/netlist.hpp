// netlist.hpp -- circuit front-end: the reference's ".out" assembler format (Circuit::ReadFile,
// src/circuit.cpp:102-366; grammar in SURVEY.md App. B) and old/new Bristol netlists
// (analyze_bristol + assemble_bristol, src/analyze.cpp:56-394, src/assemble.cpp:46-429) parsed
// straight into an integer netlist in O(G) -- no "<stem>_FHE.out" text round trip, no O(G^2) fan-out scan.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace bfhe {

enum class GateKind : uint8_t { INPUT, OUTPUT, NOT, AND, OR, XOR }; // GateEnum of src/gate.h:51 (DFF/LUT3/LUT4 are unimplemented stubs there)

struct NetGate {
  GateKind kind;
  uint32_t in0 = 0, in1 = 0; // wire ids (INPUT: bus, bit)
  uint32_t out = 0;          // wire id   (OUTPUT: output bit index)
};

struct Netlist {
  std::vector<NetGate> gates;      // file order
  uint32_t n_wires = 0;
  std::vector<uint32_t> in_bits;   // width of each input bus (In1, In2, ...)
  uint32_t out_bits = 0;           // single output bus "OUT:0" (src/circuit.cpp:183-185)
  uint32_t n_input = 0, n_output = 0, n_and = 0, n_or = 0, n_xor = 0, n_not = 0;
};

// returns empty string on success, else an error message
std::string parse_out_file(const std::string &path, Netlist &nl);
std::string parse_bristol_file(const std::string &path, bool new_format, Netlist &nl);

} // namespace bfhe
