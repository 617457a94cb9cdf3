// netlist.cpp -- see netlist.hpp.
#include "netlist.hpp"
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>
#include <unordered_map>

namespace bfhe {

namespace {
struct RegMap { // register number -> dense wire id
  std::unordered_map<uint32_t, uint32_t> m;
  uint32_t next = 0;
  uint32_t get(uint32_t r) {
    auto it = m.find(r);
    if (it != m.end()) return it->second;
    m.emplace(r, next);
    return next++;
  }
};
bool has(const std::string &s, const char *t) { return s.find(t) != std::string::npos; }
} // namespace

// Same substring dispatch order as Circuit::ReadFile: LOAD, STORE, NOT, AND, " OR", XOR, BOOT
// (src/circuit.cpp:144-294); '#' lines skipped (:135-137); trailing "!depth = k" ignored by sscanf.
std::string parse_out_file(const std::string &path, Netlist &nl) {
  std::ifstream f(path);
  if (!f) return "error opening file " + path;
  nl = Netlist();
  RegMap regs;
  std::string line;
  unsigned lineNo = 0;
  int max_out = -1;
  std::vector<uint8_t> written;
  auto mark_written = [&](uint32_t w) -> bool {
    if (written.size() <= w) written.resize(w + 1, 0);
    if (written[w]) return false;
    written[w] = 1;
    return true;
  };
  while (std::getline(f, line)) {
    lineNo++;
    if (line.empty() || line[0] == '#') continue;
    unsigned n1, n2, n3;
    NetGate g{};
    auto err = [&](const char *what) { return std::string(what) + " parse error line " + std::to_string(lineNo); };
    if (has(line, "LOAD")) {
      if (sscanf(line.c_str(), "R%u = LOAD(In%u, %u)", &n1, &n2, &n3) != 3) return err("LOAD");
      if (n2 < 1) return err("LOAD (input numbers are 1-based)");
      g.kind = GateKind::INPUT; g.in0 = n2 - 1; g.in1 = n3; g.out = regs.get(n1);
      if (nl.in_bits.size() < n2) nl.in_bits.resize(n2, 0);
      if (nl.in_bits[n2 - 1] < n3 + 1) nl.in_bits[n2 - 1] = n3 + 1;
      nl.n_input++;
    } else if (has(line, "STORE")) {
      if (sscanf(line.c_str(), "Out%u = STORE(R%u)", &n1, &n2) != 2) return err("STORE");
      g.kind = GateKind::OUTPUT; g.in0 = regs.get(n2); g.out = n1;
      if ((int)n1 > max_out) max_out = (int)n1;
      nl.n_output++;
    } else if (has(line, "NOT")) {
      if (sscanf(line.c_str(), "R%u = NOT(R%u)", &n1, &n2) != 2) return err("NOT");
      g.kind = GateKind::NOT; g.in0 = regs.get(n2); g.out = regs.get(n1);
      nl.n_not++;
    } else if (has(line, "AND")) {
      if (sscanf(line.c_str(), "R%u = AND(R%u, R%u)", &n1, &n2, &n3) != 3) return err("AND");
      g.kind = GateKind::AND; g.in0 = regs.get(n2); g.in1 = regs.get(n3); g.out = regs.get(n1);
      nl.n_and++;
    } else if (has(line, " OR")) {
      if (sscanf(line.c_str(), "R%u = OR(R%u, R%u)", &n1, &n2, &n3) != 3) return err("OR");
      g.kind = GateKind::OR; g.in0 = regs.get(n2); g.in1 = regs.get(n3); g.out = regs.get(n1);
      nl.n_or++;
    } else if (has(line, "XOR")) {
      if (sscanf(line.c_str(), "R%u = XOR(R%u, R%u)", &n1, &n2, &n3) != 3) return err("XOR");
      g.kind = GateKind::XOR; g.in0 = regs.get(n2); g.in1 = regs.get(n3); g.out = regs.get(n1);
      nl.n_xor++;
    } else {
      continue; // BOOT and anything else: no-op, as in the reference
    }
    if (g.kind != GateKind::OUTPUT && !mark_written(g.out))
      return "register written twice (line " + std::to_string(lineNo) + "): the netlist must be single-assignment";
    nl.gates.push_back(g);
  }
  nl.n_wires = regs.next;
  nl.out_bits = (uint32_t)(max_out + 1);
  return "";
}

// Old Bristol: "ngates nwires" / "n1 n2 n3" / blank / gates.  New ("Bristol fashion"): "ngates nwires" /
// "niv n1 n2 .." / "nov m1 .." / blank / gates.  Inputs are the first wires, outputs the last
// (src/assemble.cpp:155-193).  Gate lines: "2 1 a b o XOR|AND", "1 1 a o INV", new format also "1 1 a o EQW".
std::string parse_bristol_file(const std::string &path, bool new_format, Netlist &nl) {
  FILE *f = std::fopen(path.c_str(), "r");
  if (!f) return "error opening file " + path;
  nl = Netlist();
  char buf[512];
  auto fail = [&](const std::string &m) { std::fclose(f); return m; };
  unsigned ngates = 0, nwires = 0;
  if (!std::fgets(buf, sizeof buf, f) || sscanf(buf, "%u %u", &ngates, &nwires) != 2) return fail("bad Bristol header");
  std::vector<uint32_t> outs;
  if (!std::fgets(buf, sizeof buf, f)) return fail("bad Bristol header");
  {
    std::istringstream ss(buf);
    std::vector<uint32_t> v;
    uint32_t x;
    while (ss >> x) v.push_back(x);
    if (new_format) {
      if (v.empty() || v.size() != v[0] + 1) return fail("bad Bristol-fashion input line");
      nl.in_bits.assign(v.begin() + 1, v.end());
      if (!std::fgets(buf, sizeof buf, f)) return fail("bad Bristol-fashion output line");
      std::istringstream s2(buf);
      std::vector<uint32_t> o;
      while (s2 >> x) o.push_back(x);
      if (o.empty() || o.size() != o[0] + 1) return fail("bad Bristol-fashion output line");
      outs.assign(o.begin() + 1, o.end());
    } else {
      if (v.size() != 3) return fail("bad old-Bristol I/O line");
      nl.in_bits = {v[0], v[1]};
      outs = {v[2]};
    }
  }
  while (!nl.in_bits.empty() && nl.in_bits.back() == 0) nl.in_bits.pop_back(); // e.g. "512 0 160"
  uint32_t n_in = 0, n_out = 0;
  for (auto b : nl.in_bits) n_in += b;
  for (auto b : outs) n_out += b;
  if (n_in + n_out > nwires && n_out > nwires) return fail("inconsistent Bristol header");
  nl.n_wires = nwires;
  nl.out_bits = n_out; // all output buses concatenated onto OUT:0
  uint32_t w = 0;
  for (uint32_t bus = 0; bus < nl.in_bits.size(); bus++)
    for (uint32_t bit = 0; bit < nl.in_bits[bus]; bit++) {
      NetGate g{};
      g.kind = GateKind::INPUT; g.in0 = bus; g.in1 = bit; g.out = w++;
      nl.gates.push_back(g);
      nl.n_input++;
    }
  unsigned lineNo = new_format ? 3 : 2, seen = 0;
  while (std::fgets(buf, sizeof buf, f)) {
    lineNo++;
    unsigned ni, no, a, b, o;
    char op[32];
    NetGate g{};
    if (sscanf(buf, "%u %u", &ni, &no) != 2) continue; // blank line
    if (ni == 2 && no == 1 && sscanf(buf, "%*u %*u %u %u %u %31s", &a, &b, &o, op) == 4) {
      if (!std::strcmp(op, "XOR")) { g.kind = GateKind::XOR; nl.n_xor++; }
      else if (!std::strcmp(op, "AND")) { g.kind = GateKind::AND; nl.n_and++; }
      else if (!std::strcmp(op, "OR")) { g.kind = GateKind::OR; nl.n_or++; }
      else return fail(std::string("unsupported gate ") + op + " at line " + std::to_string(lineNo));
      g.in0 = a; g.in1 = b; g.out = o;
    } else if (ni == 1 && no == 1 && sscanf(buf, "%*u %*u %u %u %31s", &a, &o, op) == 3) {
      if (!std::strcmp(op, "INV") || !std::strcmp(op, "NOT")) {
        g.kind = GateKind::NOT; g.in0 = a; g.out = o; nl.n_not++;
      } else if (!std::strcmp(op, "EQW")) { // wire copy: two inversions keep the netlist in the reference's gate set
        NetGate t{};
        t.kind = GateKind::NOT; t.in0 = a; t.out = nl.n_wires++;
        nl.gates.push_back(t);
        g.kind = GateKind::NOT; g.in0 = t.out; g.out = o;
        nl.n_not += 2;
      } else return fail(std::string("unsupported gate ") + op + " at line " + std::to_string(lineNo));
    } else {
      return fail("unsupported gate arity at line " + std::to_string(lineNo));
    }
    if (g.in0 >= nl.n_wires || g.in1 >= nl.n_wires || g.out >= nl.n_wires) return fail("wire id out of range at line " + std::to_string(lineNo));
    nl.gates.push_back(g);
    seen++;
  }
  std::fclose(f);
  if (seen != ngates) return "gate count mismatch: header says " + std::to_string(ngates) + ", file has " + std::to_string(seen);
  for (uint32_t i = 0; i < n_out; i++) {
    NetGate g{};
    g.kind = GateKind::OUTPUT; g.in0 = nwires - n_out + i; g.out = i;
    nl.gates.push_back(g);
    nl.n_output++;
  }
  return "";
}

} // namespace bfhe
