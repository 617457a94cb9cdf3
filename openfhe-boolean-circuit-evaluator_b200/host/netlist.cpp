// netlist.cpp -- see netlist.hpp.
#include "netlist.hpp"
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>
#include <algorithm>
#include <map>
#include <unordered_map>

namespace bfhe {

namespace {
struct RegMap { // register number -> dense wire id
  std::unordered_map<uint32_t, uint32_t> m;
  std::vector<uint32_t> regs; // wire id -> register number
  uint32_t next = 0;
  uint32_t get(uint32_t r) {
    auto it = m.find(r);
    if (it != m.end()) return it->second;
    m.emplace(r, next);
    regs.push_back(r);
    return next++;
  }
};
bool has(const std::string &s, const char *t) { return s.find(t) != std::string::npos; }
} // namespace

void Netlist::count(GateKind k) {
  switch (k) {
  case GateKind::INPUT: n_input++; break;
  case GateKind::OUTPUT: n_output++; break;
  case GateKind::NOT: n_not++; break;
  case GateKind::AND: n_and++; break;
  case GateKind::OR: n_or++; break;
  case GateKind::XOR: n_xor++; break;
  case GateKind::DFF: n_dff++; break;
  case GateKind::LUT3: n_lut3++; break;
  case GateKind::LUT4: n_lut4++; break;
  case GateKind::NAND: n_nand++; break;
  case GateKind::NOR: n_nor++; break;
  case GateKind::XNOR: n_xnor++; break;
  case GateKind::XOR_FAST: n_xor_fast++; break;
  case GateKind::XNOR_FAST: n_xnor_fast++; break;
  default: break;
  }
}

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
    } else if (has(line, "LUT3")) { // extensions of the reference grammar (its GateEnum declares DFF / LUT3 / LUT4, src/gate.h:51)
      unsigned n4, tt;
      if (sscanf(line.c_str(), "R%u = LUT3(R%u, R%u, R%u, %i)", &n1, &n2, &n3, &n4, (int *)&tt) != 5 || tt > 0xffu) return err("LUT3");
      g.kind = GateKind::LUT3; g.in0 = regs.get(n2); g.in1 = regs.get(n3); g.in2 = regs.get(n4); g.table = tt; g.out = regs.get(n1);
    } else if (has(line, "LUT4")) {
      unsigned n4, n5, tt;
      if (sscanf(line.c_str(), "R%u = LUT4(R%u, R%u, R%u, R%u, %i)", &n1, &n2, &n3, &n4, &n5, (int *)&tt) != 6 || tt > 0xffffu) return err("LUT4");
      g.kind = GateKind::LUT4; g.in0 = regs.get(n2); g.in1 = regs.get(n3); g.in2 = regs.get(n4); g.in3 = regs.get(n5); g.table = tt;
      g.out = regs.get(n1);
    } else if (has(line, "DFF")) {
      if (sscanf(line.c_str(), "R%u = DFF(R%u)", &n1, &n2) != 2) return err("DFF");
      g.kind = GateKind::DFF; g.in0 = regs.get(n2); g.out = regs.get(n1);
    } else if (has(line, "XNOR_FAST") || has(line, "XOR_FAST") || has(line, "XNOR") || has(line, "NAND") || has(line, "NOR")) {
      char nm[16];
      if (sscanf(line.c_str(), "R%u = %15[A-Z_](R%u, R%u)", &n1, nm, &n2, &n3) != 4) return err("gate");
      const std::string k(nm);
      g.kind = k == "XNOR_FAST" ? GateKind::XNOR_FAST : k == "XOR_FAST" ? GateKind::XOR_FAST : k == "XNOR" ? GateKind::XNOR
               : k == "NAND" ? GateKind::NAND : k == "NOR" ? GateKind::NOR : GateKind::KIND_COUNT;
      if (g.kind == GateKind::KIND_COUNT) return err("gate name");
      g.in0 = regs.get(n2); g.in1 = regs.get(n3); g.out = regs.get(n1);
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
    if (g.kind >= GateKind::DFF) nl.count(g.kind);
    if (g.kind != GateKind::OUTPUT && !mark_written(g.out))
      return "register written twice (line " + std::to_string(lineNo) + "): the netlist must be single-assignment";
    nl.gates.push_back(g);
  }
  nl.n_wires = regs.next;
  nl.wire_reg = regs.regs;
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
      else if (!std::strcmp(op, "NAND")) { g.kind = GateKind::NAND; nl.n_nand++; }
      else if (!std::strcmp(op, "NOR")) { g.kind = GateKind::NOR; nl.n_nor++; }
      else if (!std::strcmp(op, "XNOR")) { g.kind = GateKind::XNOR; nl.n_xnor++; }
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
  nl.wire_reg.resize(nl.n_wires);
  for (uint32_t i = 0; i < nl.n_wires; i++) nl.wire_reg[i] = i;
  return "";
}

// ---------------------------------------------------------------------------------------------------------------------
// LUT3 / LUT4 lowering.  The reference declares these gate types (src/gate.h:51) and leaves them as stubs
// (src/gate.cpp:220-225); with the Boolean (p = 4) plaintext space of the gate bootstrap a k-input table is evaluated as a
// Shannon expansion over 2-input gates: f = (s AND f1) OR (NOT s AND f0), where constant and literal cofactors fold away and
// a non-degenerate 2-input cofactor is ONE gate (AND / OR with free operand inversions, or XOR).  NOTs cost no bootstrap.
// ---------------------------------------------------------------------------------------------------------------------
namespace {
struct Lit { // a Boolean value during lowering: constant, or wire with polarity
  enum { C0, C1, W } k;
  uint32_t w;
  bool neg;
};
struct Lower {
  Netlist &nl;
  std::vector<NetGate> out;
  uint32_t fresh() { nl.wire_reg.push_back(0x80000000u | nl.n_wires); return nl.n_wires++; }
  uint32_t emit(GateKind k, uint32_t a, uint32_t b, int target) {
    NetGate g{};
    g.kind = k; g.in0 = a; g.in1 = b; g.out = target >= 0 ? (uint32_t)target : fresh();
    out.push_back(g);
    return g.out;
  }
  uint32_t as_wire(const Lit &l) { return l.neg ? emit(GateKind::NOT, l.w, 0, -1) : l.w; } // a NOT is an operand flag for its consumers
  Lit gate2(GateKind k, const Lit &a, const Lit &b, int target) { return Lit{Lit::W, emit(k, as_wire(a), as_wire(b), target), false}; }
  static Lit inv(Lit l) {
    if (l.k == Lit::C0) l.k = Lit::C1;
    else if (l.k == Lit::C1) l.k = Lit::C0;
    else l.neg = !l.neg;
    return l;
  }
  static bool same(const Lit &a, const Lit &b) { return a.k == b.k && (a.k != Lit::W || (a.w == b.w && a.neg == b.neg)); }
  // f over in[0..k): bit (x0 | x1 << 1 | ...) of tt; target >= 0: the LAST gate emitted for a non-literal result drives that wire
  Lit build(const uint32_t *in, int k, uint32_t tt, int target) {
    const uint32_t full = (k >= 5) ? 0xffffffffu : ((1u << (1u << k)) - 1);
    tt &= full;
    if (tt == 0) return Lit{Lit::C0, 0, false};
    if (tt == full) return Lit{Lit::C1, 0, false};
    if (k == 1) return Lit{Lit::W, in[0], tt == 1};
    const int half = 1 << (k - 1);
    const uint32_t lo = tt & ((1u << half) - 1), hi = tt >> half; // cofactors for in[k-1] = 0 / 1
    if (lo == hi) return build(in, k - 1, lo, target);
    const Lit s{Lit::W, in[k - 1], false};
    if (k == 2) {
      const Lit a{Lit::W, in[0], false};
      if (in[0] == in[1]) return build(in, 1, ((tt >> 3) & 1) << 1 | (tt & 1), target); // f(x, x)
      switch (tt) {
      case 0x8: return gate2(GateKind::AND, a, s, target);
      case 0x4: return gate2(GateKind::AND, inv(a), s, target);
      case 0x2: return gate2(GateKind::AND, a, inv(s), target);
      case 0x1: return gate2(GateKind::AND, inv(a), inv(s), target);
      case 0xe: return gate2(GateKind::OR, a, s, target);
      case 0xd: return gate2(GateKind::OR, inv(a), s, target);
      case 0xb: return gate2(GateKind::OR, a, inv(s), target);
      case 0x7: return gate2(GateKind::OR, inv(a), inv(s), target);
      case 0x6: return gate2(GateKind::XOR, a, s, target);
      case 0x9: return inv(gate2(GateKind::XOR, a, s, -1));
      default: break; // 0xa / 0x5 (a, !a) have lo == hi; 0xc / 0x3 (s, !s) fall to the mux below and fold to a literal
      }
    }
    const Lit f0 = build(in, k - 1, lo, -1), f1 = build(in, k - 1, hi, -1);
    // mux(s, f1, f0) with constant cofactors folded
    if (f1.k == Lit::C1 && f0.k == Lit::C0) return s;
    if (f1.k == Lit::C0 && f0.k == Lit::C1) return inv(s);
    if (f1.k == Lit::C1) return gate2(GateKind::OR, s, f0, target);
    if (f1.k == Lit::C0) return gate2(GateKind::AND, inv(s), f0, target);
    if (f0.k == Lit::C1) return gate2(GateKind::OR, inv(s), f1, target);
    if (f0.k == Lit::C0) return gate2(GateKind::AND, s, f1, target);
    if (same(f0, inv(f1)) && !(f0.w == s.w)) return f0.neg ? inv(gate2(GateKind::XOR, s, inv(f0), -1)) : gate2(GateKind::XOR, s, f0, target);
    if (f0.w == s.w || f1.w == s.w) { // a cofactor that is the selector itself: s ? f1 : s = s AND f1, s ? s : f0 = s OR f0, ...
      const Lit g1 = f1.w == s.w ? (f1.neg ? Lit{Lit::C0, 0, false} : Lit{Lit::C1, 0, false}) : f1;
      const Lit g0 = f0.w == s.w ? (f0.neg ? Lit{Lit::C1, 0, false} : Lit{Lit::C0, 0, false}) : f0;
      if (g1.k == Lit::C1 && g0.k == Lit::C0) return s;
      if (g1.k == Lit::C0 && g0.k == Lit::C1) return inv(s);
      if (g1.k == Lit::C1) return gate2(GateKind::OR, s, g0, target);
      if (g1.k == Lit::C0) return gate2(GateKind::AND, inv(s), g0, target);
      if (g0.k == Lit::C1) return gate2(GateKind::OR, inv(s), g1, target);
      return gate2(GateKind::AND, s, g1, target);
    }
    const Lit t1 = gate2(GateKind::AND, s, f1, -1), t0 = gate2(GateKind::AND, inv(s), f0, -1);
    return gate2(GateKind::OR, t1, t0, target);
  }
};
} // namespace

std::string lower_luts(Netlist &nl) {
  bool any = false;
  for (const NetGate &g : nl.gates) any = any || g.kind == GateKind::LUT3 || g.kind == GateKind::LUT4 || g.kind == GateKind::XNOR;
  if (!any) return "";
  if (nl.wire_reg.size() < nl.n_wires) nl.wire_reg.resize(nl.n_wires, 0);
  Lower L{nl, {}};
  L.out.reserve(nl.gates.size() * 2);
  for (const NetGate &g : nl.gates) {
    if (g.kind == GateKind::XNOR) { // EvalBinGate(XNOR) = the composite XOR followed by EvalNOT
      L.emit(GateKind::NOT, L.emit(GateKind::XOR, g.in0, g.in1, -1), 0, (int)g.out);
      continue;
    }
    if (g.kind != GateKind::LUT3 && g.kind != GateKind::LUT4) { L.out.push_back(g); continue; }
    const int k = g.kind == GateKind::LUT3 ? 3 : 4;
    const uint32_t in[4] = {g.in0, g.in1, g.in2, g.in3};
    const size_t before = L.out.size();
    const Lit r = L.build(in, k, g.table, (int)g.out);
    if (r.k != Lit::W) return "constant LUT (table " + std::to_string(g.table) + "): a circuit has no constant wires";
    const bool drives = L.out.size() > before && L.out.back().out == g.out && !r.neg && r.w == g.out;
    if (!drives) { // literal or inverted result: out = NOT(w) or NOT(NOT(w)); both fold into operand flags downstream
      if (r.neg) L.emit(GateKind::NOT, r.w, 0, (int)g.out);
      else L.emit(GateKind::NOT, L.emit(GateKind::NOT, r.w, 0, -1), 0, (int)g.out);
    }
  }
  nl.gates.swap(L.out);
  return "";
}

static std::string gate_name(GateKind k) {
  static const char *names[] = {"INPUT", "OUTPUT", "NOT", "AND", "OR", "XOR", "DFF", "LUT3", "LUT4", "NAND", "NOR", "XNOR", "XOR_FAST", "XNOR_FAST"};
  return names[(int)k];
}
static std::string wire_name(const Netlist &nl, uint32_t w) {
  const uint32_t r = w < nl.wire_reg.size() ? nl.wire_reg[w] : w;
  return (r & 0x80000000u) ? "T:" + std::to_string(r & 0x7fffffffu) : "R:" + std::to_string(r); // T: wires made by LUT lowering
}
// Circuit::dumpNetList (src/circuit.cpp:844-855): the map wire name -> names of the gates reading it.  The reference builds it
// from the output wires of every gate (src/circuit.cpp:326-354: INPUT gates first, then all others; OUTPUT gates contribute
// "OUT:0" and "BIT:<n>" with no readers) and prints it in std::map order; gate names are "<KIND>:<gate number in file order>".
std::string dump_netlist_text(const Netlist &nl) {
  std::map<std::string, std::vector<std::string>> m;
  std::vector<std::vector<uint32_t>> readers(nl.n_wires); // O(G): the reference rescans every gate per wire
  for (uint32_t i = 0; i < nl.gates.size(); i++) {
    const NetGate &g = nl.gates[i];
    if (g.kind == GateKind::INPUT) continue;
    const int nin = g.kind == GateKind::OUTPUT || g.kind == GateKind::NOT || g.kind == GateKind::DFF ? 1 : g.kind == GateKind::LUT3 ? 3
                    : g.kind == GateKind::LUT4 ? 4 : 2;
    const uint32_t in[4] = {g.in0, g.in1, g.in2, g.in3};
    for (int k = 0; k < nin; k++) readers[in[k]].push_back(i);
  }
  for (uint32_t i = 0; i < nl.gates.size(); i++) {
    const NetGate &g = nl.gates[i];
    if (g.kind == GateKind::OUTPUT) { m.insert({"OUT:0", {}}); m.insert({"BIT:" + std::to_string(g.out), {}}); continue; }
    std::vector<std::string> fan;
    for (uint32_t r : readers[g.out]) fan.push_back(gate_name(nl.gates[r].kind) + ":" + std::to_string(r));
    m.insert({wire_name(nl, g.out), fan});
  }
  std::string s = "Netlist \n";
  for (const auto &it : m) {
    s += it.first;
    for (const auto &f : it.second) s += " " + f;
    s += "\n";
  }
  return s;
}
// Circuit::dumpGates (src/circuit.cpp:856-865)
std::string dump_gates_text(const Netlist &nl) {
  std::string s = "Inputlist \n";
  for (uint32_t i = 0; i < nl.gates.size(); i++)
    if (nl.gates[i].kind == GateKind::INPUT) s += "INPUT:" + std::to_string(i) + "\n";
  s += "Alllist \n";
  for (uint32_t i = 0; i < nl.gates.size(); i++)
    if (nl.gates[i].kind != GateKind::INPUT) s += gate_name(nl.gates[i].kind) + ":" + std::to_string(i) + "\n";
  return s;
}

} // namespace bfhe
