// See frontend.hpp.  Integer/text semantics follow the reference line by line;
// the code is ours (no Boost, no bison/flex output).
#include "frontend.hpp"
#include <thread>
#include <iterator>
#include <cmath>

#include <pwd.h>
#include <string.h>
#include <time.h>
#include <unistd.h>
#include <fcntl.h>
#include <sys/stat.h>

#include <algorithm>
#include <type_traits>
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <regex>
#include <set>
#include <sstream>

namespace lbfe {

// ======================================================================= PDB

// parseStringAs<std::string> (src/utils.cpp:228-242): blanks removed; empty when the
// field is not entirely inside the line.
static std::string fieldStr(const std::string& s, unsigned pos, unsigned n) {
  std::string v;
  if ((size_t)pos + n > s.size()) return v;
  for (unsigned i = pos; i < pos + n; ++i)
    if (s[i] != ' ') v += s[i];
  return v;
}

// A field that is just [blanks][sign]digits[.digits][blanks] -- every number a PDB writer prints -- goes
// straight to strtof, which is what the stream extraction below ends up calling for it (libstdc++ num_get
// collects the characters and converts them with strtof): same value, a tenth of the time.
static bool plainDecimal(const char* p, unsigned n, float* out) {
  char buf[24];
  if (n >= sizeof(buf)) return false;
  unsigned i = 0, digits = 0, dots = 0;
  while (i < n && p[i] == ' ') ++i;
  const unsigned start = i;
  if (i < n && (p[i] == '-' || p[i] == '+')) ++i;
  for (; i < n; ++i) {
    if (p[i] >= '0' && p[i] <= '9') ++digits;
    else if (p[i] == '.' && !dots) ++dots;
    else break;
  }
  const unsigned end = i;
  for (; i < n; ++i)
    if (p[i] != ' ') return false;
  if (!digits) return false;
  memcpy(buf, p + start, end - start);
  buf[end - start] = 0;
  *out = strtof(buf, nullptr);
  return true;
}

// parseStringAs<T> (src/utils.hpp:324-360): stream extraction from the (clipped) field.
template <typename T>
static T fieldNum(const std::string& s, unsigned pos, unsigned nelem) {
  if (pos >= s.size()) throw Error("Missing Field at position " + std::to_string(pos) + "\n> " + s + "\n");
  unsigned n = nelem;
  if ((size_t)pos + n > s.size()) n = (unsigned)(s.size() - pos + 1);
  if (std::is_same<T, float>::value) {
    float f;
    if (plainDecimal(s.data() + pos, (unsigned)std::min<size_t>(n, s.size() - pos), &f)) return (T)f;
  }
  std::istringstream iss(s.substr(pos, n));
  T val(0);
  if (!(iss >> val))  // the reference underlines the field (src/utils.hpp:343-355)
    throw Error("PARSE ERROR\n" + s + "\n" + std::string(pos, ' ') + std::string(n > 1 ? n : 1, '^') + "\n");
  return val;
}

// Hybrid-36 (the cctbx / iotbx.pdb convention for serial numbers that outgrow a fixed-width field): a
// field of w characters is an ordinary decimal number while the count fits; the values after 10^w - 1
// are written in base 36 with upper-case letters -- "A000..." being the first, 10^w -- and after those
// 26 * 36^(w-1) words the same again in lower case.  What the reference's reader
// (parseStringAsHybrid36, src/utils.cpp:253-311) does with malformed fields is kept, because the PDB
// fixtures pin it: a leading '-' then blanks are skipped, the width that selects the alphabet is what
// remains after them, only fields of 1..5 significant characters can be base 36, a field that runs
// past the end of the line reads as 0, and a stray letter inside a decimal field counts from 'a' = 10.
int parseHybrid36(const std::string& line, unsigned pos, unsigned width) {
  const size_t field = width ? width : line.size() - pos;
  if ((size_t)pos + field > line.size()) return 0;
  if (field > 6) throw std::logic_error("Requested size exceeds max");
  const char* p = line.data() + pos;
  const char* const end = p + field;
  const bool minus = p < end && *p == '-';
  if (minus) ++p;
  while (p < end && *p == ' ') ++p;
  const size_t w = (size_t)(end - p);
  enum Alphabet { DECIMAL, UPPER, LOWER };
  Alphabet alphabet = DECIMAL;
  if (w >= 1 && w <= 5) alphabet = *p >= 'a' ? LOWER : (*p >= 'A' ? UPPER : DECIMAL);
  const int radix = alphabet == DECIMAL ? 10 : 36;
  const char ten = alphabet == UPPER ? 'A' : 'a';  // the character worth 10
  int value = 0;
  for (; p < end; ++p) value = value * radix + (*p >= ten ? *p - ten + 10 : *p - '0');
  if (alphabet != DECIMAL) {
    int first_word = 1, decimal_span = 10;  // 36^(w-1) and 10^w
    for (size_t k = 1; k < w; ++k) { first_word *= 36; decimal_span *= 10; }
    // upper case starts at 10^w with the word worth 10 * 36^(w-1); lower case 26 * 36^(w-1) later
    value += decimal_span + (alphabet == UPPER ? -10 : 16) * first_word;
  }
  return minus ? -value : value;
}

static void parseAtomRecord(const std::string& s, Model& m) {  // src/pdb.cpp:66-141
  Atom a;
  a.index = (uint32_t)m.size();
  a.record = fieldStr(s, 0, 6);
  a.id = parseHybrid36(s, 6, 5);
  a.name = fieldStr(s, 12, 4);
  a.altloc = fieldStr(s, 16, 1);
  a.resname = fieldStr(s, 17, 4);
  a.chain = fieldStr(s, 21, 1);
  a.resid = parseHybrid36(s, 22, 4);
  a.icode = fieldStr(s, 26, 1);
  if (!a.icode.empty() && isdigit((unsigned char)a.icode[0])) {  // frame-shifted resid (non-strict policy)
    a.resid = fieldNum<int>(s, 22, 5);
    a.icode = " ";  // the reference stores the blank it substitutes (src/pdb.cpp:106-112)
  }
  a.x = fieldNum<float>(s, 30, 8);
  a.y = fieldNum<float>(s, 38, 8);
  a.z = fieldNum<float>(s, 46, 8);
  if (s.size() > 54) {
    a.occupancy = fieldNum<float>(s, 54, 6);
    if (s.size() > 60) {
      a.bfactor = fieldNum<float>(s, 60, 6);
      if (s.size() > 72) {
        a.segid = fieldStr(s, 72, 4);
        if (s.size() > 76) a.element = fieldStr(s, 76, 2);
      }
    }
  }
  m.push_back(std::move(a));
}

Model readPDB(std::istream& is) {
  Model m;
  std::string line;
  while (std::getline(is, line)) {
    if (line.compare(0, 4, "ATOM") == 0 || line.compare(0, 6, "HETATM") == 0) parseAtomRecord(line, m);
    else if (line.compare(0, 3, "END") == 0) break;
    // REMARK / CONECT / CRYST1 / TER and unknown records do not affect the RMSD path
  }
  return m;
}

Model readPDB(const std::string& path) {
  // the whole file in one read, then lines out of memory (an ifstream costs more per line than the parsing)
  PosFile pf;
  pf.open(path);
  std::string text((size_t)pf.size(), '\0');
  if (pf.size()) pf.readAt(0, &text[0], text.size(), "Cannot read " + path);
  Model m;
  m.reserve(text.size() / 64 + 16);  // ATOM lines are 67-81 bytes
  std::string line;
  size_t p = 0;
  while (p < text.size()) {
    size_t e = text.find('\n', p);
    if (e == std::string::npos) e = text.size();
    line.assign(text, p, e - p);
    p = e + 1;
    if (line.compare(0, 4, "ATOM") == 0 || line.compare(0, 6, "HETATM") == 0) parseAtomRecord(line, m);
    else if (line.compare(0, 3, "END") == 0) break;
  }
  return m;
}

// ================================================================= selection

namespace {

enum Tok { T_END, T_NUM, T_STR, T_SKEY, T_NKEY, T_ALL, T_HYD, T_BB, T_LT, T_LE, T_GE, T_GT, T_EQ, T_NE,
           T_REGEX, T_ARROW, T_NOT, T_AND, T_OR, T_LP, T_RP, T_CHAR };
struct Token { Tok t; std::string s; long n; };

// src/scanner.ll:29-81 -- flex semantics: longest match, first rule on ties, no word
// boundaries.
std::vector<Token> lex(const std::string& in) {
  static const struct { const char* lit; Tok t; } fixed[] = {
      {"<", T_LT}, {"<=", T_LE}, {">=", T_GE}, {">", T_GT}, {"==", T_EQ}, {"!=", T_NE}, {"ne", T_NE},
      {"=~", T_REGEX}, {"&&", T_AND}, {"||", T_OR}, {"!", T_NOT}, {"not", T_NOT}, {"->", T_ARROW},
      {"all", T_ALL}, {"hydrogen", T_HYD}, {"backbone", T_BB},
      {"name", T_SKEY}, {"resname", T_SKEY}, {"segid", T_SKEY}, {"segname", T_SKEY}, {"chainid", T_SKEY},
      {"id", T_NKEY}, {"resid", T_NKEY}, {"index", T_NKEY}};
  std::vector<Token> out;
  size_t p = 0;
  while (p < in.size()) {
    const char c = in[p];
    if (c == '\n' || c == '+') { ++p; continue; }
    if (c == ' ' || c == '\t') { ++p; continue; }
    if (c == '#' && p + 1 < in.size() && in[p + 1] != '\n') {  // "#.+"
      while (p < in.size() && in[p] != '\n') ++p;
      continue;
    }
    if (isdigit((unsigned char)c)) {
      size_t q = p;
      while (q < in.size() && isdigit((unsigned char)in[q])) ++q;
      Token t; t.t = T_NUM; t.s = in.substr(p, q - p); t.n = atoi(t.s.c_str());
      out.push_back(t); p = q; continue;
    }
    size_t best = 0; Tok bt = T_CHAR;
    for (const auto& f : fixed) {
      const size_t l = strlen(f.lit);
      if (l > best && in.compare(p, l, f.lit) == 0) { best = l; bt = f.t; }
    }
    if (best) {
      Token t; t.t = bt; t.s = in.substr(p, best); t.n = 0;
      out.push_back(t); p += best; continue;
    }
    if (c == '"' || c == '\'') {
      std::string text; size_t q = p + 1; bool closed = false;
      while (q < in.size()) {
        if (in[q] == '\n' || in[q] == c) { closed = true; ++q; break; }
        text += in[q++];
      }
      if (!closed) { Token t; t.t = T_END; t.n = 0; out.push_back(t); return out; }  // EOF inside a string
      Token t; t.t = T_STR; t.s = text; t.n = 0; out.push_back(t); p = q; continue;
    }
    Token t; t.t = c == '(' ? T_LP : (c == ')' ? T_RP : T_CHAR); t.s = std::string(1, c); t.n = 0;
    out.push_back(t); ++p;
  }
  Token e; e.t = T_END; e.n = 0; out.push_back(e);
  return out;
}

enum Op { O_PUSH_INT, O_PUSH_STR, O_NAME, O_RESNAME, O_SEGID, O_CHAIN, O_ID, O_RESID, O_INDEX, O_LT, O_LE, O_GE,
          O_GT, O_EQ, O_NOT, O_AND, O_OR, O_REGEX, O_EXTRACT, O_TRUE, O_HYD, O_BB };
struct Ins { Op op; long n; std::string s; std::shared_ptr<std::regex> re; };
typedef std::vector<Ins> Program;

struct Parser {  // src/grammar.yy:80-148; actions are emitted in bison's reduction order
  const std::vector<Token>& tk;
  size_t p = 0;
  explicit Parser(const std::vector<Token>& t) : tk(t) {}
  const Token& cur() const { return tk[p]; }
  [[noreturn]] void fail(const std::string& why) const { throw Error("Bad selection syntax - " + why); }

  static std::shared_ptr<std::regex> mkre(const std::string& s) {
    try {  // boost::regex::perl | icase (src/KernelActions.hpp:190,222)
      return std::make_shared<std::regex>(s, std::regex::ECMAScript | std::regex::icase);
    } catch (const std::regex_error& e) { throw Error(std::string("Bad regular expression '") + s + "': " + e.what()); }
  }

  void pushKey(const Token& t, Program& out) {
    Ins i; i.n = 0;
    if (t.t == T_SKEY) {
      if (t.s == "name") i.op = O_NAME; else if (t.s == "resname") i.op = O_RESNAME;
      else if (t.s == "segid" || t.s == "segname") i.op = O_SEGID; else i.op = O_CHAIN;
    } else {
      if (t.s == "id") i.op = O_ID; else if (t.s == "resid") i.op = O_RESID; else i.op = O_INDEX;
    }
    out.push_back(i);
  }

  // value : '(' value ')' | NUMBER | NKEY | STRING | SKEY | SKEY "->" STRING
  bool value(Program& out) {
    const Token& t = cur();
    if (t.t == T_LP) {
      const size_t save = p; const size_t osz = out.size();
      ++p;
      if (value(out) && cur().t == T_RP) { ++p; return true; }
      p = save; out.resize(osz);
      return false;
    }
    if (t.t == T_NUM) { Ins i; i.op = O_PUSH_INT; i.n = t.n; out.push_back(i); ++p; return true; }
    if (t.t == T_NKEY) { pushKey(t, out); ++p; return true; }
    if (t.t == T_STR) { Ins i; i.op = O_PUSH_STR; i.n = 0; i.s = t.s; out.push_back(i); ++p; return true; }
    if (t.t == T_SKEY) {
      pushKey(t, out); ++p;
      if (cur().t == T_ARROW) {
        ++p;
        if (cur().t != T_STR) fail("expected a pattern after ->");
        Ins i; i.op = O_EXTRACT; i.n = 0; i.s = cur().s; i.re = mkre(cur().s); out.push_back(i); ++p;
      }
      return true;
    }
    return false;
  }

  void rexpr(Program& out) {
    const Token& t = cur();
    if (t.t == T_NOT) { ++p; rexpr(out); Ins i; i.op = O_NOT; i.n = 0; out.push_back(i); return; }
    if (t.t == T_ALL) { ++p; Ins i; i.op = O_TRUE; i.n = 0; out.push_back(i); return; }
    if (t.t == T_HYD) { ++p; Ins i; i.op = O_HYD; i.n = 0; out.push_back(i); return; }
    if (t.t == T_BB) { ++p; Ins i; i.op = O_BB; i.n = 0; out.push_back(i); return; }
    // value OP value  |  alphid =~ strval   (tried before '(' expr ')': a parenthesis may open either)
    const size_t save = p; const size_t osz = out.size();
    const bool bare_skey = t.t == T_SKEY;
    if (value(out)) {
      const Tok op = cur().t;
      if (op == T_REGEX) {
        if (!bare_skey || tk[save + 1].t != T_REGEX) fail("=~ needs name|resname|segid|chainid on its left");
        ++p;
        if (cur().t != T_STR) fail("expected a pattern after =~");
        Ins i; i.op = O_REGEX; i.n = 0; i.s = cur().s; i.re = mkre(cur().s); out.push_back(i); ++p;
        return;
      }
      if (op == T_LT || op == T_LE || op == T_GE || op == T_GT || op == T_EQ || op == T_NE) {
        ++p;
        if (!value(out)) fail("expected a value after the comparison operator");
        Ins i; i.n = 0;
        i.op = op == T_LT ? O_LT : op == T_LE ? O_LE : op == T_GE ? O_GE : op == T_GT ? O_GT : O_EQ;
        out.push_back(i);
        if (op == T_NE) { Ins j; j.op = O_NOT; j.n = 0; out.push_back(j); }
        return;
      }
      p = save; out.resize(osz);
    }
    if (cur().t == T_LP) {
      ++p; expr(out);
      if (cur().t != T_RP) fail("missing )");
      ++p; return;
    }
    fail("unexpected '" + cur().s + "'");
  }

  void expr(Program& out) {  // left-associative, && and || at the SAME precedence (grammar.yy:70)
    rexpr(out);
    while (cur().t == T_AND || cur().t == T_OR) {
      const Tok op = cur().t; ++p;
      rexpr(out);
      Ins i; i.op = op == T_AND ? O_AND : O_OR; i.n = 0; out.push_back(i);
    }
  }
};

struct Value { bool is_int; long i; std::string s; };

// src/KernelValue.cpp:31-67
int compareValues(const Value& x, const Value& y) {
  if (x.is_int != y.is_int) throw Error("Comparing values with different types.");
  if (!x.is_int) return x.s == y.s ? 0 : (x.s < y.s ? -1 : 1);
  return (int)(x.i - y.i);
}

const char* kBackboneRes[48] = {"A", "A3", "A5", "ALA", "ARG", "ASN", "ASP", "C", "C3", "C5", "CYS", "CYX", "DA", "DC",
                                "DG", "DT", "G", "G3", "G5", "GLN", "GLU", "GLY", "HID", "HIE", "HIP", "HIS", "HSD",
                                "HSE", "HSP", "ILE", "LEU", "LYS", "MET", "MSE", "PHE", "PRO", "PTR", "SER", "T", "T3",
                                "T5", "THR", "TRP", "TYR", "U", "U3", "U5", "VAL"};  // src/Selectors.cpp:37-86
const char* kBackboneAtoms[33] = {"C", "C1'", "C2'", "C3'", "C4'", "C5'", "CA", "H", "H1'", "H12'", "H15'", "H2'",
                                  "H2''", "H22'", "H25'", "H3'", "H4'", "H5'", "H5''", "HO2'", "HO3'", "HO5'", "N",
                                  "O", "O2'", "O3'", "O4'", "O5'", "OP1", "OP2", "OP3", "OXT", "P"};  // :88-122

bool isBackbone(const Atom& a) {  // src/Selectors.cpp:126-132 (binary_search over the tables as listed)
  static const std::vector<std::string> res(kBackboneRes, kBackboneRes + 48), atm(kBackboneAtoms, kBackboneAtoms + 33);
  return std::binary_search(res.begin(), res.end(), a.resname) && std::binary_search(atm.begin(), atm.end(), a.name);
}

bool extractInt(const std::string& s, long* out) {  // std::stringstream(s) >> int
  std::istringstream iss(s);
  int v;
  if (iss >> v) { *out = v; return true; }
  return false;
}

bool run(const Program& prog, const Atom& a) {  // src/KernelActions.cpp:33-297, Selectors.cpp:189-200
  std::vector<Value> st;
  auto pop = [&]() { if (st.empty()) throw Error("Execution error - stack underflow"); Value v = st.back(); st.pop_back(); return v; };
  auto pushi = [&](long v) { Value x; x.is_int = true; x.i = v; st.push_back(x); };
  auto pushs = [&](const std::string& v) { Value x; x.is_int = false; x.i = 0; x.s = v; st.push_back(x); };
  for (const Ins& in : prog) {
    switch (in.op) {
      case O_PUSH_INT: pushi(in.n); break;
      case O_PUSH_STR: pushs(in.s); break;
      case O_NAME: pushs(a.name); break;
      case O_RESNAME: pushs(a.resname); break;
      case O_SEGID: pushs(a.segid); break;
      case O_CHAIN: pushs(a.chain); break;
      case O_ID: pushi(a.id); break;
      case O_RESID: pushi(a.resid); break;
      case O_INDEX: pushi((long)a.index); break;
      case O_LT: case O_LE: {
        if (st.size() < 2) throw Error("Execution error - stack underflow");
        const Value& v1 = st[st.size() - 1]; const Value& v2 = st[st.size() - 2];
        if ((v1.is_int && v1.i < 0) || (v2.is_int && v2.i < 0)) { st.pop_back(); st.pop_back(); pushi(0); break; }
        Value b = pop(), c = pop(); const int e = compareValues(c, b);
        pushi(in.op == O_LT ? e < 0 : e <= 0); break;
      }
      case O_GE: case O_GT: case O_EQ: {
        Value b = pop(), c = pop(); const int e = compareValues(c, b);
        pushi(in.op == O_GE ? e >= 0 : in.op == O_GT ? e > 0 : e == 0); break;
      }
      case O_NOT: { Value v = pop(); if (!v.is_int) throw Error("Invalid operand to logicalNot"); pushi(!v.i); break; }
      case O_AND: case O_OR: {
        Value b = pop(), c = pop();
        if (!b.is_int || !c.is_int) throw Error(in.op == O_AND ? "Invalid operands to logicalAnd" : "Invalid operands to logicalOr");
        pushi(in.op == O_AND ? (c.i && b.i) : (c.i || b.i)); break;
      }
      case O_REGEX: {
        Value v = pop();
        if (v.is_int) throw Error("Expected a string value but got something else...");
        pushi(std::regex_search(v.s, *in.re) ? 1 : 0); break;
      }
      case O_EXTRACT: {
        Value v = pop();
        if (v.is_int) throw Error("Expected a string value but got something else...");
        long r = -1; std::smatch what;
        if (std::regex_search(v.s, what, *in.re))
          for (size_t k = 0; k < what.size(); ++k)
            if (extractInt(what[k].str(), &r)) break;
        pushi(r); break;
      }
      case O_TRUE: pushi(1); break;
      case O_HYD: pushi(!a.name.empty() && a.name[0] == 'H'); break;  // PDB models carry no mass (KernelActions.cpp:270-284)
      case O_BB: pushi(isBackbone(a)); break;
    }
  }
  if (st.size() != 1) throw Error("Execution error - unexpected values on stack");
  if (!st.back().is_int) throw Error("Execution error - unexpected value on top of stack");
  return st.back().i != 0;
}

}  // namespace

std::vector<uint32_t> selectAtoms(const Model& model, const std::string& selection) {
  const std::vector<Token> toks = lex(selection);
  Parser ps(toks);
  Program prog;
  ps.expr(prog);
  if (ps.cur().t != T_END) throw Error("Bad selection syntax - unexpected '" + ps.cur().s + "'");
  std::vector<uint32_t> out;
  for (size_t i = 0; i < model.size(); ++i)
    if (run(prog, model[i])) out.push_back((uint32_t)i);
  return out;
}

std::vector<uint32_t> selectionToFrameIndices(const Model& model, const std::vector<uint32_t>& picked) {
  std::vector<uint32_t> out(picked.size());
  for (size_t i = 0; i < picked.size(); ++i) out[i] = model[picked[i]].index;
  return out;
}

// ============================================================== frame ranges

namespace {
struct RangeLex {  // Boost.Spirit qi with an ascii::space skipper: blanks are skipped before every terminal
  const std::string& s; size_t p;
  void skip() { while (p < s.size() && isspace((unsigned char)s[p])) ++p; }
  bool uint_(uint32_t* v) {
    skip(); size_t q = p; unsigned long long x = 0;
    while (q < s.size() && isdigit((unsigned char)s[q])) { x = x * 10 + (s[q] - '0'); if (x > 0xffffffffULL) return false; ++q; }
    if (q == p) return false;
    *v = (uint32_t)x; p = q; return true;
  }
  bool int_(int* v) {
    skip(); size_t q = p; bool neg = false;
    if (q < s.size() && (s[q] == '+' || s[q] == '-')) { neg = s[q] == '-'; ++q; }
    size_t d = q; long long x = 0;
    while (q < s.size() && isdigit((unsigned char)s[q])) { x = x * 10 + (s[q] - '0'); if (x > 0x80000000LL) return false; ++q; }
    if (q == d) return false;
    *v = (int)(neg ? -x : x); p = q; return true;
  }
  bool lit(char c) { skip(); if (p < s.size() && s[p] == c) { ++p; return true; } return false; }
};
struct RangeItem { uint32_t start, stop; int step; };
}  // namespace

std::vector<uint32_t> parseIndexRange(const std::string& spec, uint32_t maxsize) {
  RangeLex L{spec, 0};
  std::vector<RangeItem> items;
  auto endpoint = [&](uint32_t* v) { if (!L.uint_(v)) *v = maxsize; return true; };  // uint_ | eps[maxsize]
  auto ranger = [&](RangeItem* it) {  // ordered alternatives, src/index_range_parser.cpp:126-130
    const size_t save = L.p; uint32_t a, b; int st;
    if (L.uint_(&a) && L.lit(':') && L.int_(&st) && L.lit(':') && endpoint(&b)) { *it = {a, b, st}; return true; }
    L.p = save;
    if (endpoint(&a) && L.lit(':') && L.int_(&st) && L.lit(':') && L.uint_(&b)) { *it = {a, b, st}; return true; }
    L.p = save;
    if (L.uint_(&a) && L.lit(':') && endpoint(&b)) { *it = {a, b, 1}; return true; }
    L.p = save;
    if (endpoint(&a) && L.lit(':') && L.uint_(&b)) { *it = {a, b, -1}; return true; }
    L.p = save;
    if (L.uint_(&a)) { *it = {a, a, 0}; return true; }
    L.p = save;
    return false;
  };
  RangeItem it;
  bool ok = ranger(&it);
  if (ok) {
    items.push_back(it);
    for (;;) {  // ranger % ','
      const size_t save = L.p;
      if (!L.lit(',')) break;
      if (!ranger(&it)) { L.p = save; break; }
      items.push_back(it);
    }
  }
  L.skip();
  if (!ok || L.p != spec.size()) throw Error("Could not parse range: " + spec);
  std::vector<uint32_t> out;
  for (const RangeItem& r : items) {  // RangeItem::generate, :79-96
    if ((r.start < r.stop && r.step < 0) || (r.start > r.stop && r.step > 0)) throw Error("Invalid range");
    if (r.step == 0) out.push_back(r.start);
    else if (r.step > 0) { for (uint32_t i = r.start; i <= r.stop; i += (uint32_t)r.step) { out.push_back(i); if (i + (uint32_t)r.step < i) break; } }
    else { for (uint32_t i = r.start; i >= r.stop && i <= r.start; i += (uint32_t)r.step) out.push_back(i); }
  }
  return out;
}

std::vector<uint32_t> assignTrajectoryFrames(uint32_t nframes, const std::string& spec, uint32_t skip, uint32_t stride) {
  std::vector<uint32_t> frames;
  if (spec.empty()) { for (uint32_t i = skip; i < nframes; i += stride) frames.push_back(i); }
  else frames = parseIndexRange(spec, nframes - 1);
  return frames;
}

std::vector<uint32_t> uniquify(const std::vector<uint32_t>& v) {
  std::set<uint32_t> s(v.begin(), v.end());
  return std::vector<uint32_t>(s.begin(), s.end());
}

// ======================================================================= DCD

static uint32_t swab32(uint32_t x) { return __builtin_bswap32(x); }

// ---- positional file reads ---------------------------------------------------------------------
PosFile::~PosFile() { if (fd_ >= 0) ::close(fd_); }
void PosFile::open(const std::string& path) {
  fd_ = ::open(path.c_str(), O_RDONLY);
  struct stat st;
  if (fd_ < 0 || fstat(fd_, &st) != 0) throw Error("Cannot open " + path);
  size_ = (uint64_t)st.st_size;
}
void PosFile::readAt(uint64_t off, void* dst, size_t n, const std::string& what) const {
  char* p = (char*)dst;
  while (n > 0) {
    const ssize_t got = ::pread(fd_, p, n, (off_t)off);
    if (got < 0 && errno == EINTR) continue;
    if (got <= 0) throw Error(what);
    p += got; off += (uint64_t)got; n -= (size_t)got;
  }
}

namespace {
// per-thread staging area for the raw bytes of one frame
std::vector<unsigned char>& frameScratch(size_t n) {
  static thread_local std::vector<unsigned char> buf;
  if (buf.size() < n) buf.resize(n);
  return buf;
}
}  // namespace

uint32_t DCD::recordLen() {
  uint32_t n = 0;
  ifs_.read((char*)&n, 4);
  if (ifs_.eof()) return 0;
  if (ifs_.fail()) throw Error("Error while reading F77 record from DCD " + path_);
  return swab_ ? swab32(n) : n;
}

DCD::DCD(const std::string& path) : path_(path), ifs_(path.c_str(), std::ios::binary) {
  if (!ifs_) throw Error("Cannot open " + path);
  pf_.open(path);
  uint32_t datum = 0;
  ifs_.read((char*)&datum, 4);
  if (ifs_.eof() || ifs_.fail()) throw Error("Unable to read first datum from DCD file " + path);
  ifs_.seekg(0);
  if (datum != 0x54) {  // src/dcd.cpp:109-125
    if (swab32(datum) != 0x54) throw Error("Unable to determine endian-ness of DCD file " + path);
    swab_ = true;
  }
  uint32_t len = recordLen();
  if (len != 84) throw Error("Error while reading DCD header of " + path);
  unsigned char hdr[84];
  ifs_.read((char*)hdr, 84);
  if (recordLen() != 84) throw Error("Mismatch in record length while reading from DCD " + path);
  if (memcmp(hdr, "CORD", 4) != 0) throw Error("DCD is missing CORD magic marker: " + path);
  for (int i = 0; i < 20; ++i) {
    uint32_t w; memcpy(&w, hdr + 4 * (i + 1), 4);
    icntrl_[i] = (int)(swab_ ? swab32(w) : w);
  }
  { uint32_t w; memcpy(&w, hdr + 40, 4); if (swab_) w = swab32(w); memcpy(&delta_, &w, 4); }
  nframes_ = (uint32_t)icntrl_[0];
  if (icntrl_[8] != 0) throw Error("Fixed atoms not yet supported by LOOS DCD reader");
  len = recordLen();  // TITLE
  if (len == 0) throw Error("Unexpected EOF reading DCD TITLE of " + path);
  ifs_.seekg(len, std::ios::cur);
  if (recordLen() != len) throw Error("Mismatch in record length while reading from DCD " + path);
  len = recordLen();  // NATOMS
  if (len != 4) throw Error("Error reading number of atoms from DCD " + path);
  uint32_t na = 0; ifs_.read((char*)&na, 4);
  natoms_ = swab_ ? swab32(na) : na;
  if (recordLen() != 4) throw Error("Mismatch in record length while reading from DCD " + path);
  first_frame_pos_ = (uint64_t)ifs_.tellg();
  frame_size_ = 12ull * (2 + natoms_) + (hasCrystal() ? 56 : 0);  // src/dcd.cpp:246-248
  if (nframes_ == 0) {  // header without a frame count: derive it from the file size (:254-260,156-164)
    ifs_.seekg(0, std::ios::end);
    const uint64_t datasize = (uint64_t)ifs_.tellg() - first_frame_pos_;
    nframes_ = (datasize % frame_size_) ? 0 : (uint32_t)(datasize / frame_size_);
    if (nframes_ == 0) std::cerr << "Warning- DCD '" << path << "' appears empty and could not determine size\n";
  }
}

void DCD::readFramePlanes(uint32_t i, float* out) {  // src/dcd.cpp:329-371
  if (i >= nframes_) throw Error("Requested DCD frame is out of range");
  std::vector<unsigned char>& raw = frameScratch(frame_size_);
  pf_.readAt(first_frame_pos_ + (uint64_t)i * frame_size_, raw.data(), frame_size_, "Error reading data record from DCD");
  const unsigned char* p = raw.data();
  auto marker = [&]() {  // F77 record length word
    uint32_t n;
    memcpy(&n, p, 4);
    p += 4;
    return swab_ ? swab32(n) : n;
  };
  if (hasCrystal()) {
    if (marker() != 48) throw Error("Cannot read crystal parameters");
    p += 48;
    if (marker() != 48) throw Error("Mismatch in record length while reading from DCD");
  }
  for (int k = 0; k < 3; ++k) {
    const uint32_t len = marker();
    if (len != 4 * natoms_) throw Error("Size of coords stored in frame does not match model size");
    float* dst = out + (size_t)k * natoms_;
    memcpy(dst, p, len);
    p += len;
    if (marker() != len) throw Error("Mismatch in record length while reading from DCD");
    if (swab_) {
      uint32_t* w = (uint32_t*)dst;
      for (uint32_t a = 0; a < natoms_; ++a) w[a] = swab32(w[a]);
    }
  }
}

// ================================================================= MultiTraj

// =================================================================== Amber NetCDF

namespace {

struct NcCursor {  // big-endian NetCDF-3 header reader
  std::ifstream& f;
  const std::string& name;
  uint32_t u32() {
    unsigned char b[4];
    f.read((char*)b, 4);
    if (!f) throw Error("Error- " + name + ": truncated NetCDF header");
    return ((uint32_t)b[0] << 24) | ((uint32_t)b[1] << 16) | ((uint32_t)b[2] << 8) | b[3];
  }
  uint64_t u64() { const uint64_t h = u32(); return (h << 32) | u32(); }
  std::string str() {
    const uint32_t n = u32();
    if (n > (1u << 20)) throw Error("Error- " + name + ": implausible name length in NetCDF header");
    std::string s(n, '\0');
    f.read(&s[0], n);
    f.ignore((4 - n % 4) % 4);
    if (!f) throw Error("Error- " + name + ": truncated NetCDF header");
    return s;
  }
};

size_t ncTypeSize(uint32_t t, const std::string& name) {
  switch (t) {
    case 1: case 2: return 1;  // NC_BYTE, NC_CHAR
    case 3: return 2;          // NC_SHORT
    case 4: case 5: return 4;  // NC_INT, NC_FLOAT
    case 6: return 8;          // NC_DOUBLE
  }
  throw Error("Error- " + name + ": unknown NetCDF type " + std::to_string(t));
}

// att_list = ABSENT | NC_ATTRIBUTE nelems [name type nelems values...]; text attributes are collected
void ncAttributes(NcCursor& c, std::map<std::string, std::string>* text) {
  const uint32_t tag = c.u32(), n = c.u32();
  if (tag == 0 && n == 0) return;
  if (tag != 0x0C) throw Error("Error- " + c.name + ": malformed NetCDF attribute list");
  for (uint32_t i = 0; i < n; ++i) {
    const std::string an = c.str();
    const uint32_t type = c.u32(), ne = c.u32();
    const size_t bytes = (size_t)ne * ncTypeSize(type, c.name);
    std::string v(bytes, '\0');
    c.f.read(&v[0], (std::streamsize)bytes);
    c.f.ignore((4 - bytes % 4) % 4);
    if (!c.f) throw Error("Error- " + c.name + ": truncated NetCDF header");
    if (text && type == 2) {
      while (!v.empty() && v.back() == '\0') v.pop_back();
      (*text)[an] = v;
    }
  }
}

}  // namespace

bool AmberNetcdf::isNetcdf(const std::string& path) {
  std::ifstream f(path.c_str(), std::ios::binary);
  char m[4] = {0, 0, 0, 0};
  f.read(m, 4);
  return f && m[0] == 'C' && m[1] == 'D' && m[2] == 'F';
}

AmberNetcdf::AmberNetcdf(const std::string& path, uint32_t natoms_expected)
    : path_(path), ifs_(path.c_str(), std::ios::binary) {
  if (!ifs_) throw Error("Error- cannot open " + path);
  char magic[4];
  ifs_.read(magic, 4);
  if (!ifs_) throw Error("Error- " + path + ": not a NetCDF file");
  if (magic[0] == '\x89' && magic[1] == 'H' && magic[2] == 'D' && magic[3] == 'F')
    throw Error("Error- " + path + ": NetCDF-4 (HDF5) container is not supported, only NetCDF-3 classic / 64-bit offset");
  if (!(magic[0] == 'C' && magic[1] == 'D' && magic[2] == 'F' && (magic[3] == 1 || magic[3] == 2)))
    throw Error("Error- " + path + ": not a NetCDF-3 file");
  const bool off64 = magic[3] == 2;
  NcCursor c{ifs_, path_};
  uint32_t numrecs = c.u32();

  // dim_list
  std::vector<std::pair<std::string, uint32_t>> dims;
  {
    const uint32_t tag = c.u32(), n = c.u32();
    if (!(tag == 0 && n == 0)) {
      if (tag != 0x0A) throw Error("Error- " + path + ": malformed NetCDF dimension list");
      for (uint32_t i = 0; i < n; ++i) { std::string dn = c.str(); const uint32_t len = c.u32(); dims.push_back(std::make_pair(dn, len)); }
    }
  }
  std::map<std::string, std::string> gatt;
  ncAttributes(c, &gatt);
  // AmberNetcdf::init, src/amber_netcdf.cpp:28-38
  if (gatt["Conventions"].empty() || gatt["ConventionVersion"].empty())
    throw Error("Error- " + path + ": Unable to find convention global attributes.  Is this really an Amber NetCDF trajectory?");
  if (gatt["Conventions"].find("AMBER") == std::string::npos)
    throw Error("Error- " + path + ": Cannot find AMBER tag in global attribues.  Is this really an Amber NetCDF trajectory?");
  if (gatt["ConventionVersion"] != "1.0")
    throw Error("Error- " + path + ": Convention version is '" + gatt["ConventionVersion"] + "', but only 1.0 is supported for Amber NetCDF trajectories.");

  int atom_dim = -1, frame_dim = -1, spatial_dim = -1;
  for (size_t i = 0; i < dims.size(); ++i) {
    if (dims[i].first == "atom") atom_dim = (int)i;
    if (dims[i].first == "frame") frame_dim = (int)i;
    if (dims[i].first == "spatial") spatial_dim = (int)i;
  }
  if (atom_dim < 0) throw Error("Error- " + path + ": Cannot read atom id");
  natoms_ = dims[atom_dim].second;
  if (natoms_ != natoms_expected) throw Error("Error- " + path + ": AmberNetcdf has different number of atoms than expected");
  if (frame_dim < 0) throw Error("Error- " + path + ": Cannot read frame information");

  // var_list: every record variable adds its (padded) size to the record
  bool have_coords = false;
  uint64_t rec_size = 0, coord_vsize = 0;
  int nrecvars = 0;
  {
    const uint32_t tag = c.u32(), n = c.u32();
    if (!(tag == 0 && n == 0)) {
      if (tag != 0x0B) throw Error("Error- " + path + ": malformed NetCDF variable list");
      for (uint32_t i = 0; i < n; ++i) {
        const std::string vn = c.str();
        const uint32_t nd = c.u32();
        std::vector<uint32_t> vd(nd);
        for (uint32_t k = 0; k < nd; ++k) { vd[k] = c.u32(); if (vd[k] >= dims.size()) throw Error("Error- " + path + ": bad dimension id"); }
        ncAttributes(c, nullptr);
        const uint32_t type = c.u32();
        const uint32_t vsize = c.u32();
        const uint64_t begin = off64 ? c.u64() : c.u32();
        const bool is_rec = nd > 0 && dims[vd[0]].second == 0;
        if (is_rec) { rec_size += vsize; ++nrecvars; }
        if (vn == "cell_lengths") periodic_ = true;
        if (vn == "coordinates") {
          if (!(nd == 3 && (int)vd[0] == frame_dim && (int)vd[1] == atom_dim && (spatial_dim < 0 || (int)vd[2] == spatial_dim) &&
                dims[vd[2]].second == 3))
            throw Error("Error- " + path + ": coordinates variable is not (frame, atom, spatial)");
          if (type != 5 && type != 6) throw Error("Error- " + path + ": coordinates are neither float nor double");
          coords_double_ = type == 6;
          coord_begin_ = begin;
          coord_vsize = (uint64_t)natoms_ * 3 * (coords_double_ ? 8 : 4);
          have_coords = true;
          if (!is_rec) throw Error("Error- " + path + ": coordinates do not vary along the record dimension");
        }
      }
    }
  }
  if (!have_coords) throw Error("Error- " + path + ": Cannot get id for coordinates");
  // a single record variable is stored without padding between records
  rec_size_ = nrecvars == 1 ? coord_vsize : rec_size;
  if (numrecs == 0xFFFFFFFFu) {  // "streaming": the count was never finalised, take it from the file size
    ifs_.seekg(0, std::ios::end);
    const uint64_t sz = (uint64_t)ifs_.tellg();
    numrecs = rec_size_ && sz > coord_begin_ ? (uint32_t)((sz - coord_begin_ + rec_size_ - coord_vsize) / rec_size_) : 0;
  }
  nframes_ = numrecs;  // length of the unlimited dimension "frame"
  if (dims[frame_dim].second != 0) nframes_ = dims[frame_dim].second;  // fixed-size frame dimension
  coord_vsize_ = coord_vsize;
  pf_.open(path);
}

void AmberNetcdf::readFramePlanes(uint32_t i, float* out) {  // readRawFrame + updateGroupCoords, :108-150
  if (i >= nframes_) throw Error("Error- " + path_ + ": Cannot read Amber netcdf frame (coords)");
  std::vector<unsigned char>& raw = frameScratch(coord_vsize_);
  pf_.readAt(coord_begin_ + (uint64_t)i * rec_size_, raw.data(), coord_vsize_, "Error- " + path_ + ": Cannot read Amber netcdf frame (coords)");
  const unsigned char* p = raw.data();
  const size_t n = natoms_;
  if (!coords_double_) {
    for (size_t k = 0; k < n; ++k)
      for (int ax = 0; ax < 3; ++ax, p += 4) {
        const uint32_t b = ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3];
        memcpy(out + ax * n + k, &b, 4);
      }
  } else {  // nc_get_vara_float converts doubles to float (GCoord is double in the reference; the
            // staged planes are float32 like every other trajectory format)
    for (size_t k = 0; k < n; ++k)
      for (int ax = 0; ax < 3; ++ax, p += 8) {
        uint64_t b = 0;
        for (int q = 0; q < 8; ++q) b = (b << 8) | p[q];
        double d;
        memcpy(&d, &b, 8);
        out[ax * n + k] = (float)d;
      }
  }
}

// =================================================================== GROMACS XTC

namespace {

// the xdrfile table of "small" ranges: index smallidx selects how many values a small difference may take
const int kXtcMagicInts[] = {
    0, 0, 0, 0, 0, 0, 0, 0, 0, 8, 10, 12, 16, 20, 25, 32, 40, 50, 64, 80, 101, 128, 161, 203, 256, 322, 406, 512,
    645, 812, 1024, 1290, 1625, 2048, 2580, 3250, 4096, 5060, 6501, 8192, 10321, 13003, 16384, 20642, 26007,
    32768, 41285, 52015, 65536, 82570, 104031, 131072, 165140, 208063, 262144, 330280, 416127, 524287, 660561,
    832255, 1048576, 1321122, 1664510, 2097152, 2642245, 3329021, 4194304, 5284491, 6658042, 8388607, 10568983,
    13316085, 16777216};
const int kXtcFirstIdx = 9;
const int kXtcMagic = 1995;

struct XdrIn {  // big-endian 4-byte items
  std::ifstream& f;
  const std::string& name;
  bool u32(uint32_t* v) {
    unsigned char b[4];
    f.read((char*)b, 4);
    if (!f) return false;
    *v = ((uint32_t)b[0] << 24) | ((uint32_t)b[1] << 16) | ((uint32_t)b[2] << 8) | b[3];
    return true;
  }
  bool i32(int* v) { uint32_t u; if (!u32(&u)) return false; *v = (int)u; return true; }
  bool f32(float* v) { uint32_t u; if (!u32(&u)) return false; memcpy(v, &u, 4); return true; }
};

struct BitIn {  // most-significant-bit-first reader over the compressed block
  const unsigned char* p;
  size_t n, pos = 0;  // pos in bits
  // the 64 bits that start at byte `byte`, first byte most significant; zeros past the end
  uint64_t window(size_t byte) const {
    if (byte + 8 <= n) {
      uint64_t v;
      memcpy(&v, p + byte, 8);
      return __builtin_bswap64(v);
    }
    uint64_t v = 0;
    for (size_t k = 0; k < 8; ++k) v = (v << 8) | (byte + k < n ? p[byte + k] : 0u);
    return v;
  }
  uint64_t take56(unsigned nbits) {  // 1 <= nbits <= 56: with the bit offset (<= 7) still inside one window
    const uint64_t w = window(pos >> 3);
    const uint64_t v = (w << (pos & 7)) >> (64 - nbits);
    pos += nbits;
    return v;
  }
  uint32_t take(unsigned nbits) { return nbits ? (uint32_t)take56(nbits) : 0u; }  // nbits <= 32
  // three values packed as one number in the mixed radix (sizes[1], sizes[2]), stored least significant
  // BYTE first, each byte most significant bit first (decodeints, src/xtc.cpp:132-158)
  void take3(unsigned nbits, const uint32_t sizes[3], int out[3]) {
    if (nbits >= 1 && nbits <= 56) {  // the usual case: one window, 64-bit arithmetic
      const uint64_t x = take56(nbits);  // bytes b0 b1 ... (b0 first) then the `rem` leftover bits, as one big-endian string
      const unsigned nfull = nbits >> 3, rem = nbits & 7;
      const uint64_t top = x & ((1ull << rem) - 1);  // the leftover bits are the most significant part
      const uint64_t be = x >> rem;                  // nfull bytes, b0 most significant
      uint64_t v = nfull ? __builtin_bswap64(be) >> (64 - 8 * nfull) : 0;  // b0 least significant
      v |= top << (8 * nfull);
      const uint64_t q2 = v / sizes[2];
      out[2] = (int)(uint32_t)(v - q2 * sizes[2]);
      const uint64_t q1 = q2 / sizes[1];
      out[1] = (int)(uint32_t)(q2 - q1 * sizes[1]);
      out[0] = (int)(uint32_t)q1;
      return;
    }
    unsigned __int128 v = 0;
    unsigned shift = 0;
    while (nbits > 8) { v |= (unsigned __int128)take(8) << shift; shift += 8; nbits -= 8; }
    if (nbits > 0) v |= (unsigned __int128)take(nbits) << shift;
    out[2] = (int)(uint32_t)(v % sizes[2]); v /= sizes[2];
    out[1] = (int)(uint32_t)(v % sizes[1]); v /= sizes[1];
    out[0] = (int)(uint32_t)v;
  }
};

unsigned bitLength(unsigned __int128 v) { unsigned n = 0; while (v) { ++n; v >>= 1; } return n; }

}  // namespace

XTC::XTC(const std::string& path) : path_(path), ifs_(path.c_str(), std::ios::binary) {
  if (!ifs_) throw Error("Error- cannot open " + path);
  XdrIn x{ifs_, path_};
  ifs_.seekg(0, std::ios::end);
  const uint64_t file_size = (uint64_t)ifs_.tellg();
  ifs_.seekg(0);
  // scanFrames (src/xtc.cpp:421-477): walk the frames, remember where each one starts
  for (;;) {
    const uint64_t pos = (uint64_t)ifs_.tellg();
    int magic;
    if (!x.i32(&magic)) break;  // clean end of file
    if (magic != kXtcMagic)
      throw Error("Error- " + path + ": Invalid XTC magic number (got " + std::to_string(magic) + " but expected " + std::to_string(kXtcMagic) + ")");
    uint32_t na, step;
    float time, box[9];
    bool ok = x.u32(&na) && x.u32(&step) && x.f32(&time);
    for (int k = 0; k < 9 && ok; ++k) ok = x.f32(&box[k]);
    if (!ok) throw Error("Error- " + path + ": Problem reading XTC header");
    if (natoms_ == 0) natoms_ = na;
    else if (natoms_ != na) throw Error("Error- " + path + ": XTC frames have differing numbers of atoms");
    uint32_t nbytes;
    if (natoms_ <= 9) {
      uint32_t lsize;
      if (!x.u32(&lsize) || lsize != natoms_) throw Error("Error- " + path + ": XTC small system vector size is not what was expected");
      nbytes = natoms_ * 12;
    } else {
      ifs_.seekg(9 * 4, std::ios::cur);  // natoms, precision, minint[3], maxint[3], smallidx
      if (!x.u32(&nbytes)) throw Error("Error- " + path + ": Problem scanning XTC trajectory to build frame indices");
    }
    ifs_.seekg((std::streamoff)((nbytes + 3) / 4 * 4), std::ios::cur);
    // a seek past the end succeeds on a file; a frame cut short is an error here, not a frame
    if (!ifs_ || (uint64_t)ifs_.tellg() > file_size)
      throw Error("Error- " + path + ": XTC frame " + std::to_string(frame_pos_.size()) + " is truncated");
    frame_pos_.push_back(pos);
  }
  ifs_.clear();
  if (frame_pos_.empty()) throw Error("Error- " + path + ": Unable to read in the first frame");
  // one more entry: where the last frame ends (every frame is fetched whole with one positional read)
  ifs_.seekg((std::streamoff)frame_pos_.back());
  {
    XdrIn y{ifs_, path_};
    uint32_t nbytes = 0;
    if (natoms_ <= 9) {
      nbytes = natoms_ * 12;
      frame_pos_.push_back(frame_pos_.back() + 14 * 4 + nbytes);
    } else {
      ifs_.seekg((13 + 9) * 4, std::ios::cur);
      if (!y.u32(&nbytes)) throw Error("Error- " + path + ": Problem scanning XTC trajectory to build frame indices");
      frame_pos_.push_back(frame_pos_.back() + (13 + 9 + 1) * 4 + (nbytes + 3) / 4 * 4);
    }
  }
  ifs_.clear();
  pf_.open(path);
}

void XTC::readFramePlanes(uint32_t fi, float* out) {
  if (fi + 1 >= frame_pos_.size()) throw Error("Error- " + path_ + ": Requested XTC frame is out of range");
  const size_t frame_bytes = (size_t)(frame_pos_[fi + 1] - frame_pos_[fi]);
  std::vector<unsigned char>& raw = frameScratch(frame_bytes + 8);
  pf_.readAt(frame_pos_[fi], raw.data(), frame_bytes, "Error- " + path_ + ": Problem reading XTC frame");
  size_t off = 13 * 4;  // past magic, natoms, step, time, box
  auto word = [&](uint32_t* v) {  // XDR: big-endian 32-bit words
    if (off + 4 > frame_bytes) return false;
    const unsigned char* b = raw.data() + off;
    *v = ((uint32_t)b[0] << 24) | ((uint32_t)b[1] << 16) | ((uint32_t)b[2] << 8) | b[3];
    off += 4;
    return true;
  };
  auto wordf = [&](float* f) { uint32_t u; if (!word(&u)) return false; memcpy(f, &u, 4); return true; };
  auto wordi = [&](int* i) { uint32_t u; if (!word(&u)) return false; *i = (int)u; return true; };
  uint32_t lsize;
  if (!word(&lsize) || lsize != natoms_) throw Error("Error- " + path_ + ": Problem reading XTC frame");
  const size_t n = natoms_;
  if (lsize <= 9) {  // plain floats, nm
    for (size_t k = 0; k < n; ++k)
      for (int ax = 0; ax < 3; ++ax) {
        float v;
        if (!wordf(&v)) throw Error("Error- " + path_ + ": XTC Error: number of uncompressed coords read did not match number expected");
        out[ax * n + k] = (float)((double)v * 10.0);
      }
    return;
  }
  float precision;
  int minint[3], maxint[3], smallidx;
  bool ok = wordf(&precision);
  for (int k = 0; k < 3 && ok; ++k) ok = wordi(&minint[k]);
  for (int k = 0; k < 3 && ok; ++k) ok = wordi(&maxint[k]);
  uint32_t nbytes = 0;
  ok = ok && wordi(&smallidx) && word(&nbytes);
  if (!ok || smallidx < kXtcFirstIdx || smallidx >= (int)(sizeof(kXtcMagicInts) / sizeof(int)) || off + nbytes > frame_bytes)
    throw Error("Error- " + path_ + ": Problem reading XTC frame");
  const unsigned char* bits = raw.data() + off;

  uint32_t sizeint[3], bitsizeint[3] = {0, 0, 0};
  for (int k = 0; k < 3; ++k) {  // unsigned arithmetic: a corrupt header must not overflow (or divide by zero below)
    sizeint[k] = (uint32_t)maxint[k] - (uint32_t)minint[k] + 1u;
    if (sizeint[k] == 0 || maxint[k] < minint[k]) throw Error("Error- " + path_ + ": corrupt XTC frame (bounding box)");
  }
  unsigned bitsize;
  if ((sizeint[0] | sizeint[1] | sizeint[2]) > 0xffffffu) {  // too large to pack as one number: separate fields
    for (int k = 0; k < 3; ++k) bitsizeint[k] = bitLength(sizeint[k]);
    bitsize = 0;
  } else {
    bitsize = bitLength((unsigned __int128)sizeint[0] * sizeint[1] * sizeint[2]);
  }
  int smaller = kXtcMagicInts[std::max(kXtcFirstIdx, smallidx - 1)] / 2;
  int smallnum = kXtcMagicInts[smallidx] / 2;
  uint32_t sizesmall[3] = {(uint32_t)kXtcMagicInts[smallidx], (uint32_t)kXtcMagicInts[smallidx], (uint32_t)kXtcMagicInts[smallidx]};
  const float inv_precision = 1.0f / precision;
  auto emit = [&](size_t atom, const int c[3]) {
    if (atom >= n) throw Error("Error- " + path_ + ": corrupt XTC frame (more atoms than announced)");
    for (int ax = 0; ax < 3; ++ax) out[ax * n + atom] = (float)((double)((float)c[ax] * inv_precision) * 10.0);
  };
  BitIn in{bits, nbytes};
  size_t i = 0;
  int run = 0;
  while (i < n) {
    int cur[3], prev[3];
    if (bitsize == 0) for (int k = 0; k < 3; ++k) cur[k] = (int)in.take(bitsizeint[k]);
    else in.take3(bitsize, sizeint, cur);
    for (int k = 0; k < 3; ++k) { cur[k] = (int)((uint32_t)cur[k] + (uint32_t)minint[k]); prev[k] = cur[k]; }
    int is_smaller = 0;
    if (in.take(1)) {  // the run length (and the change of the small range) is stated anew
      run = (int)in.take(5);
      is_smaller = run % 3;
      run -= is_smaller;
      --is_smaller;
    }
    if (run > 0) {
      for (int k = 0; k < run; k += 3) {
        int d[3];
        in.take3((unsigned)smallidx, sizesmall, d);
        for (int q = 0; q < 3; ++q) d[q] = (int)((uint32_t)d[q] + (uint32_t)prev[q] - (uint32_t)smallnum);
        if (k == 0) {  // the writer stored the second atom of the run first (water: O after H): swap back
          emit(i, d);
          emit(i + 1, cur);
          i += 2;
        } else {
          emit(i, d);
          ++i;
        }
        for (int q = 0; q < 3; ++q) prev[q] = d[q];  // differences chain through the small-coded atoms
      }
    } else {
      emit(i, cur);
      ++i;
    }
    smallidx += is_smaller;
    if (is_smaller < 0) {
      smallnum = smaller;
      smaller = smallidx > kXtcFirstIdx ? kXtcMagicInts[smallidx - 1] / 2 : 0;
    } else if (is_smaller > 0) {
      smaller = smallnum;
      smallnum = kXtcMagicInts[smallidx] / 2;
    }
    if (smallidx < kXtcFirstIdx || smallidx >= (int)(sizeof(kXtcMagicInts) / sizeof(int)))
      throw Error("Error- " + path_ + ": corrupt XTC frame");
    sizesmall[0] = sizesmall[1] = sizesmall[2] = (uint32_t)kXtcMagicInts[smallidx];
  }
}

// ---- Python float text ----------------------------------------------------------------------
namespace {

// repr() of a double: shortest digit string that converts back to the same double
void pyRepr(double x, std::string& out) {
  if (std::isnan(x)) { out += "nan"; return; }
  if (std::isinf(x)) { out += x < 0 ? "-inf" : "inf"; return; }
  if (x == 0.0) { out += std::signbit(x) ? "-0.0" : "0.0"; return; }
  char buf[40];
  int prec = 1;
  for (; prec <= 17; ++prec) {
    snprintf(buf, sizeof(buf), "%.*e", prec - 1, x);
    if (strtod(buf, nullptr) == x) break;
  }
  // buf = [-]d[.ddd]e[+-]XX
  const char* p = buf;
  if (*p == '-') { out += '-'; ++p; }
  std::string digits;
  for (; *p && *p != 'e'; ++p)
    if (*p != '.') digits += *p;
  const int exp10 = atoi(p + 1);
  while (digits.size() > 1 && digits.back() == '0') digits.pop_back();
  const int decpt = exp10 + 1;  // value = 0.DIGITS x 10^decpt
  if (decpt <= -4 || decpt > 16) {  // format_float_short, 'r': exponent form
    out += digits[0];
    if (digits.size() > 1) { out += '.'; out.append(digits, 1, std::string::npos); }
    char e[8];
    snprintf(e, sizeof(e), "e%+03d", decpt - 1);
    out += e;
  } else if (decpt <= 0) {
    out += "0.";
    out.append((size_t)(-decpt), '0');
    out += digits;
  } else if ((size_t)decpt >= digits.size()) {
    out += digits;
    out.append((size_t)decpt - digits.size(), '0');
    out += ".0";
  } else {
    out.append(digits, 0, (size_t)decpt);
    out += '.';
    out.append(digits, (size_t)decpt, std::string::npos);
  }
}

}  // namespace

void pyRoundRepr(double v, int ndigits, std::string& out) {
  if (!std::isfinite(v) || v == 0.0) { pyRepr(v, out); return; }
  // Fast exact path for the values a distance matrix holds: 1e-4 <= |v| < 1e15 and 0..15 decimals.
  // v = m 2^-s exactly (m < 2^53, 1 <= s <= 66 in this range), so R = round-half-even(m 10^nd / 2^s) is
  // integer arithmetic on 128 bits -- the correctly rounded decimal float.__round__ produces.  With at most
  // 15 significant digits that decimal is also the shortest text of the double it converts to (repr).
  const double av = std::fabs(v);
  if (ndigits >= 0 && ndigits <= 15 && av >= 1e-4 && av < 1e15) {
    static const uint64_t kPow10[16] = {1ull, 10ull, 100ull, 1000ull, 10000ull, 100000ull, 1000000ull, 10000000ull, 100000000ull,
                                        1000000000ull, 10000000000ull, 100000000000ull, 1000000000000ull, 10000000000000ull,
                                        100000000000000ull, 1000000000000000ull};
    uint64_t bits;
    memcpy(&bits, &av, 8);
    const int be = (int)(bits >> 52);  // normal here: av >= 1e-4
    const uint64_t m = (bits & ((1ull << 52) - 1)) | (1ull << 52);
    const int sh = 1075 - be;          // v = m * 2^-sh
    if (sh >= 1 && sh <= 70) {
      const unsigned __int128 P = (unsigned __int128)m * kPow10[ndigits];
      unsigned __int128 R = P >> sh;
      const unsigned __int128 rem = P & (((unsigned __int128)1 << sh) - 1), half = (unsigned __int128)1 << (sh - 1);
      if (rem > half || (rem == half && (R & 1))) ++R;
      if (R < (unsigned __int128)kPow10[15]) {
        const uint64_t r = (uint64_t)R;
        if (v < 0) out += '-';
        if (r == 0) { out += "0.0"; return; }
        const uint64_t ip = r / kPow10[ndigits];
        uint64_t fp = r % kPow10[ndigits];
        char tmp[24];
        int n = 0;
        uint64_t t = ip;
        do { tmp[n++] = (char)('0' + t % 10); t /= 10; } while (t);
        while (n) out += tmp[--n];
        out += '.';
        if (fp == 0) { out += '0'; return; }
        int nd = ndigits;
        while (fp % 10 == 0) { fp /= 10; --nd; }  // trailing zeros are not printed
        for (int k = nd - 1; k >= 0; --k) { tmp[k] = (char)('0' + fp % 10); fp /= 10; }
        out.append(tmp, (size_t)nd);
        return;
      }
    }
  }
  char buf[64];
  // common case: 0..15 decimals of a value of ordinary size.  glibc's %.{n}f is the correctly rounded
  // decimal (ties decided on the exact binary value, as _Py_dg_dtoa mode 3 does); with at most 15
  // significant digits it IS the shortest string of the double it converts to.
  if (ndigits >= 0 && ndigits <= 15 && std::fabs(v) < 1e15) {
    const int n = snprintf(buf, sizeof(buf), "%.*f", ndigits, v);
    int sig = 0;
    bool lead = true, allzero = true;
    for (int k = 0; k < n; ++k) {
      const char c = buf[k];
      if (c < '0' || c > '9') continue;
      if (c != '0') { lead = false; allzero = false; }
      if (!lead) ++sig;
    }
    if (allzero) { out += buf[0] == '-' ? "-0.0" : "0.0"; return; }
    const double r = strtod(buf, nullptr);
    if (sig <= 15 && std::fabs(r) >= 1e-4) {
      int end = n;
      if (ndigits > 0) {
        while (buf[end - 1] == '0') --end;
        if (buf[end - 1] == '.') ++end;  // keep one zero: "2.0"
        out.append(buf, (size_t)end);
      } else {
        out.append(buf, (size_t)n);
        out += ".0";
      }
      return;
    }
    pyRepr(r, out);
    return;
  }
  // general case (float.__round__, Objects/floatobject.c double_round): exact decimal expansion,
  // round half even at the requested position, convert back, repr
  if (ndigits > 323) { pyRepr(v, out); return; }
  if (ndigits < -308) { pyRepr(v < 0 ? -0.0 : 0.0, out); return; }
  std::vector<char> big(1500);
  const int n = snprintf(big.data(), big.size(), "%.1100f", std::fabs(v));  // exact: a double has <= 1074 fractional bits
  std::string s(big.data(), (size_t)n);
  const size_t dot = s.find('.');
  std::string ip = s.substr(0, dot), fp = s.substr(dot + 1);
  std::string all = ip + fp;                     // digits, point after ip.size()
  long keep = (long)ip.size() + ndigits;         // digits kept
  std::string res;
  if (keep < 0) res = "0";
  else {
    bool up = false;
    if ((size_t)keep < all.size()) {
      const char d = all[(size_t)keep];
      bool rest = false;
      for (size_t k = (size_t)keep + 1; k < all.size() && !rest; ++k) rest = all[k] != '0';
      if (d > '5' || (d == '5' && rest)) up = true;
      else if (d == '5') up = keep > 0 && ((all[(size_t)keep - 1] - '0') & 1);
    }
    res = all.substr(0, (size_t)keep);
    if (up) {
      long k = (long)res.size() - 1;
      for (; k >= 0; --k) {
        if (res[(size_t)k] == '9') res[(size_t)k] = '0';
        else { ++res[(size_t)k]; break; }
      }
      if (k < 0) { res.insert(res.begin(), '1'); ++keep; ip.insert(ip.begin(), '0'); }
    }
    if (res.empty()) res = "0";
  }
  // res holds the kept digits; its decimal point sits ndigits from the right (ndigits may be negative)
  std::string text = (v < 0 ? "-" : "") + res + "e" + std::to_string(-ndigits);
  pyRepr(strtod(text.c_str(), nullptr), out);
}

void writePyRoundMatrix(const double* M, size_t rows, size_t cols, size_t ld, int precision, unsigned threads,
                        const std::function<void(const char*, size_t)>& write) {
  if (rows == 0) return;
  const size_t block = std::max<size_t>(1, (1u << 20) / std::max<size_t>(1, cols));
  const unsigned nt = (unsigned)std::max<size_t>(1, std::min<size_t>(threads, (rows + block - 1) / block));
  std::vector<std::string> text(nt);
  for (size_t r0 = 0; r0 < rows; r0 += block * nt) {
    auto work = [&](unsigned t) {
      std::string& out = text[t];
      out.clear();
      const size_t lo = std::min(rows, r0 + t * block), hi = std::min(rows, lo + block);
      for (size_t i = lo; i < hi; ++i) {
        for (size_t j = 0; j < cols; ++j) { pyRoundRepr(M[j * ld + i], precision, out); out += ' '; }
        out += '\n';
      }
    };
    if (nt == 1) work(0);
    else {
      std::vector<std::thread> pool;
      for (unsigned t = 0; t < nt; ++t) pool.emplace_back(work, t);
      for (std::thread& th : pool) th.join();
    }
    for (unsigned t = 0; t < nt; ++t)
      if (!text[t].empty()) write(text[t].data(), text[t].size());
  }
}

std::shared_ptr<Trajectory> openTrajectory(const std::string& path, uint32_t natoms) {
  const size_t dot = path.rfind('.');
  std::string suffix = dot == std::string::npos ? "" : path.substr(dot + 1);
  std::transform(suffix.begin(), suffix.end(), suffix.begin(), ::tolower);
  if (suffix == "dcd") return std::make_shared<DCD>(path);
  if (suffix == "xtc") return std::make_shared<XTC>(path);
  if (suffix == "h5") return std::make_shared<MDTrajH5>(path, natoms);
  if (suffix == "nc" || suffix == "netcdf") return std::make_shared<AmberNetcdf>(path, natoms);
  if ((suffix == "crd" || suffix == "mdcrd") && AmberNetcdf::isNetcdf(path)) return std::make_shared<AmberNetcdf>(path, natoms);
  throw Error("Error- unknown or unsupported trajectory format for '" + path + "' (supported here: dcd, xtc, nc/netcdf, h5)");
}

uint32_t MultiTraj::nframes(size_t i) const {
  const uint32_t n = trajs.at(i)->nframes();
  if (n <= skip) return 0;
  return (n - skip + stride - 1) / stride;
}
uint32_t MultiTraj::nframes() const {
  uint32_t t = 0;
  for (size_t i = 0; i < trajs.size(); ++i) t += nframes(i);
  return t;
}
std::pair<uint32_t, uint32_t> MultiTraj::location(uint32_t i) const {
  uint32_t k, j;
  for (j = k = 0; k < trajs.size(); ++k) {
    const uint32_t n = nframes(k);
    if (j + n > i) break;
    j += n;
  }
  return std::make_pair(k, skip + (i - j) * stride);
}

// ==================================================================== output

std::string formatG(double v, unsigned precision) {
  char buf[64];
  snprintf(buf, sizeof(buf), "%.*g", (int)precision, v);
  return buf;
}

// ---- matrix text (src/MatrixImpl.hpp:210-222) ----
// `os << setprecision(p) << M` prints every float with the stream's default notation, i.e. printf's
// "%.{p}g", followed by one space.  At BASELINE's C4 size that is 10^10 values (45 GB of text):
// glibc's snprintf (~130 ns per value) would take 20 minutes on one core, the GPU kernel takes
// 0.45 s.  formatG_fast produces the same bytes for float inputs from exact integer arithmetic
// (a float times a power of ten fits 128 bits), ~10x faster, and writeMatrix formats 16-row blocks
// on all host threads.  Anything outside the fast path's range goes through snprintf.
namespace {

const double kPow10[23] = {1e0, 1e1, 1e2, 1e3, 1e4, 1e5, 1e6, 1e7, 1e8, 1e9, 1e10, 1e11,
                           1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};
const uint64_t kPow10u[20] = {1ull, 10ull, 100ull, 1000ull, 10000ull, 100000ull, 1000000ull, 10000000ull,
                              100000000ull, 1000000000ull, 10000000000ull, 100000000000ull, 1000000000000ull,
                              10000000000000ull, 100000000000000ull, 1000000000000000ull, 10000000000000000ull,
                              100000000000000000ull, 1000000000000000000ull, 10000000000000000000ull};

// Appends "%.{p}g" of v to out (no terminator); returns the number of characters.
// Exact for every finite float with 1 <= p <= 9 and 1e-5 <= |v| < 1e9 (and for 0); others: snprintf.
inline int formatG_fast(float v, unsigned p, char* out) {
  uint32_t bits;
  memcpy(&bits, &v, 4);
  const uint32_t mag = bits & 0x7fffffffu;
  if (mag == 0) {
    int n = 0;
    if (bits >> 31) out[n++] = '-';
    out[n++] = '0';
    return n;
  }
  const double av = std::fabs((double)v);
  if (p < 1 || p > 9 || !(av >= 1e-5) || !(av < 1e9)) return snprintf(out, 40, "%.*g", (int)(p ? p : 1), (double)v);
  // decimal exponent of the leading digit: 10^X <= av < 10^(X+1).  Exact: powers of ten up to
  // 10^22 are doubles, and a float times 10^k (k <= 5) still fits the 53-bit significand.
  int X;
  if (av >= 1.0) {
    X = 0;
    while (X < 8 && av >= kPow10[X + 1]) ++X;
  } else {
    X = -1;
    while (X > -5 && av * kPow10[-X] < 1.0) --X;
  }
  // n = round_half_even(av * 10^k), k = p - 1 - X, from the float's exact value m * 2^e
  const int e_raw = (int)(mag >> 23);
  uint64_t m = mag & 0x7fffffu;
  int e;
  if (e_raw == 0) e = -149; else { m |= 0x800000u; e = e_raw - 150; }
  const int k = (int)p - 1 - X;  // -8 .. 13
  uint64_t n;
  if (k >= 0 && k <= 11 && e < 0 && e > -64) {
    // the common case: the denominator is a power of two and everything fits 64 bits
    const uint64_t num = m * kPow10u[k];  // < 2^24 * 10^11 < 2^61
    const int sh = -e;
    const uint64_t q = num >> sh, r = num & ((1ull << sh) - 1), half = 1ull << (sh - 1);
    n = q + ((r > half || (r == half && (q & 1))) ? 1 : 0);
  } else {
    unsigned __int128 num = m, den = 1;
    if (k >= 0) num *= kPow10u[k]; else den *= kPow10u[-k];
    if (e >= 0) num <<= e; else den <<= -e;   // av >= 1e-5 bounds -e <= 40: den < 2^41 * 10^8
    unsigned __int128 q = num / den, r = num - q * den;
    const unsigned __int128 twice = r * 2;
    if (twice > den || (twice == den && (q & 1))) ++q;
    n = (uint64_t)q;
  }
  if (n >= kPow10u[p]) { n /= 10; ++X; }  // rounding carried into a new decade (n == 10^p exactly)
  // digits of n: exactly p of them
  char dig[20];
  for (int i = (int)p - 1; i >= 0; --i) { dig[i] = (char)('0' + n % 10); n /= 10; }
  int nd = (int)p;
  while (nd > 1 && dig[nd - 1] == '0') --nd;  // %g strips trailing zeros
  int len = 0;
  if (bits >> 31) out[len++] = '-';
  if (X < -4 || X >= (int)p) {  // scientific: d[.ddd]e+XX
    out[len++] = dig[0];
    if (nd > 1) { out[len++] = '.'; for (int i = 1; i < nd; ++i) out[len++] = dig[i]; }
    out[len++] = 'e';
    int ax = X;
    if (ax < 0) { out[len++] = '-'; ax = -ax; } else out[len++] = '+';
    out[len++] = (char)('0' + ax / 10);
    out[len++] = (char)('0' + ax % 10);
  } else if (X >= 0) {  // X + 1 integer digits, the rest after the point
    for (int i = 0; i <= X; ++i) out[len++] = i < nd ? dig[i] : '0';
    if (nd > X + 1) { out[len++] = '.'; for (int i = X + 1; i < nd; ++i) out[len++] = dig[i]; }
  } else {  // 0.000ddd
    out[len++] = '0'; out[len++] = '.';
    for (int i = 0; i < -X - 1; ++i) out[len++] = '0';
    for (int i = 0; i < nd; ++i) out[len++] = dig[i];
  }
  return len;
}

}  // namespace

void writeMatrix(std::ostream& os, const float* M, size_t m, size_t n, size_t ld, unsigned precision) {
  os << "# " << m << " " << n << " (0)\n";
  // Blocks of RB rows: a column-major matrix is read one cache line (16 consecutive rows of one
  // column) at a time; a batch of blocks is formatted by all threads, then written in order.
  constexpr size_t RB = 16;
  const size_t nblocks = (m + RB - 1) / RB;
  unsigned nthreads = std::thread::hardware_concurrency();
  if (nthreads == 0) nthreads = 1;
  if (const char* env = getenv("LOOSB200_WRITER_THREADS")) nthreads = (unsigned)std::max(1, atoi(env));
  if ((size_t)nthreads > nblocks) nthreads = (unsigned)std::max<size_t>(1, nblocks);
  const size_t batch = (size_t)nthreads * 4;  // blocks formatted per round
  std::vector<std::string> text(batch);
  for (size_t b0 = 0; b0 < nblocks; b0 += batch) {
    const size_t nb = std::min(batch, nblocks - b0);
    auto work = [&](unsigned t) {
      // one flat buffer per block: row r writes at rowbuf + r * cap (cap = worst-case line length),
      // values are formatted in place -- no per-value string append
      const size_t cap = n * (size_t)(precision + 9) + 2;  // sign, digits, '.', "e+XX", ' ' per value
      std::vector<char> rowbuf(RB * cap);
      char* pos[RB];
      for (size_t b = t; b < nb; b += nthreads) {
        const size_t j0 = (b0 + b) * RB, rows = std::min(RB, m - j0);
        for (size_t r = 0; r < rows; ++r) pos[r] = rowbuf.data() + r * cap;
        for (size_t i = 0; i < n; ++i) {
          const float* col = M + i * ld + j0;
          for (size_t r = 0; r < rows; ++r) {
            char* q = pos[r];
            q += formatG_fast(col[r], precision, q);
            *q++ = ' ';
            pos[r] = q;
          }
        }
        std::string& out = text[b];
        out.clear();
        size_t total = 0;
        for (size_t r = 0; r < rows; ++r) total += (size_t)(pos[r] - (rowbuf.data() + r * cap)) + 1;
        out.reserve(total);
        for (size_t r = 0; r < rows; ++r) { out.append(rowbuf.data() + r * cap, (size_t)(pos[r] - (rowbuf.data() + r * cap))); out += '\n'; }
      }
    };
    if (nthreads == 1) {
      work(0);
    } else {
      std::vector<std::thread> pool;
      for (unsigned t = 0; t < nthreads; ++t) pool.emplace_back(work, t);
      for (auto& th : pool) th.join();
    }
    for (size_t b = 0; b < nb; ++b) os.write(text[b].data(), (std::streamsize)text[b].size());
  }
  os.flush();
}

void writeNpy(const std::string& path, const float* M, size_t m, size_t n, size_t ld) {
  std::ofstream f(path.c_str(), std::ios::binary);
  if (!f) throw Error("Error- cannot open " + path + " for writing");
  std::string dict = "{'descr': '<f4', 'fortran_order': True, 'shape': (" + std::to_string(m) + ", " + std::to_string(n) + "), }";
  size_t total = 10 + dict.size() + 1;          // magic(6) + version(2) + header length(2) + dict + newline
  dict.append((64 - total % 64) % 64, ' ');     // data starts on a 64-byte boundary
  dict += '\n';
  if (dict.size() > 65535) throw Error("Error- npy header too long");
  const unsigned char head[10] = {0x93, 'N', 'U', 'M', 'P', 'Y', 1, 0, (unsigned char)(dict.size() & 0xff), (unsigned char)(dict.size() >> 8)};
  f.write((const char*)head, 10);
  f.write(dict.data(), (std::streamsize)dict.size());
  if (ld == m) {
    f.write((const char*)M, (std::streamsize)(m * n * sizeof(float)));
  } else {
    for (size_t j = 0; j < n; ++j) f.write((const char*)(M + j * ld), (std::streamsize)(m * sizeof(float)));
  }
  if (!f) throw Error("Error- write to " + path + " failed");
}

std::vector<float> readMatrix(std::istream& is, size_t* m_out, size_t* n_out) {
  std::string line;
  int m = 0, n = 0;
  while (std::getline(is, line).good())
    if (sscanf(line.c_str(), "# %d %d", &m, &n) == 2) break;
  if (m == 0 && n == 0) throw Error("Could not find magic marker in matrix file");
  if (m <= 0 || n <= 0) throw Error("Error while reading magic marker");
  // the rest of the stream in one piece: strtof per token (operator>> per value is ~10x slower)
  std::string body((std::istreambuf_iterator<char>(is)), std::istreambuf_iterator<char>());
  // every value takes at least one character and a separator: a marker that promises more than the
  // stream can hold is refused before anything of that size is allocated
  if ((uint64_t)m * (uint64_t)n > body.size() / 2 + 1)
    throw Error("Read error: the matrix file ends before the " + std::to_string(m) + " x " + std::to_string(n) + " values its marker announces");
  std::vector<float> M((size_t)m * (size_t)n);
  const char* p = body.c_str();
  const char *last = p, *last_end = p;  // the previous token (empty before the first)
  for (int j = 0; j < m; ++j)
    for (int i = 0; i < n; ++i) {
      while (*p == ' ' || *p == '\n' || *p == '\t' || *p == '\r') ++p;
      // the reference's `is >> inbuf` fails at the end of the stream without touching inbuf: its message shows the LAST token read
      if (!*p) throw Error("Read error at (" + std::to_string(j) + "," + std::to_string(i) + ") with input '" + std::string(last, last_end) + "'");
      char* end = nullptr;
      const float v = strtof(p, &end);
      if (end == p) {
        const char* q = p;
        while (*q && *q != ' ' && *q != '\n') ++q;
        throw Error("Invalid conversion on matrix read at (" + std::to_string(j) + "," + std::to_string(i) + ") of '" + std::string(p, q) + "'");
      }
      M[(size_t)i * (size_t)m + (size_t)j] = v;
      last = p;
      last_end = end;
      while (*last_end && *last_end != ' ' && *last_end != '\n' && *last_end != '\t' && *last_end != '\r') ++last_end;  // a token is whitespace-delimited
      p = end;
    }
  *m_out = (size_t)m; *n_out = (size_t)n;
  return M;
}

// test hook: the fast formatter on one value (frontend_capi.cpp)
int formatFloatG(float v, unsigned precision, char* out) { return formatG_fast(v, precision, out); }

std::string invocationHeader(int argc, char* argv[]) {  // src/utils.cpp:119-164
  time_t t = time(0);
  std::string timestamp(asctime(localtime(&t)));
  timestamp.erase(timestamp.length() - 1, 1);
  struct passwd* pwd = getpwuid(getuid());
  const std::string user = pwd ? pwd->pw_name : "UNKNOWN USER";
  char cwd[4096];
  const char* cd = getcwd(cwd, sizeof(cwd));
  std::string invoke = std::string(argv[0]) + " ";
  for (int i = 1; i < argc; ++i) invoke += "'" + std::string(argv[i]) + "'" + (i == argc - 1 ? "" : " ");
  invoke += " - " + user + " (" + timestamp + ")";
  if (cd) invoke += " {" + std::string(cd) + "}";
  invoke += " [loos-b200 0.1 BUILT ON " __DATE__ " " __TIME__ "]";
  std::string out;
  for (char c : invoke) { if (c == '\n') out += "\\n"; else out += c; }
  return out;
}

// =================================================================== options

void Options::add(const std::string& longname, char shortname, const std::string& defval, const std::string& help, bool is_bool) {
  Opt o; o.name = longname; o.shortname = shortname; o.defval = defval; o.help = help; o.is_bool = is_bool; o.set = false;
  opts_.push_back(o);
}
void Options::addPositional(const std::string& name, int max_count) { positional_.push_back(std::make_pair(name, max_count)); }
Options::Opt* Options::find(const std::string& name) { for (auto& o : opts_) if (o.name == name) return &o; return nullptr; }
const Options::Opt* Options::find(const std::string& name) const { for (auto& o : opts_) if (o.name == name) return &o; return nullptr; }

static bool parseBool(const std::string& v, bool* out) {  // boost::program_options bool syntax
  std::string s; for (char c : v) s += (char)tolower((unsigned char)c);
  if (s == "1" || s == "true" || s == "yes" || s == "on") { *out = true; return true; }
  if (s == "0" || s == "false" || s == "no" || s == "off") { *out = false; return true; }
  return false;
}

bool Options::parse(int argc, char* argv[]) {
  size_t pos_i = 0; size_t pos_count = 0;
  auto store = [&](Opt* o, const std::string& v) -> bool {
    bool b;
    if (o->is_bool && !parseBool(v, &b)) { std::cerr << "Error- the argument ('" << v << "') for option '--" << o->name << "' is invalid\n"; return false; }
    o->values.push_back(v); o->set = true; return true;
  };
  for (int i = 1; i < argc; ++i) {
    const std::string a = argv[i];
    Opt* o = nullptr; std::string val; bool have_val = false;
    if (a.size() > 2 && a[0] == '-' && a[1] == '-') {
      const size_t eq = a.find('=');
      const std::string name = a.substr(2, eq == std::string::npos ? std::string::npos : eq - 2);
      o = find(name);  // exact names only: prefix guessing is off (OptionsFramework.cpp:754-755)
      if (!o) { std::cerr << "Error- unrecognised option '--" << name << "'\n"; return false; }
      if (eq != std::string::npos) { val = a.substr(eq + 1); have_val = true; }
    } else if (a.size() >= 2 && a[0] == '-' && !isdigit((unsigned char)a[1]) && a[1] != '.') {
      for (auto& c : opts_) if (c.shortname && c.shortname == a[1]) o = &c;
      if (!o) { std::cerr << "Error- unrecognised option '" << a << "'\n"; return false; }
      if (a.size() > 2) { val = a.substr(2); have_val = true; }
    }
    if (o) {
      if (!have_val) {
        if (i + 1 >= argc) { std::cerr << "Error- the required argument for option '--" << o->name << "' is missing\n"; return false; }
        val = argv[++i];
      }
      if (!store(o, val)) return false;
      continue;
    }
    if (pos_i >= positional_.size()) { std::cerr << "Error- too many positional options have been specified on the command line\n"; return false; }
    Opt* po = find(positional_[pos_i].first);
    if (!po) { add(positional_[pos_i].first, 0, "", "", false); po = find(positional_[pos_i].first); }
    po->values.push_back(a); po->set = true;
    ++pos_count;
    if (positional_[pos_i].second >= 0 && (int)pos_count >= positional_[pos_i].second) { ++pos_i; pos_count = 0; }
  }
  // --config file (src/OptionsFramework.cpp:746,765-771): boost's parse_config_file -- one
  // "name = value" per line, '#' starts a comment, "[section]" prefixes the names that follow with
  // "section."; what the command line already set wins (first store wins in a variables_map), a
  // name may repeat for options that take several values (the trajectory list).
  if (const Opt* cfg = find("config")) {
    if (cfg->set) {
      std::ifstream ifs(cfg->values.back().c_str());
      if (!ifs) { std::cerr << "Error- Cannot open options config file\n"; return false; }
      std::vector<std::string> from_cmdline;
      for (const auto& o : opts_) if (o.set) from_cmdline.push_back(o.name);
      std::string line, section;
      while (std::getline(ifs, line)) {
        const size_t hash = line.find('#');
        if (hash != std::string::npos) line.erase(hash);
        auto trim = [](std::string& t) {
          const size_t a = t.find_first_not_of(" \t\r\n");
          const size_t b = t.find_last_not_of(" \t\r\n");
          t = a == std::string::npos ? "" : t.substr(a, b - a + 1);
        };
        trim(line);
        if (line.empty()) continue;
        if (line.front() == '[' && line.back() == ']') { section = line.substr(1, line.size() - 2); if (!section.empty()) section += "."; continue; }
        const size_t eq = line.find('=');
        if (eq == std::string::npos) { std::cerr << "Error- the options configuration file contains an invalid line '" << line << "'\n"; return false; }
        std::string name = line.substr(0, eq), val = line.substr(eq + 1);
        trim(name); trim(val);
        name = section + name;
        Opt* o = find(name);
        if (!o) {  // positionals are ordinary named options in the reference's options_description
          for (const auto& pz : positional_) if (pz.first == name) { add(name, 0, "", "", false); o = find(name); }
        }
        if (!o) { std::cerr << "Error- unrecognised option '" << name << "'\n"; return false; }
        if (std::find(from_cmdline.begin(), from_cmdline.end(), name) != from_cmdline.end()) continue;
        if (!store(o, val)) return false;
      }
    }
  }
  return true;
}

bool Options::has(const std::string& name) const { const Opt* o = find(name); return o && o->set; }
std::string Options::str(const std::string& name) const {
  const Opt* o = find(name);
  if (!o) return "";
  return o->set ? o->values.back() : o->defval;
}
unsigned long Options::uns(const std::string& name) const {
  const std::string v = str(name);
  char* end = nullptr;
  if (v.empty() || v[0] == '-') throw Error("the argument ('" + v + "') for option '--" + name + "' is invalid");
  const unsigned long x = strtoul(v.c_str(), &end, 10);
  if (!end || *end) throw Error("the argument ('" + v + "') for option '--" + name + "' is invalid");
  return x;
}
double Options::dbl(const std::string& name) const {
  const std::string v = str(name);
  char* end = nullptr;
  const double x = strtod(v.c_str(), &end);
  if (v.empty() || !end || *end) throw Error("the argument ('" + v + "') for option '--" + name + "' is invalid");
  return x;
}
bool Options::flag(const std::string& name) const { bool b = false; parseBool(str(name), &b); return b; }
const std::vector<std::string>& Options::multi(const std::string& name) const {
  static const std::vector<std::string> empty;
  const Opt* o = find(name);
  return o ? o->values : empty;
}
std::string Options::usage(const std::string& prog, const std::string& positional_help) const {
  std::ostringstream os;
  os << "Usage- " << prog << " [options] " << positional_help << "\nAllowed Options:\n";
  for (const auto& o : opts_) {
    if (o.help.empty()) continue;
    std::string lhs = "  ";
    if (o.shortname) lhs += std::string("-") + o.shortname + " [ --" + o.name + " ]"; else lhs += "--" + o.name;
    lhs += " arg";
    if (!o.defval.empty()) lhs += " (=" + o.defval + ")";
    os << lhs << (lhs.size() < 40 ? std::string(40 - lhs.size(), ' ') : " ") << o.help << "\n";
  }
  return os.str();
}

}  // namespace lbfe
