// params.cc -- the reference's parameter set (shared_mod.f95:28-111) read at run time.
//
// The reference fixes every parameter at compile time in shared_mod.f95's "USER-MODIFIABLE SECTION"
// (sm:38-79) and derives the rest (sm:83-111).  Here the same text -- either a whole shared_mod.f95 or
// the block that testcases/print_params.m prints for pasting -- is parsed with Fortran's literal rules,
// so the doubles the model runs with are bit-identical to what a Fortran build would hold:
//   * `9.8`, `1.e-3`, `82.8` are default-real (float32) literals, widened when assigned to real(rw);
//   * `0.5_rw`, `10._r8`, `1.d0` are doubles; `400` is an integer;
//   * mixed arithmetic promotes int -> r4 -> r8, and r4 (op) r4 is evaluated in float32.
#include "beom_host.h"

#include <cctype>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace {

std::string g_err;

enum Kind { K_INT = 0, K_R4 = 1, K_R8 = 2, K_STR = 3 };

struct Value {
  Kind kind = K_INT;
  std::vector<double> v;  // scalar = 1 element; arrays / complex = several
  std::string s;
  double scalar() const {
    if (v.empty()) throw std::runtime_error("empty value");
    return v[0];
  }
};

double round_kind(double x, Kind k) {
  if (k == K_R4) return (double)(float)x;
  if (k == K_INT) return std::trunc(x);
  return x;
}

struct Parser {
  const std::string &t;
  size_t i = 0;
  std::map<std::string, Value> &sym;
  Parser(const std::string &text, std::map<std::string, Value> &symbols) : t(text), sym(symbols) {}

  void ws() {
    while (i < t.size() && (t[i] == ' ' || t[i] == '\t')) i++;
  }
  bool eat(char c) {
    ws();
    if (i < t.size() && t[i] == c) { i++; return true; }
    return false;
  }
  bool peek2(const char *two) {
    ws();
    return i + 1 < t.size() && t[i] == two[0] && t[i + 1] == two[1];
  }

  static std::string lower(std::string s) {
    for (auto &c : s) c = (char)std::tolower((unsigned char)c);
    return s;
  }

  Kind kind_of_suffix(const std::string &sfx) {
    std::string s = lower(sfx);
    if (s == "rw" || s == "r8" || s == "8") return K_R8;
    if (s == "r4" || s == "4") return K_R4;
    auto it = sym.find(s);
    if (it != sym.end()) return it->second.scalar() >= 8 ? K_R8 : K_R4;
    throw std::runtime_error("unknown kind suffix _" + sfx);
  }

  Value number() {
    ws();
    size_t s = i;
    bool is_real = false, is_d = false;
    while (i < t.size() && std::isdigit((unsigned char)t[i])) i++;
    if (i < t.size() && t[i] == '.') {
      // not the start of an operator like .and.
      if (!(i + 1 < t.size() && std::isalpha((unsigned char)t[i + 1]) && std::tolower(t[i + 1]) != 'e' && std::tolower(t[i + 1]) != 'd')) {
        is_real = true;
        i++;
        while (i < t.size() && std::isdigit((unsigned char)t[i])) i++;
      }
    }
    std::string mant = t.substr(s, i - s);
    std::string expo;
    if (i < t.size() && (std::tolower(t[i]) == 'e' || std::tolower(t[i]) == 'd')) {
      size_t save = i;
      char ec = (char)std::tolower(t[i]);
      i++;
      size_t es = i;
      if (i < t.size() && (t[i] == '+' || t[i] == '-')) i++;
      size_t ds = i;
      while (i < t.size() && std::isdigit((unsigned char)t[i])) i++;
      if (i == ds) i = save;  // not an exponent
      else {
        is_real = true;
        is_d = (ec == 'd');
        expo = t.substr(es, i - es);
      }
    }
    Kind k = is_real ? (is_d ? K_R8 : K_R4) : K_INT;
    if (i < t.size() && t[i] == '_') {
      i++;
      size_t ks = i;
      while (i < t.size() && (std::isalnum((unsigned char)t[i]) || t[i] == '_')) i++;
      Kind kk = kind_of_suffix(t.substr(ks, i - ks));
      if (is_real) k = kk;
    }
    std::string lit = mant + (expo.empty() ? "" : "e" + expo);
    Value r;
    r.kind = k;
    if (k == K_R4) r.v.push_back((double)std::strtof(lit.c_str(), nullptr));  // correctly rounded to float32
    else r.v.push_back(std::strtod(lit.c_str(), nullptr));
    return r;
  }

  Value primary() {
    ws();
    if (i >= t.size()) throw std::runtime_error("unexpected end of expression");
    char c = t[i];
    if (c == '\'' || c == '"') {
      char q = c;
      i++;
      Value r;
      r.kind = K_STR;
      while (i < t.size()) {
        if (t[i] == q) {
          if (i + 1 < t.size() && t[i + 1] == q) { r.s.push_back(q); i += 2; continue; }
          i++;
          break;
        }
        r.s.push_back(t[i++]);
      }
      return r;
    }
    if (peek2("(/")) {  // array constructor
      i += 2;
      Value r;
      r.kind = K_INT;
      for (;;) {
        Value e = expr();
        if (e.kind > r.kind) r.kind = e.kind;
        r.v.push_back(e.scalar());
        ws();
        if (eat(',')) continue;
        if (peek2("/)")) { i += 2; break; }
        throw std::runtime_error("bad array constructor");
      }
      return r;
    }
    if (c == '(') {
      i++;
      Value a = expr();
      ws();
      if (eat(',')) {  // complex literal (re, im)
        Value b = expr();
        if (!eat(')')) throw std::runtime_error("expected )");
        Value r;
        r.kind = a.kind > b.kind ? a.kind : b.kind;
        r.v = {a.scalar(), b.scalar()};
        return r;
      }
      if (!eat(')')) throw std::runtime_error("expected )");
      return a;
    }
    if (std::isdigit((unsigned char)c) || c == '.') return number();
    if (std::isalpha((unsigned char)c)) {
      size_t s = i;
      while (i < t.size() && (std::isalnum((unsigned char)t[i]) || t[i] == '_')) i++;
      std::string name = lower(t.substr(s, i - s));
      ws();
      if (name == "selected_real_kind" || name == "selected_int_kind") {
        // selected_real_kind(P = 6) -> 4, (P = 12) -> 8; selected_int_kind -> 4
        int depth = 0;
        std::string args;
        do {
          if (t[i] == '(') depth++;
          if (t[i] == ')') depth--;
          args.push_back(t[i]);
          i++;
        } while (i < t.size() && depth > 0);
        int digits = 0;
        for (char ch : args)
          if (std::isdigit((unsigned char)ch)) digits = digits * 10 + (ch - '0');
        Value r;
        r.kind = K_INT;
        r.v.push_back(name == "selected_real_kind" ? (digits > 6 ? 8 : 4) : 4);
        return r;
      }
      auto it = sym.find(name);
      if (it == sym.end()) throw std::runtime_error("unknown name '" + name + "' in expression");
      if (i < t.size() && t[i] == '(') {  // array element
        i++;
        Value idx = expr();
        if (!eat(')')) throw std::runtime_error("expected ) after index");
        long k = (long)idx.scalar();
        if (k < 1 || (size_t)k > it->second.v.size()) throw std::runtime_error("index out of range for " + name);
        Value r;
        r.kind = it->second.kind;
        r.v.push_back(it->second.v[(size_t)k - 1]);
        return r;
      }
      return it->second;
    }
    throw std::runtime_error(std::string("unexpected character '") + c + "'");
  }

  Value power() {
    Value a = primary();
    ws();
    if (peek2("**")) {
      i += 2;
      Value b = unary();  // right associative
      Value r;
      if (b.kind == K_INT) {  // integer exponent: repeated multiplication in the base's kind
        long n = (long)b.scalar();
        r.kind = a.kind;
        double x = a.scalar(), y = 1.0;
        for (long k = 0; k < std::labs(n); k++) y = round_kind(y * x, a.kind);
        if (n < 0) y = (a.kind == K_INT) ? std::trunc(1.0 / y) : round_kind(1.0 / y, a.kind);
        r.v.push_back(y);
      } else {
        r.kind = a.kind > b.kind ? a.kind : b.kind;
        r.v.push_back(round_kind(std::pow(a.scalar(), b.scalar()), r.kind));
      }
      return r;
    }
    return a;
  }
  Value unary() {
    ws();
    if (eat('-')) {
      Value a = unary();
      for (auto &x : a.v) x = -x;
      return a;
    }
    if (eat('+')) return unary();
    return power();
  }
  static Value binop(const Value &a, const Value &b, char op) {
    Value r;
    r.kind = a.kind > b.kind ? a.kind : b.kind;
    if (r.kind == K_STR) throw std::runtime_error("arithmetic on strings");
    size_t n = a.v.size() > b.v.size() ? a.v.size() : b.v.size();
    for (size_t k = 0; k < n; k++) {
      double x = a.v[a.v.size() == 1 ? 0 : k], y = b.v[b.v.size() == 1 ? 0 : k], z = 0;
      if (r.kind == K_R4) {
        float fx = (float)x, fy = (float)y, fz = 0;
        if (op == '+') fz = fx + fy;
        if (op == '-') fz = fx - fy;
        if (op == '*') fz = fx * fy;
        if (op == '/') fz = fx / fy;
        z = (double)fz;
      } else {
        if (op == '+') z = x + y;
        if (op == '-') z = x - y;
        if (op == '*') z = x * y;
        if (op == '/') z = x / y;
        if (r.kind == K_INT) z = std::trunc(z);
      }
      r.v.push_back(z);
    }
    return r;
  }
  Value term() {
    Value a = unary();
    for (;;) {
      ws();
      if (peek2("**") || peek2("/)")) return a;
      if (i < t.size() && (t[i] == '*' || t[i] == '/')) {
        char op = t[i++];
        Value b = unary();
        a = binop(a, b, op);
      } else
        return a;
    }
  }
  Value expr() {
    Value a = term();
    for (;;) {
      ws();
      if (i < t.size() && (t[i] == '+' || t[i] == '-')) {
        char op = t[i++];
        Value b = term();
        a = binop(a, b, op);
      } else
        return a;
    }
  }
};

// names whose declared type in shared_mod.f95 is real(rw)/real(r8)/complex(rw): assignment widens to double
const char *const k_real_names[] = {"dl", "cext", "f0", "rhon", "topl", "dt_s", "dt_o", "dt_r", "dt3d", "bvis", "dvis",
                                    "bdrg", "hmin", "hsbl", "hbbl", "g_fb", "uadv", "qdrg", "ocrp", "rsta", "xper", "yper",
                                    "diag", "rgld", "mcbc", "tauw", "svis", "tdrg", "topt", "plum", "dt", "hsal", "hdry",
                                    "tole", "pi", "grav", "rho0", "beta", "epsi", "gamm", "del1", "del2", "sor", nullptr};

bool is_real_name(const std::string &n) {
  for (int k = 0; k_real_names[k]; k++)
    if (n == k_real_names[k]) return true;
  return false;
}

// strip comments, join continuation lines; statements are separated by '\n'
std::string normalise(const std::string &text) {
  std::string out;
  std::istringstream in(text);
  std::string line;
  bool cont = false;
  while (std::getline(in, line)) {
    std::string s;
    char q = 0;
    for (size_t k = 0; k < line.size(); k++) {
      char c = line[k];
      if (q) {
        s.push_back(c);
        if (c == q) q = 0;
        continue;
      }
      if (c == '\'' || c == '"') { q = c; s.push_back(c); continue; }
      if (c == '!') break;
      if (c == '\r') continue;
      s.push_back(c);
    }
    size_t e = s.find_last_not_of(" \t");
    bool will_cont = false;
    if (e != std::string::npos && s[e] == '&') { will_cont = true; s.erase(e); }
    size_t b = s.find_first_not_of(" \t");
    if (cont && b != std::string::npos && s[b] == '&') s.erase(0, b + 1);
    out += s;
    if (!will_cont) out += '\n';
    else out += ' ';
    cont = will_cont;
  }
  return out;
}

void store(std::map<std::string, Value> &sym, const std::string &name, Value v) {
  if (is_real_name(name) && v.kind != K_STR) v.kind = K_R8;  // widened on assignment, value unchanged
  sym[name] = v;
}

void parse_statement(const std::string &st, std::map<std::string, Value> &sym) {
  // find "name [(dims)] = expr" items at parenthesis depth 0, separated by commas
  size_t start = 0;
  size_t dc = st.find("::");
  if (dc != std::string::npos) start = dc + 2;
  else {
    // a bare "name = value" line (print_params block) or something else: require an identifier first
    size_t b = st.find_first_not_of(" \t");
    if (b == std::string::npos || !std::isalpha((unsigned char)st[b])) return;
    // skip executable / structural statements of a full shared_mod.f95
    static const char *const skip[] = {"module", "use", "implicit", "private", "public", "contains", "end", "subroutine",
                                       "function", "if", "do", "write", "inquire", "close", "stop", "return", "else",
                                       "call", "exit", "integer", "logical", "character", "real", "complex", nullptr};
    size_t e = b;
    while (e < st.size() && (std::isalnum((unsigned char)st[e]) || st[e] == '_')) e++;
    std::string w = Parser::lower(st.substr(b, e - b));
    for (int k = 0; skip[k]; k++)
      if (w == skip[k]) return;
  }
  Parser p(st, sym);
  p.i = start;
  for (;;) {
    p.ws();
    if (p.i >= st.size()) return;
    if (!std::isalpha((unsigned char)st[p.i])) return;
    size_t s = p.i;
    while (p.i < st.size() && (std::isalnum((unsigned char)st[p.i]) || st[p.i] == '_')) p.i++;
    std::string name = Parser::lower(st.substr(s, p.i - s));
    p.ws();
    if (p.i < st.size() && st[p.i] == '(') {  // dims, e.g. rhon(nlay)
      int depth = 0;
      do {
        if (st[p.i] == '(') depth++;
        if (st[p.i] == ')') depth--;
        p.i++;
      } while (p.i < st.size() && depth > 0);
      p.ws();
    }
    if (p.i < st.size() && st[p.i] == '=' && !(p.i + 1 < st.size() && (st[p.i + 1] == '=' || st[p.i + 1] == '>'))) {
      p.i++;
      Value v = p.expr();
      store(sym, name, v);
      p.ws();
    }
    if (p.i < st.size() && st[p.i] == ',') { p.i++; continue; }
    return;
  }
}

double getd(const std::map<std::string, Value> &sym, const char *n, double dflt) {
  auto it = sym.find(n);
  return it == sym.end() ? dflt : it->second.scalar();
}

void copy_str(char *dst, int slen, const std::map<std::string, Value> &sym, const char *n) {
  if (!dst || slen <= 0) return;
  auto it = sym.find(n);
  std::string s = it == sym.end() ? std::string() : it->second.s;
  std::snprintf(dst, (size_t)slen, "%s", s.c_str());
}

}  // namespace

extern "C" {

int beom_host_last_error(char *buf, int len) {
  if (buf && len > 0) std::snprintf(buf, (size_t)len, "%s", g_err.c_str());
  return (int)g_err.size();
}

void beom_params_defaults(beom_params *p) {
  std::memset(p, 0, sizeof *p);
  p->mcbc = 1.0;  // shared_mod.f95:70-71 as shipped
  p->rgld = 0.0;
  p->hdry = (double)1.e-3f;  // sm:86-93: default-real literals
  p->tole = (double)1.e-6f;
  p->pi = (double)3.1415927f;
  p->grav = (double)9.8f;
  p->beta = (double)0.281105f;
  p->epsi = (double)0.013f;
  p->gamm = (double)0.088f;
  p->sor = (double)1.9f;  // sm:99
  p->itmx = 99999;        // sm:104-105
  p->nsal = 4;
  p->variant = BEOM_VARIANT_STANDARD;
}

void beom_params_derive(beom_params *p) {
  p->dt = 0.5 * p->dl / p->cext;  // sm:84
  p->hsal = 10.0 * p->hmin;       // sm:85
  if (p->nlay >= 1 && p->nlay <= BEOM_MAXLAY) p->rho0 = p->rhon[p->nlay - 1];  // sm:90
  p->del1 = 0.5 + p->gamm + 2.0 * p->epsi;                                     // sm:94
  p->del2 = 1.0 - p->del1 - p->gamm - p->epsi;                                 // sm:95
}

int beom_params_parse(const char *text, beom_params *out, char *idir, char *odir, char *desc, int slen) {
  g_err.clear();
  try {
    std::map<std::string, Value> sym;
    // kinds (sm:28-32) and a few names the expressions may use
    Value k4, k8;
    k4.v = {4};
    k8.v = {8};
    sym["i4"] = k4;
    sym["r4"] = k4;
    sym["r8"] = k8;
    sym["rw"] = k8;
    std::string norm = normalise(text ? text : "");
    std::istringstream in(norm);
    std::string st;
    while (std::getline(in, st)) {
      size_t b = st.find_first_not_of(" \t");
      if (b == std::string::npos) continue;
      if (Parser::lower(st.substr(b, 8)) == "contains") break;  // module procedures follow (sm:118)
      try {
        parse_statement(st, sym);
      } catch (const std::exception &) {
        // only declarations and known parameters must parse; anything else is not ours
        size_t e = b;
        while (e < st.size() && (std::isalnum((unsigned char)st[e]) || st[e] == '_')) e++;
        std::string w = Parser::lower(st.substr(b, e - b));
        if (st.find("::") != std::string::npos || is_real_name(w) || w == "lm" || w == "mm" || w == "nlay" || w == "ndeg") throw;
      }
    }
    if (sym["rw"].scalar() < 8) throw std::runtime_error("rw = r4 (single-precision build) is not supported");

    beom_params p;
    beom_params_defaults(&p);
    p.lm = (int32_t)getd(sym, "lm", 0);
    p.mm = (int32_t)getd(sym, "mm", 0);
    p.nlay = (int32_t)getd(sym, "nlay", 0);
    p.ndeg = (int32_t)getd(sym, "ndeg", 0);
    if (p.nlay < 1 || p.nlay > BEOM_MAXLAY) throw std::runtime_error("nlay missing or out of range (1..16)");
    if (p.lm < 1 || p.mm < 1) throw std::runtime_error("grid dimensions (lm,mm) should be >= 1.");
#define GET(n) p.n = getd(sym, #n, p.n)
    GET(dl); GET(cext); GET(f0); GET(dt_s); GET(dt_o); GET(dt_r); GET(dt3d); GET(bvis); GET(dvis); GET(bdrg);
    GET(hmin); GET(hsbl); GET(hbbl); GET(g_fb); GET(uadv); GET(qdrg); GET(ocrp); GET(rsta); GET(xper); GET(yper);
    GET(diag); GET(rgld); GET(mcbc); GET(svis); GET(tdrg); GET(topt); GET(plum);
    GET(hdry); GET(tole); GET(pi); GET(grav); GET(beta); GET(epsi); GET(gamm); GET(sor);
#undef GET
    p.itmx = (int32_t)getd(sym, "itmx", p.itmx);
    p.nsal = (int32_t)getd(sym, "nsal", p.nsal);
    auto arr = [&](const char *n, double *dst) {
      auto it = sym.find(n);
      if (it == sym.end()) throw std::runtime_error(std::string(n) + " missing");
      if ((int)it->second.v.size() < p.nlay) throw std::runtime_error(std::string(n) + " has fewer than nlay entries");
      for (int k = 0; k < p.nlay; k++) dst[k] = it->second.v[(size_t)k];
    };
    arr("rhon", p.rhon);
    arr("topl", p.topl);
    auto tw = sym.find("tauw");
    if (tw != sym.end() && tw->second.v.size() >= 2) { p.tauw[0] = tw->second.v[0]; p.tauw[1] = tw->second.v[1]; }
    beom_params_derive(&p);
    // a full shared_mod.f95 spells the derived constants out; honour them if present
#define GETD(n) p.n = getd(sym, #n, p.n)
    GETD(dt); GETD(hsal); GETD(rho0); GETD(del1); GETD(del2);
#undef GETD
    *out = p;
    copy_str(idir, slen, sym, "idir");
    copy_str(odir, slen, sym, "odir");
    copy_str(desc, slen, sym, "desc");
    return 0;
  } catch (const std::exception &e) {
    g_err = std::string("beom_params_parse: ") + e.what();
    return -1;
  }
}

int beom_params_parse_file(const char *path, beom_params *out, char *idir, char *odir, char *desc, int slen) {
  std::ifstream f(path);
  if (!f) {
    g_err = std::string("beom_params_parse_file: cannot open ") + path;
    return -2;
  }
  std::stringstream ss;
  ss << f.rdbuf();
  return beom_params_parse(ss.str().c_str(), out, idir, odir, desc, slen);
}

}  // extern "C"

// shared with the other host translation units
void beom_host_set_error(const std::string &s) { g_err = s; }
