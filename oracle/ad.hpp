// ORACLE (test infrastructure, never shipped, never on the product path).
//
// CPU restatement of the reference's tape-based reverse-mode AD, qs/lib/ad.t
// (the type the reference calls "dualnum"), with its bump allocator,
// qs/lib/memoryPool.t.  The CUDA engine replaces all of this with closed-form
// gradients; the oracle keeps it because (i) it is the independent check on those
// closed forms and (ii) it is where the reference spends its time, so it is what
// the CPU baseline must time.
//
//   * every arithmetic op allocates one node {val, adj, captured operands} from the
//     pool and pushes {node, adjoint fn} on a global tape          ad.t:41-79
//   * constants cast to num get a node but no tape entry            ad.t:91-97
//   * accumadj skips when the output adjoint is exactly 0           ad.t:166-172
//   * grad(): adj=1 on the result, then walk the tape backwards     ad.t:679-685
//   * recoverMemory(): clear tape, rewind pool                      ad.t:673-676
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <vector>

namespace qso {
namespace ad {

// ---------------------------------------------------------------- memoryPool.t:25-137
struct MemoryPool {
  std::vector<char*> blocks;
  std::vector<size_t> sizes;
  size_t cur_block = 0;
  char* cur_end = nullptr;
  char* next_loc = nullptr;
  size_t currAlloced = 0, maxAlloced = 0;
  MemoryPool() {
    const size_t n0 = 65536;  // DEFAULT_INITIAL_BYTES, memoryPool.t:12
    blocks.push_back((char*)std::malloc(n0));
    sizes.push_back(n0);
    next_loc = blocks[0];
    cur_end = blocks[0] + n0;
  }
  ~MemoryPool() { for (char* b : blocks) std::free(b); }
  char* move_to_next_block(size_t len) {  // memoryPool.t:64-90
    ++cur_block;
    while (cur_block < blocks.size() && sizes[cur_block] < len) ++cur_block;
    if (cur_block >= blocks.size()) {
      size_t newsize = sizes.back() * 2;
      if (newsize < len) newsize = len;
      blocks.push_back((char*)std::malloc(newsize));
      sizes.push_back(newsize);
    }
    char* result = blocks[cur_block];
    next_loc = result + len;
    cur_end = result + sizes[cur_block];
    return result;
  }
  inline void* alloc(size_t len) {  // memoryPool.t:92-102
    char* result = next_loc;
    next_loc += len;
    if (next_loc >= cur_end) result = move_to_next_block(len);
    currAlloced += len;
    return result;
  }
  void recoverAll() {  // memoryPool.t:105-114
    cur_block = 0;
    next_loc = blocks[0];
    cur_end = next_loc + sizes[0];
    if (currAlloced > maxAlloced) maxAlloced = currAlloced;
    currAlloced = 0;
  }
};

struct Base { double val, adj; };            // ad.t:43-47
typedef void (*AdjointFn)(void*);
struct TapeEntry { void* datum; AdjointFn fn; };   // ad.t:14-19

struct Globals {
  MemoryPool pool;
  std::vector<TapeEntry> tape;
  uint64_t nodes = 0;   // statistics: nodes allocated since process start
};
inline Globals& G() { static thread_local Globals g; return g; }

struct num {                                  // ad.t:83-97
  Base* impl;
  num() : impl(nullptr) {}
  explicit num(Base* p) : impl(p) {}
  num(double v) {                             // __cast: node without tape entry (ad.t:91-97)
    impl = (Base*)G().pool.alloc(sizeof(Base));
    impl->val = v; impl->adj = 0.0; ++G().nodes;
  }
  double val() const { return impl->val; }
  double adj() const { return impl->adj; }
};

template <class N> inline N* newnode(double v, AdjointFn fn) {   // DualNumT.new, ad.t:56-75
  Globals& g = G();
  N* p = (N*)g.pool.alloc(sizeof(N));
  p->val = v; p->adj = 0.0; ++g.nodes;
  if (fn) g.tape.push_back(TapeEntry{p, fn});
  return p;
}

// accumadj (ad.t:166-172)
#define QSO_ACC(outp, v, amount) do { if ((outp)->adj != 0.0) (v).impl->adj += (outp)->adj * (amount); } while (0)

struct N1 : Base { num a; };
struct N2 : Base { num a, b; };
struct N1d : Base { num a; double b; };   // num op double (both operands captured when the adjoint reads them)
struct Nd1 : Base { double a; num b; };

// ---- ADD (ad.t:375-383)
inline void adj_add_nn(void* p) { N2* n = (N2*)p; QSO_ACC(n, n->a, 1.0); QSO_ACC(n, n->b, 1.0); }
inline void adj_add_n(void* p) { N1* n = (N1*)p; QSO_ACC(n, n->a, 1.0); }
inline num operator+(num a, num b) { N2* n = newnode<N2>(a.val() + b.val(), adj_add_nn); n->a = a; n->b = b; return num(n); }
inline num operator+(num a, double b) { N1* n = newnode<N1>(a.val() + b, adj_add_n); n->a = a; return num(n); }
inline num operator+(double a, num b) { N1* n = newnode<N1>(a + b.val(), adj_add_n); n->a = b; return num(n); }
// ---- SUB (ad.t:385-393)
inline void adj_sub_nn(void* p) { N2* n = (N2*)p; QSO_ACC(n, n->a, 1.0); QSO_ACC(n, n->b, -1.0); }
inline void adj_sub_dn(void* p) { N1* n = (N1*)p; QSO_ACC(n, n->a, -1.0); }
inline num operator-(num a, num b) { N2* n = newnode<N2>(a.val() - b.val(), adj_sub_nn); n->a = a; n->b = b; return num(n); }
inline num operator-(num a, double b) { N1* n = newnode<N1>(a.val() - b, adj_add_n); n->a = a; return num(n); }
inline num operator-(double a, num b) { N1* n = newnode<N1>(a - b.val(), adj_sub_dn); n->a = b; return num(n); }
// ---- MUL (ad.t:395-403)
inline void adj_mul_nn(void* p) { N2* n = (N2*)p; QSO_ACC(n, n->a, n->b.val()); QSO_ACC(n, n->b, n->a.val()); }
inline void adj_mul_nd(void* p) { N1d* n = (N1d*)p; QSO_ACC(n, n->a, n->b); }
inline num operator*(num a, num b) { N2* n = newnode<N2>(a.val() * b.val(), adj_mul_nn); n->a = a; n->b = b; return num(n); }
inline num operator*(num a, double b) { N1d* n = newnode<N1d>(a.val() * b, adj_mul_nd); n->a = a; n->b = b; return num(n); }
inline num operator*(double a, num b) { N1d* n = newnode<N1d>(a * b.val(), adj_mul_nd); n->a = b; n->b = a; return num(n); }
// ---- DIV (ad.t:405-413)
inline void adj_div_nn(void* p) { N2* n = (N2*)p; double a = n->a.val(), b = n->b.val(); QSO_ACC(n, n->a, 1.0 / b); QSO_ACC(n, n->b, -a / (b * b)); }
inline void adj_div_nd(void* p) { N1d* n = (N1d*)p; QSO_ACC(n, n->a, 1.0 / n->b); }
inline void adj_div_dn(void* p) { Nd1* n = (Nd1*)p; double b = n->b.val(); QSO_ACC(n, n->b, -n->a / (b * b)); }
inline num operator/(num a, num b) { N2* n = newnode<N2>(a.val() / b.val(), adj_div_nn); n->a = a; n->b = b; return num(n); }
inline num operator/(num a, double b) { N1d* n = newnode<N1d>(a.val() / b, adj_div_nd); n->a = a; n->b = b; return num(n); }
inline num operator/(double a, num b) { Nd1* n = newnode<Nd1>(a / b.val(), adj_div_dn); n->a = a; n->b = b; return num(n); }
// ---- UNM (ad.t:415-422)
inline void adj_neg(void* p) { N1* n = (N1*)p; QSO_ACC(n, n->a, -1.0); }
inline num operator-(num a) { N1* n = newnode<N1>(-a.val(), adj_neg); n->a = a; return num(n); }
// ---- comparisons act on primal values (ad.t:424-442)
#define QSO_CMP(op) \
  inline bool operator op(num a, num b) { return a.val() op b.val(); } \
  inline bool operator op(num a, double b) { return a.val() op b; } \
  inline bool operator op(double a, num b) { return a op b.val(); }
QSO_CMP(==) QSO_CMP(<) QSO_CMP(<=) QSO_CMP(>) QSO_CMP(>=)
#undef QSO_CMP

// ---- functions
// EXP (ad.t:519-526): d/da = val(v)
inline void adj_exp(void* p) { N1* n = (N1*)p; QSO_ACC(n, n->a, n->val); }
inline num exp(num a) { N1* n = newnode<N1>(std::exp(a.val()), adj_exp); n->a = a; return num(n); }
// LOG (ad.t:577-584)
inline void adj_log(void* p) { N1* n = (N1*)p; QSO_ACC(n, n->a, 1.0 / n->a.val()); }
inline num log(num a) { N1* n = newnode<N1>(std::log(a.val()), adj_log); n->a = a; return num(n); }
// SQRT (ad.t:631-638): d/da = 1/(2 val(v))
inline void adj_sqrt(void* p) { N1* n = (N1*)p; QSO_ACC(n, n->a, 1.0 / (2.0 * n->val)); }
inline num sqrt(num a) { N1* n = newnode<N1>(std::sqrt(a.val()), adj_sqrt); n->a = a; return num(n); }
// POW (ad.t:596-606)
inline void adj_pow_nn(void* p) { N2* n = (N2*)p; double a = n->a.val(); if (a != 0.0) { QSO_ACC(n, n->a, n->b.val() * n->val / a); QSO_ACC(n, n->b, std::log(a) * n->val); } }
inline void adj_pow_nd(void* p) { N1d* n = (N1d*)p; double a = n->a.val(); if (a != 0.0) { QSO_ACC(n, n->a, n->b * n->val / a); } }
inline void adj_pow_dn(void* p) { Nd1* n = (Nd1*)p; if (n->a != 0.0) { QSO_ACC(n, n->b, std::log(n->a) * n->val); } }
inline num pow(num a, num b) { N2* n = newnode<N2>(std::pow(a.val(), b.val()), adj_pow_nn); n->a = a; n->b = b; return num(n); }
inline num pow(num a, double b) { N1d* n = newnode<N1d>(std::pow(a.val(), b), adj_pow_nd); n->a = a; n->b = b; return num(n); }
inline num pow(double a, num b) { Nd1* n = newnode<Nd1>(std::pow(a, b.val()), adj_pow_dn); n->a = a; n->b = b; return num(n); }
// TANH (ad.t:649-657)
inline void adj_tanh(void* p) { N1* n = (N1*)p; double c = std::cosh(n->a.val()); QSO_ACC(n, n->a, 1.0 / (c * c)); }
inline num tanh(num a) { N1* n = newnode<N1>(std::tanh(a.val()), adj_tanh); n->a = a; return num(n); }
// FABS (ad.t:529-536): returns a or -a
inline num fabs(num a) { if (a.val() >= 0.0) return a; return -a; }
// FMAX / FMIN (ad.t:545-575): the constant branch yields a fresh constant => zero gradient
inline num fmax(num a, double b) { if (a.val() >= b) return a; return num(b); }
inline num fmax(double a, num b) { if (a > b.val()) return num(a); return b; }
inline num fmax(num a, num b) { if (a.val() > b.val()) return a; return b; }
inline num fmin(num a, double b) { if (a.val() <= b) return a; return num(b); }
inline num fmin(double a, num b) { if (a < b.val()) return num(a); return b; }
inline num fmin(num a, num b) { if (a.val() < b.val()) return a; return b; }
#undef QSO_ACC

// ---- derivative computation (ad.t:662-711)
inline void recoverMemory() { G().tape.clear(); G().pool.recoverAll(); }
inline void grad(num n) {
  n.impl->adj = 1.0;
  std::vector<TapeEntry>& t = G().tape;
  for (size_t i = 0; i < t.size(); ++i) { size_t j = t.size() - i - 1; t[j].fn(t[j].datum); }
}
inline void grad(num n, const std::vector<num>& indeps, std::vector<double>& gradient) {   // ad.t:699-708
  grad(n);
  gradient.clear();
  gradient.reserve(indeps.size());
  for (size_t i = 0; i < indeps.size(); ++i) gradient.push_back(indeps[i].adj());
  recoverMemory();
}

}  // namespace ad

// qs.val (ad.t:131-137 / globals.t:23)
inline double val(double x) { return x; }
inline double val(ad::num x) { return x.val(); }

// tmath = ad.math (qs/lib/tmath.t:7-8): math.h overloaded for num
namespace tmath {
using std::exp; using std::log; using std::sqrt; using std::pow; using std::fabs; using std::fmax; using std::fmin; using std::tanh;
using ad::exp; using ad::log; using ad::sqrt; using ad::pow; using ad::fabs; using ad::fmax; using ad::fmin; using ad::tanh;
}  // namespace tmath

}  // namespace qso
