/* counted.h -- op-counting scalar for the fp64 oracle (TEST INFRASTRUCTURE, like everything
 * under oracle/).  oracle/opcount/count_ops.py compiles mjstep_oracle.c as C++ with every
 * `double` of the oracle's own code replaced by cnt_t, so each arithmetic operation of the
 * restated MuJoCo pipeline bumps a counter: the "algorithmic FLOPs per env-step" figure
 * SURVEY.md section 8(d) / BASELINE.md section 3 ask builder and judge to share.
 * cnt_t has the layout of a double, so the NumPy-owned arrays of oracle.py pass unchanged. */
#pragma once
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <type_traits>

enum { OPC_ADD = 0, OPC_MUL, OPC_DIV, OPC_SQRT, OPC_TRANS, OPC_CMP, OPC_N };
extern long long g_opc[OPC_N];

struct cnt_t {
  double v;
  cnt_t() = default;
  cnt_t(double x) : v(x) {}
  explicit operator double() const { return v; }
  explicit operator int() const { return (int)v; }
  explicit operator bool() const { return v != 0; }
  cnt_t operator-() const { return cnt_t(-v); }
  cnt_t operator+() const { return *this; }
  bool operator!() const { return v == 0; }
  cnt_t &operator+=(cnt_t o) { g_opc[OPC_ADD]++; v += o.v; return *this; }
  cnt_t &operator-=(cnt_t o) { g_opc[OPC_ADD]++; v -= o.v; return *this; }
  cnt_t &operator*=(cnt_t o) { g_opc[OPC_MUL]++; v *= o.v; return *this; }
  cnt_t &operator/=(cnt_t o) { g_opc[OPC_DIV]++; v /= o.v; return *this; }
};
static_assert(sizeof(cnt_t) == sizeof(double), "cnt_t must have the layout of a double");

template <class T> using arith_t = typename std::enable_if<std::is_arithmetic<T>::value, int>::type;
#define OPC_BIN(op_, slot_) \
  inline cnt_t operator op_(cnt_t a, cnt_t b) { g_opc[slot_]++; return cnt_t(a.v op_ b.v); } \
  template <class T, arith_t<T> = 0> inline cnt_t operator op_(cnt_t a, T b) { g_opc[slot_]++; return cnt_t(a.v op_ (double)b); } \
  template <class T, arith_t<T> = 0> inline cnt_t operator op_(T a, cnt_t b) { g_opc[slot_]++; return cnt_t((double)a op_ b.v); }
OPC_BIN(+, OPC_ADD) OPC_BIN(-, OPC_ADD) OPC_BIN(*, OPC_MUL) OPC_BIN(/, OPC_DIV)
#undef OPC_BIN
#define OPC_REL(op_) \
  inline bool operator op_(cnt_t a, cnt_t b) { g_opc[OPC_CMP]++; return a.v op_ b.v; } \
  template <class T, arith_t<T> = 0> inline bool operator op_(cnt_t a, T b) { g_opc[OPC_CMP]++; return a.v op_ (double)b; } \
  template <class T, arith_t<T> = 0> inline bool operator op_(T a, cnt_t b) { g_opc[OPC_CMP]++; return (double)a op_ b.v; }
OPC_REL(<) OPC_REL(>) OPC_REL(<=) OPC_REL(>=) OPC_REL(==) OPC_REL(!=)
#undef OPC_REL

#define OPC_FN1(name_, slot_) inline cnt_t name_(cnt_t a) { g_opc[slot_]++; return cnt_t(std::name_(a.v)); }
OPC_FN1(sqrt, OPC_SQRT) OPC_FN1(sin, OPC_TRANS) OPC_FN1(cos, OPC_TRANS) OPC_FN1(tan, OPC_TRANS)
OPC_FN1(exp, OPC_TRANS) OPC_FN1(log, OPC_TRANS) OPC_FN1(acos, OPC_TRANS) OPC_FN1(asin, OPC_TRANS)
OPC_FN1(atan, OPC_TRANS) OPC_FN1(tanh, OPC_TRANS)
#undef OPC_FN1
inline cnt_t fabs(cnt_t a) { return cnt_t(std::fabs(a.v)); }      /* sign-bit operation: not counted */
inline cnt_t floor(cnt_t a) { return cnt_t(std::floor(a.v)); }
inline cnt_t ceil(cnt_t a) { return cnt_t(std::ceil(a.v)); }
inline int isfinite(cnt_t a) { return std::isfinite(a.v); }
inline int isnan(cnt_t a) { return std::isnan(a.v); }
#define OPC_FN2(name_, slot_) \
  inline cnt_t name_(cnt_t a, cnt_t b) { g_opc[slot_]++; return cnt_t(std::name_(a.v, b.v)); } \
  template <class T, arith_t<T> = 0> inline cnt_t name_(cnt_t a, T b) { g_opc[slot_]++; return cnt_t(std::name_(a.v, (double)b)); } \
  template <class T, arith_t<T> = 0> inline cnt_t name_(T a, cnt_t b) { g_opc[slot_]++; return cnt_t(std::name_((double)a, b.v)); }
OPC_FN2(fmax, OPC_CMP) OPC_FN2(fmin, OPC_CMP) OPC_FN2(pow, OPC_TRANS) OPC_FN2(atan2, OPC_TRANS)
OPC_FN2(copysign, OPC_CMP) OPC_FN2(fmod, OPC_DIV)
#undef OPC_FN2
