// ORACLE — TEST INFRASTRUCTURE ONLY (never linked into the product library).
// Forward-mode automatic differentiation scalars used by the CPU oracle.
//
// The reference obtains every Jacobian on the MPC path from CppAD tapes
// (qm_interface/src/dynamics/QMDynamicsAD.cpp:22-33, QMInterface.cpp:363-379) and every
// Jacobian / Jacobian time-derivative on the WBC path from Pinocchio
// (qm_wbc/src/WbcBase.cpp:150-190).  Neither library exists in this container, so the oracle
// differentiates the same forward-kinematics composition with dual numbers instead.  This is
// deliberately a different technique from the hand-derived analytic Jacobians in the CUDA
// kernels, so that agreement between the two is evidence and not tautology.
#pragma once
#include <cmath>

namespace orc {

// Dual<N,S>: value + N directional derivatives, scalar field S (double or another Dual → nesting).
template <int N, class S = double>
struct Dual {
  S v;
  S d[N];
  Dual() : v(S(0)) { for (int i = 0; i < N; ++i) d[i] = S(0); }
  Dual(double c) : v(S(c)) { for (int i = 0; i < N; ++i) d[i] = S(0); }
  template <class U = S, class = typename std::enable_if<!std::is_same<U, double>::value>::type>
  Dual(const S& c) : v(c) { for (int i = 0; i < N; ++i) d[i] = S(0); }
  static Dual variable(const S& value, int idx) { Dual r; r.v = value; r.d[idx] = S(1.0); return r; }
};

template <int N, class S> inline Dual<N, S> operator+(const Dual<N, S>& a, const Dual<N, S>& b) {
  Dual<N, S> r; r.v = a.v + b.v; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] + b.d[i]; return r; }
template <int N, class S> inline Dual<N, S> operator-(const Dual<N, S>& a, const Dual<N, S>& b) {
  Dual<N, S> r; r.v = a.v - b.v; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] - b.d[i]; return r; }
template <int N, class S> inline Dual<N, S> operator-(const Dual<N, S>& a) {
  Dual<N, S> r; r.v = -a.v; for (int i = 0; i < N; ++i) r.d[i] = -a.d[i]; return r; }
template <int N, class S> inline Dual<N, S> operator*(const Dual<N, S>& a, const Dual<N, S>& b) {
  Dual<N, S> r; r.v = a.v * b.v; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * b.v + a.v * b.d[i]; return r; }
template <int N, class S> inline Dual<N, S> operator/(const Dual<N, S>& a, const Dual<N, S>& b) {
  Dual<N, S> r; S inv = S(1.0) / b.v; r.v = a.v * inv;
  for (int i = 0; i < N; ++i) r.d[i] = (a.d[i] - r.v * b.d[i]) * inv; return r; }
template <int N, class S> inline Dual<N, S> operator+(const Dual<N, S>& a, double b) { Dual<N, S> r = a; r.v = r.v + S(b); return r; }
template <int N, class S> inline Dual<N, S> operator+(double b, const Dual<N, S>& a) { return a + b; }
template <int N, class S> inline Dual<N, S> operator-(const Dual<N, S>& a, double b) { Dual<N, S> r = a; r.v = r.v - S(b); return r; }
template <int N, class S> inline Dual<N, S> operator-(double b, const Dual<N, S>& a) { return (-a) + b; }
template <int N, class S> inline Dual<N, S> operator*(const Dual<N, S>& a, double b) {
  Dual<N, S> r; r.v = a.v * S(b); for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * S(b); return r; }
template <int N, class S> inline Dual<N, S> operator*(double b, const Dual<N, S>& a) { return a * b; }
template <int N, class S> inline Dual<N, S> operator/(const Dual<N, S>& a, double b) { return a * (1.0 / b); }
template <int N, class S> inline Dual<N, S> operator/(double b, const Dual<N, S>& a) { return Dual<N, S>(b) / a; }
template <int N, class S> inline Dual<N, S>& operator+=(Dual<N, S>& a, const Dual<N, S>& b) { a = a + b; return a; }
template <int N, class S> inline Dual<N, S>& operator-=(Dual<N, S>& a, const Dual<N, S>& b) { a = a - b; return a; }
template <int N, class S> inline Dual<N, S>& operator*=(Dual<N, S>& a, const Dual<N, S>& b) { a = a * b; return a; }
template <int N, class S> inline Dual<N, S>& operator*=(Dual<N, S>& a, double b) { a = a * b; return a; }

inline double value_of(double x) { return x; }
template <int N, class S> inline double value_of(const Dual<N, S>& x) { return value_of(x.v); }

using std::sin; using std::cos; using std::sqrt; using std::acos; using std::atan2; using std::log;

template <int N, class S> inline Dual<N, S> sin(const Dual<N, S>& a) {
  Dual<N, S> r; r.v = sin(a.v); S c = cos(a.v); for (int i = 0; i < N; ++i) r.d[i] = c * a.d[i]; return r; }
template <int N, class S> inline Dual<N, S> cos(const Dual<N, S>& a) {
  Dual<N, S> r; r.v = cos(a.v); S s = -sin(a.v); for (int i = 0; i < N; ++i) r.d[i] = s * a.d[i]; return r; }
template <int N, class S> inline Dual<N, S> sqrt(const Dual<N, S>& a) {
  Dual<N, S> r; r.v = sqrt(a.v); S h = S(0.5) / r.v; for (int i = 0; i < N; ++i) r.d[i] = h * a.d[i]; return r; }
template <int N, class S> inline Dual<N, S> log(const Dual<N, S>& a) {
  Dual<N, S> r; r.v = log(a.v); S h = S(1.0) / a.v; for (int i = 0; i < N; ++i) r.d[i] = h * a.d[i]; return r; }

template <int N, class S> inline bool operator>(const Dual<N, S>& a, const Dual<N, S>& b) { return value_of(a) > value_of(b); }
template <int N, class S> inline bool operator<(const Dual<N, S>& a, const Dual<N, S>& b) { return value_of(a) < value_of(b); }
template <int N, class S> inline bool operator>(const Dual<N, S>& a, double b) { return value_of(a) > b; }
template <int N, class S> inline bool operator<(const Dual<N, S>& a, double b) { return value_of(a) < b; }
template <int N, class S> inline bool operator>=(const Dual<N, S>& a, double b) { return value_of(a) >= b; }
template <int N, class S> inline bool operator<=(const Dual<N, S>& a, double b) { return value_of(a) <= b; }

// Second-order Taylor scalar along ONE direction: x(t) = v + d1 t + d2 t^2/2.
// Pushing q(t) = q + t*qdot through forward kinematics yields position, velocity (J qdot) and the
// zero-joint-acceleration bias acceleration (Jdot qdot) of every frame in one pass.
struct Jet2 {
  double v, d1, d2;
  Jet2() : v(0), d1(0), d2(0) {}
  Jet2(double c) : v(c), d1(0), d2(0) {}
  Jet2(double a, double b, double c) : v(a), d1(b), d2(c) {}
};
inline Jet2 operator+(const Jet2& a, const Jet2& b) { return {a.v + b.v, a.d1 + b.d1, a.d2 + b.d2}; }
inline Jet2 operator-(const Jet2& a, const Jet2& b) { return {a.v - b.v, a.d1 - b.d1, a.d2 - b.d2}; }
inline Jet2 operator-(const Jet2& a) { return {-a.v, -a.d1, -a.d2}; }
inline Jet2 operator*(const Jet2& a, const Jet2& b) {
  return {a.v * b.v, a.d1 * b.v + a.v * b.d1, a.d2 * b.v + 2.0 * a.d1 * b.d1 + a.v * b.d2}; }
inline Jet2 operator*(const Jet2& a, double b) { return {a.v * b, a.d1 * b, a.d2 * b}; }
inline Jet2 operator*(double b, const Jet2& a) { return a * b; }
inline Jet2 operator+(const Jet2& a, double b) { return {a.v + b, a.d1, a.d2}; }
inline Jet2 operator-(const Jet2& a, double b) { return {a.v - b, a.d1, a.d2}; }
inline Jet2& operator+=(Jet2& a, const Jet2& b) { a = a + b; return a; }
inline Jet2& operator-=(Jet2& a, const Jet2& b) { a = a - b; return a; }
inline Jet2 sin(const Jet2& a) { double s = std::sin(a.v), c = std::cos(a.v); return {s, c * a.d1, c * a.d2 - s * a.d1 * a.d1}; }
inline Jet2 cos(const Jet2& a) { double s = std::sin(a.v), c = std::cos(a.v); return {c, -s * a.d1, -s * a.d2 - c * a.d1 * a.d1}; }
inline double value_of(const Jet2& x) { return x.v; }

}  // namespace orc
