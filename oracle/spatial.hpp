// TEST INFRASTRUCTURE ONLY — part of the fp64 CPU oracle (see oracle/nb_oracle.cpp).
// Scalar-generic spatial algebra restating dart/math/Geometry.cpp.  Spatial vectors are
// [angular; linear] in the body frame (Geometry.cpp:1300-1312).
#pragma once
#include <cmath>
#include <cstring>

namespace orc {

// ------------------------------------------------------------------ Dual numbers
// forward-mode AD carrier: value + N directional derivatives
template <int N>
struct Dual {
  double v;
  double d[N];
  Dual() : v(0) { for (int i = 0; i < N; i++) d[i] = 0; }
  Dual(double x) : v(x) { for (int i = 0; i < N; i++) d[i] = 0; }
};
template <int N> inline Dual<N> operator+(const Dual<N>& a, const Dual<N>& b) { Dual<N> r; r.v = a.v + b.v; for (int i = 0; i < N; i++) r.d[i] = a.d[i] + b.d[i]; return r; }
template <int N> inline Dual<N> operator-(const Dual<N>& a, const Dual<N>& b) { Dual<N> r; r.v = a.v - b.v; for (int i = 0; i < N; i++) r.d[i] = a.d[i] - b.d[i]; return r; }
template <int N> inline Dual<N> operator-(const Dual<N>& a) { Dual<N> r; r.v = -a.v; for (int i = 0; i < N; i++) r.d[i] = -a.d[i]; return r; }
template <int N> inline Dual<N> operator*(const Dual<N>& a, const Dual<N>& b) { Dual<N> r; r.v = a.v * b.v; for (int i = 0; i < N; i++) r.d[i] = a.d[i] * b.v + a.v * b.d[i]; return r; }
template <int N> inline Dual<N> operator/(const Dual<N>& a, const Dual<N>& b) { Dual<N> r; double ib = 1.0 / b.v; r.v = a.v * ib; for (int i = 0; i < N; i++) r.d[i] = (a.d[i] - r.v * b.d[i]) * ib; return r; }
template <int N> inline Dual<N> operator+(const Dual<N>& a, double b) { Dual<N> r = a; r.v += b; return r; }
template <int N> inline Dual<N> operator+(double b, const Dual<N>& a) { return a + b; }
template <int N> inline Dual<N> operator-(const Dual<N>& a, double b) { Dual<N> r = a; r.v -= b; return r; }
template <int N> inline Dual<N> operator-(double b, const Dual<N>& a) { return (-a) + b; }
template <int N> inline Dual<N> operator*(const Dual<N>& a, double b) { Dual<N> r; r.v = a.v * b; for (int i = 0; i < N; i++) r.d[i] = a.d[i] * b; return r; }
template <int N> inline Dual<N> operator*(double b, const Dual<N>& a) { return a * b; }
template <int N> inline Dual<N> operator/(const Dual<N>& a, double b) { return a * (1.0 / b); }
template <int N> inline Dual<N> operator/(double a, const Dual<N>& b) { return Dual<N>(a) / b; }
template <int N> inline Dual<N>& operator+=(Dual<N>& a, const Dual<N>& b) { a = a + b; return a; }
template <int N> inline Dual<N>& operator-=(Dual<N>& a, const Dual<N>& b) { a = a - b; return a; }
template <int N> inline Dual<N>& operator*=(Dual<N>& a, const Dual<N>& b) { a = a * b; return a; }
template <int N> inline Dual<N>& operator*=(Dual<N>& a, double b) { a = a * b; return a; }
template <int N> inline Dual<N> sin(const Dual<N>& a) { Dual<N> r; r.v = std::sin(a.v); double c = std::cos(a.v); for (int i = 0; i < N; i++) r.d[i] = c * a.d[i]; return r; }
template <int N> inline Dual<N> cos(const Dual<N>& a) { Dual<N> r; r.v = std::cos(a.v); double s = -std::sin(a.v); for (int i = 0; i < N; i++) r.d[i] = s * a.d[i]; return r; }
template <int N> inline Dual<N> sqrt(const Dual<N>& a) { Dual<N> r; r.v = std::sqrt(a.v); double k = (r.v > 0) ? 0.5 / r.v : 0.0; for (int i = 0; i < N; i++) r.d[i] = k * a.d[i]; return r; }
template <int N> inline Dual<N> acos(const Dual<N>& a) { Dual<N> r; r.v = std::acos(a.v); double s = 1.0 - a.v * a.v; double k = (s > 0) ? -1.0 / std::sqrt(s) : 0.0; for (int i = 0; i < N; i++) r.d[i] = k * a.d[i]; return r; }
inline double val(double x) { return x; }
inline double val(float x) { return (double)x; }
template <int N> inline double val(const Dual<N>& x) { return x.v; }
using std::sin; using std::cos; using std::sqrt; using std::acos;

// ------------------------------------------------------------------ small fixed types
template <class S> struct Vec3 { S x[3]; S& operator[](int i) { return x[i]; } const S& operator[](int i) const { return x[i]; } };
template <class S> struct Vec6 { S x[6]; S& operator[](int i) { return x[i]; } const S& operator[](int i) const { return x[i]; } };
template <class S> struct Mat3 { S m[9]; S& operator()(int r, int c) { return m[3 * r + c]; } const S& operator()(int r, int c) const { return m[3 * r + c]; } };
template <class S> struct Mat6 { S m[36]; S& operator()(int r, int c) { return m[6 * r + c]; } const S& operator()(int r, int c) const { return m[6 * r + c]; } };
template <class S> struct Iso { Mat3<S> R; Vec3<S> p; };  // x_parent = R x_child + p

template <class S> inline Vec3<S> v3(S a, S b, S c) { Vec3<S> r; r[0] = a; r[1] = b; r[2] = c; return r; }
template <class S> inline Vec3<S> operator+(const Vec3<S>& a, const Vec3<S>& b) { return v3<S>(a[0] + b[0], a[1] + b[1], a[2] + b[2]); }
template <class S> inline Vec3<S> operator-(const Vec3<S>& a, const Vec3<S>& b) { return v3<S>(a[0] - b[0], a[1] - b[1], a[2] - b[2]); }
template <class S> inline Vec3<S> operator*(const Vec3<S>& a, const S& s) { return v3<S>(a[0] * s, a[1] * s, a[2] * s); }
template <class S> inline Vec3<S> neg(const Vec3<S>& a) { return v3<S>(-a[0], -a[1], -a[2]); }
template <class S> inline S dot(const Vec3<S>& a, const Vec3<S>& b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
template <class S> inline Vec3<S> cross(const Vec3<S>& a, const Vec3<S>& b) { return v3<S>(a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]); }
template <class S> inline Vec3<S> mul(const Mat3<S>& M, const Vec3<S>& v) { Vec3<S> r; for (int i = 0; i < 3; i++) r[i] = M(i, 0) * v[0] + M(i, 1) * v[1] + M(i, 2) * v[2]; return r; }
template <class S> inline Vec3<S> mulT(const Mat3<S>& M, const Vec3<S>& v) { Vec3<S> r; for (int i = 0; i < 3; i++) r[i] = M(0, i) * v[0] + M(1, i) * v[1] + M(2, i) * v[2]; return r; }
template <class S> inline Mat3<S> mul(const Mat3<S>& A, const Mat3<S>& B) { Mat3<S> C; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) C(i, j) = A(i, 0) * B(0, j) + A(i, 1) * B(1, j) + A(i, 2) * B(2, j); return C; }
template <class S> inline Mat3<S> transpose(const Mat3<S>& A) { Mat3<S> C; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) C(i, j) = A(j, i); return C; }
template <class S> inline Mat3<S> eye3() { Mat3<S> I; for (int i = 0; i < 9; i++) I.m[i] = S(0.0); I(0, 0) = I(1, 1) = I(2, 2) = S(1.0); return I; }
template <class S> inline Mat3<S> skew(const Vec3<S>& v) { Mat3<S> K; for (int i = 0; i < 9; i++) K.m[i] = S(0.0); K(0, 1) = -v[2]; K(0, 2) = v[1]; K(1, 0) = v[2]; K(1, 2) = -v[0]; K(2, 0) = -v[1]; K(2, 1) = v[0]; return K; }
template <class S> inline Vec3<S> head(const Vec6<S>& v) { return v3<S>(v[0], v[1], v[2]); }
template <class S> inline Vec3<S> tail(const Vec6<S>& v) { return v3<S>(v[3], v[4], v[5]); }
template <class S> inline Vec6<S> v6(const Vec3<S>& a, const Vec3<S>& b) { Vec6<S> r; for (int i = 0; i < 3; i++) { r[i] = a[i]; r[3 + i] = b[i]; } return r; }
template <class S> inline Vec6<S> zero6() { Vec6<S> r; for (int i = 0; i < 6; i++) r[i] = S(0.0); return r; }
template <class S> inline Vec6<S> operator+(const Vec6<S>& a, const Vec6<S>& b) { Vec6<S> r; for (int i = 0; i < 6; i++) r[i] = a[i] + b[i]; return r; }
template <class S> inline Vec6<S> operator-(const Vec6<S>& a, const Vec6<S>& b) { Vec6<S> r; for (int i = 0; i < 6; i++) r[i] = a[i] - b[i]; return r; }
template <class S> inline Vec6<S> operator*(const Vec6<S>& a, const S& s) { Vec6<S> r; for (int i = 0; i < 6; i++) r[i] = a[i] * s; return r; }
template <class S> inline S dot(const Vec6<S>& a, const Vec6<S>& b) { S r = a[0] * b[0]; for (int i = 1; i < 6; i++) r = r + a[i] * b[i]; return r; }
template <class S> inline Vec6<S> mul(const Mat6<S>& M, const Vec6<S>& v) { Vec6<S> r; for (int i = 0; i < 6; i++) { S a = M(i, 0) * v[0]; for (int j = 1; j < 6; j++) a = a + M(i, j) * v[j]; r[i] = a; } return r; }
template <class S> inline Mat6<S> zero66() { Mat6<S> r; for (int i = 0; i < 36; i++) r.m[i] = S(0.0); return r; }
template <class S> inline Mat6<S> operator+(const Mat6<S>& a, const Mat6<S>& b) { Mat6<S> r; for (int i = 0; i < 36; i++) r.m[i] = a.m[i] + b.m[i]; return r; }

template <class S> inline Iso<S> iso_identity() { Iso<S> T; T.R = eye3<S>(); T.p = v3<S>(S(0.0), S(0.0), S(0.0)); return T; }
template <class S> inline Iso<S> mul(const Iso<S>& A, const Iso<S>& B) { Iso<S> C; C.R = mul(A.R, B.R); C.p = mul(A.R, B.p) + A.p; return C; }
template <class S> inline Iso<S> inverse(const Iso<S>& A) { Iso<S> C; C.R = transpose(A.R); C.p = neg(mul(C.R, A.p)); return C; }
template <class S> inline Vec3<S> apply(const Iso<S>& A, const Vec3<S>& x) { return mul(A.R, x) + A.p; }

// ------------------------------------------------------------------ SE(3) operators (dart/math/Geometry.cpp)
// AdT  (:1302-1312):  w' = R w ; v' = p x (R w) + R v
template <class S> inline Vec6<S> AdT(const Iso<S>& T, const Vec6<S>& V) {
  Vec3<S> w = mul(T.R, head(V));
  Vec3<S> v = mul(T.R, tail(V)) + cross(T.p, w);
  return v6(w, v);
}
// AdInvT (:1437-1445): w' = R^T w ; v' = R^T (v + w x p)
template <class S> inline Vec6<S> AdInvT(const Iso<S>& T, const Vec6<S>& V) {
  Vec3<S> w = mulT(T.R, head(V));
  Vec3<S> v = mulT(T.R, tail(V) + cross(head(V), T.p));
  return v6(w, v);
}
// ad (:1470-1483)
template <class S> inline Vec6<S> ad(const Vec6<S>& X, const Vec6<S>& Y) {
  return v6(cross(head(X), head(Y)), cross(head(X), tail(Y)) + cross(tail(X), head(Y)));
}
// dad (:3506-3513): head = t.h x s.h + t.t x s.t ; tail = t.t x s.h
template <class S> inline Vec6<S> dad(const Vec6<S>& s, const Vec6<S>& t) {
  return v6(cross(head(t), head(s)) + cross(tail(t), tail(s)), cross(tail(t), head(s)));
}
// dAdT (:1504-1512): head = R^T (m + f x p) ; tail = R^T f
template <class S> inline Vec6<S> dAdT(const Iso<S>& T, const Vec6<S>& F) {
  return v6(mulT(T.R, head(F) + cross(tail(F), T.p)), mulT(T.R, tail(F)));
}
// dAdInvT (:1529-1537): tail = R f ; head = R m + p x (R f)
template <class S> inline Vec6<S> dAdInvT(const Iso<S>& T, const Vec6<S>& F) {
  Vec3<S> f = mul(T.R, tail(F));
  Vec3<S> m = mul(T.R, head(F)) + cross(T.p, f);
  return v6(m, f);
}
// 6x6 matrix of AdT (getAdTMatrix :1316-1330): [[R,0],[ [p]x R, R ]]
template <class S> inline Mat6<S> AdTMatrix(const Iso<S>& T) {
  Mat6<S> A = zero66<S>();
  Mat3<S> pR = mul(skew(T.p), T.R);
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
    A(i, j) = T.R(i, j); A(3 + i, 3 + j) = T.R(i, j); A(3 + i, j) = pR(i, j);
  }
  return A;
}
// transformInertia(T, I) (:3515-3597) == AdT(T)^T * I * AdT(T)  (dense form; same value: e.g. its
// lower-right block is R^T C R, :3576-3590).  Called with T = (child relative transform)^-1 it moves a child's
// articulated inertia into the parent frame.
template <class S> inline Mat6<S> transformInertia(const Iso<S>& T, const Mat6<S>& I) {
  Mat6<S> A = AdTMatrix(T);
  Mat6<S> IA = zero66<S>(), R = zero66<S>();
  for (int i = 0; i < 6; i++) for (int j = 0; j < 6; j++) { S a = S(0.0); for (int k = 0; k < 6; k++) a = a + I(i, k) * A(k, j); IA(i, j) = a; }
  for (int i = 0; i < 6; i++) for (int j = 0; j < 6; j++) { S a = S(0.0); for (int k = 0; k < 6; k++) a = a + A(k, i) * IA(k, j); R(i, j) = a; }
  return R;
}

// expMapRot (:539-554), EPSILON_EXPMAP_THETA = 1e-3 (Geometry.cpp top)
template <class S> inline Mat3<S> expMapRot(const Vec3<S>& q) {
  S th2 = dot(q, q);
  Mat3<S> K = skew(q), K2 = mul(K, K), R = eye3<S>();
  S a, b;
  if (val(th2) < 1e-6) { a = S(1.0); b = S(0.5); }
  else { S th = sqrt(th2); a = sin(th) / th; b = (S(1.0) - cos(th)) / th2; }
  for (int i = 0; i < 9; i++) R.m[i] = R.m[i] + a * K.m[i] + b * K2.m[i];
  return R;
}
// logMap (:720-760), DART_EPSILON = 1e-6
template <class S> inline Vec3<S> logMap(const Mat3<S>& R) {
  S c = S(0.5) * (R(0, 0) + R(1, 1) + R(2, 2) - S(1.0));
  if (val(c) > 1.0) c = c - (c - S(1.0));  // clamp keeping derivative bookkeeping finite
  if (val(c) < -1.0) c = c - (c + S(1.0));
  S theta = acos(c);
  const double PI = 3.14159265358979323846;
  if (val(theta) > PI - 1e-6) {
    S delta = S(0.5) + S(0.125) * (S(PI) - theta) * (S(PI) - theta);
    S a0 = theta * sqrt(S(1.0) + (R(0, 0) - S(1.0)) * delta);
    S a1 = theta * sqrt(S(1.0) + (R(1, 1) - S(1.0)) * delta);
    S a2 = theta * sqrt(S(1.0) + (R(2, 2) - S(1.0)) * delta);
    return v3<S>(val(R(2, 1)) > val(R(1, 2)) ? a0 : -a0, val(R(0, 2)) > val(R(2, 0)) ? a1 : -a1,
                 val(R(1, 0)) > val(R(0, 1)) ? a2 : -a2);
  }
  S alpha;
  if (val(theta) > 1e-6) alpha = S(0.5) * theta / sin(theta);
  else alpha = S(0.5) + S(1.0 / 12.0) * theta * theta;
  return v3<S>(alpha * (R(2, 1) - R(1, 2)), alpha * (R(0, 2) - R(2, 0)), alpha * (R(1, 0) - R(0, 1)));
}

}  // namespace orc
