// Contact geometry, scalar-generic: the same code runs on plain doubles (forward contact stage) and on small dual
// numbers carrying d/d(xi) for the 6 directions of a body-pose perturbation (backward: derivative of the contact
// point / normal with respect to the pose of the moving body, whatever the contact type).
// Restates  collideBoxSphere / collideSphereBox   dart/collision/dart/DARTCollide.cpp:1482-1653, 1655-1810
//           dBoxBox + intersectRectQuad             DARTCollide.cpp:764-1450, 513-580, dLineClosestApproach :270-298
#pragma once
#include "nb2_math.cuh"

namespace nb2 {

typedef double CR;

// ---- forward-mode dual number with N directions
template <int N> struct DualT {
  double v; double d[N];
  NB2_HD DualT() {}
  NB2_HD DualT(double x) : v(x) { for (int i = 0; i < N; i++) d[i] = 0; }
};
template <int N> NB2_HD DualT<N> operator+(const DualT<N>& a, const DualT<N>& b) { DualT<N> r; r.v = a.v + b.v; for (int i = 0; i < N; i++) r.d[i] = a.d[i] + b.d[i]; return r; }
template <int N> NB2_HD DualT<N> operator-(const DualT<N>& a, const DualT<N>& b) { DualT<N> r; r.v = a.v - b.v; for (int i = 0; i < N; i++) r.d[i] = a.d[i] - b.d[i]; return r; }
template <int N> NB2_HD DualT<N> operator-(const DualT<N>& a) { DualT<N> r; r.v = -a.v; for (int i = 0; i < N; i++) r.d[i] = -a.d[i]; return r; }
template <int N> NB2_HD DualT<N> operator*(const DualT<N>& a, const DualT<N>& b) { DualT<N> r; r.v = a.v * b.v; for (int i = 0; i < N; i++) r.d[i] = a.d[i] * b.v + a.v * b.d[i]; return r; }
template <int N> NB2_HD DualT<N> operator/(const DualT<N>& a, const DualT<N>& b) { DualT<N> r; const double ib = 1.0 / b.v; r.v = a.v * ib; for (int i = 0; i < N; i++) r.d[i] = (a.d[i] - r.v * b.d[i]) * ib; return r; }
template <int N> NB2_HD DualT<N>& operator+=(DualT<N>& a, const DualT<N>& b) { a = a + b; return a; }
template <int N> NB2_HD DualT<N>& operator-=(DualT<N>& a, const DualT<N>& b) { a = a - b; return a; }
template <int N> NB2_HD DualT<N>& operator*=(DualT<N>& a, const DualT<N>& b) { a = a * b; return a; }
template <int N> NB2_HD DualT<N>& operator/=(DualT<N>& a, const DualT<N>& b) { a = a / b; return a; }
template <int N> NB2_HD DualT<N> nb2_sqrt(const DualT<N>& a) { DualT<N> r; r.v = sqrt(a.v); const double k = r.v > 0 ? 0.5 / r.v : 0.0; for (int i = 0; i < N; i++) r.d[i] = k * a.d[i]; return r; }
template <int N> NB2_HD DualT<N> nb2_abs(const DualT<N>& a) { return a.v < 0 ? -a : a; }
NB2_HD double gval(double x) { return x; }
template <int N> NB2_HD double gval(const DualT<N>& x) { return x.v; }

template <class T> NB2_HD Xf<T> gxf_mul(const Xf<T>& A, const Xf<T>& B) { Xf<T> C; C.R_ = mul(A.R_, B.R_); C.p = mul(A.R_, B.p) + A.p; return C; }
template <class T> NB2_HD V3<T> gxf_apply(const Xf<T>& A, const V3<T>& x) { return mul(A.R_, x) + A.p; }
template <class T> NB2_HD V3<T> gxf_apply_inv(const Xf<T>& A, const V3<T>& x) { return mulT(A.R_, x - A.p); }
template <class T> NB2_HD V3<T> gcol3(const M3<T>& R, int j) { return j == 0 ? mk3<T>(R.m00, R.m10, R.m20) : (j == 1 ? mk3<T>(R.m01, R.m11, R.m21) : mk3<T>(R.m02, R.m12, R.m22)); }
template <class T> NB2_HD T gget3(const V3<T>& v, int k) { return k == 0 ? v.x : (k == 1 ? v.y : v.z); }
template <class T> NB2_HD void gset3(V3<T>& v, int k, const T& x) { if (k == 0) v.x = x; else if (k == 1) v.y = x; else v.z = x; }

template <class T> struct ContactOutT { V3<T> point, normal; T depth; int type; };

// contact types emitted (subset of collision::ContactType, dart/collision/Contact.hpp:50-80)
//   1 VERTEX_FACE  2 FACE_VERTEX  3 EDGE_EDGE  4 SPHERE_BOX  5 BOX_SPHERE

// ---- box vs sphere.  sphere_first=false: object 1 = box, object 2 = sphere (DARTCollide.cpp:1482-1653, normal = contact
// point - centre); sphere_first=true: object 1 = sphere (:1655-1810, normal = centre - contact point, halfspace ignored)
template <class T>
NB2_HD int collide_box_sphere(const V3<T>& size0, const Xf<T>& T0, const T& r1, const Xf<T>& T1, CR clip, int halfspace, bool sphere_first,
                              ContactOutT<T>* out) {
  const V3<T> half = size0 * T(0.5);
  bool inside = true;
  const V3<T> c0 = T1.p;
  V3<T> p = gxf_apply_inv(T0, c0);
  for (int k = 0; k < 3; k++) {
    const T pk = gget3(p, k), hk = gget3(half, k);
    if (gval(pk) < -gval(hk)) { gset3(p, k, -hk); inside = false; }
    if (gval(pk) > gval(hk)) { gset3(p, k, hk); inside = false; }
  }
  T mn = half.x - nb2_abs(p.x); int idx = 0;
  T t = half.y - nb2_abs(p.y); if (gval(t) < gval(mn)) { mn = t; idx = 1; }
  t = half.z - nb2_abs(p.z); if (gval(t) < gval(mn)) { mn = t; idx = 2; }
  V3<T> nloc = mk3<T>(T(0.0), T(0.0), T(0.0));
  const double sgn = (gval(gget3(p, idx)) > 0.0) ? 1.0 : -1.0;
  gset3(nloc, idx, T(sphere_first ? sgn : -sgn));
  const V3<T> nface = mul(T0.R_, nloc);
  if (inside) {
    const T pen = mn + r1;
    if (gval(pen) > clip) return 0;
    out->type = sphere_first ? 1 : 2; out->point = c0; out->normal = nface; out->depth = pen; return 1;
  }
  const V3<T> cp = gxf_apply(T0, p);
  const V3<T> n = sphere_first ? (c0 - cp) : (cp - c0);
  const T mag = nb2_sqrt(dot(n, n));
  const T pen = r1 - mag;
  if (gval(pen) > clip) return 0;
  if (!sphere_first) {
    const double lz = gval(gxf_apply_inv(T1, cp).z);
    if (halfspace == 2 && lz >= 0) return 0;
    if (halfspace == 1 && lz <= 0) return 0;
  }
  if (gval(pen) < 0.0) return 0;
  out->type = sphere_first ? 4 : 5; out->point = cp; out->depth = pen;
  out->normal = (gval(mag) > 1e-6) ? n * (T(1.0) / mag) : nface;
  return 1;
}

template <class T>
NB2_HD int intersect_rect_quad(const T h[2], T p[8], T ret[16]) {
  int nq = 4, nr = 0;
  T buffer[16];
  T* q = p; T* r = ret;
  for (int dir = 0; dir <= 1; dir++) {
    for (int sign = -1; sign <= 1; sign += 2) {
      T* pq = q; T* pr = r; nr = 0;
      for (int i = nq; i > 0; i--) {
        if (sign * gval(pq[dir]) < gval(h[dir])) { pr[0] = pq[0]; pr[1] = pq[1]; pr += 2; nr++; if (nr & 8) { q = r; goto done; } }
        T* nextq = (i > 1) ? pq + 2 : q;
        if ((sign * gval(pq[dir]) < gval(h[dir])) ^ (sign * gval(nextq[dir]) < gval(h[dir]))) {
          pr[1 - dir] = pq[1 - dir] + (nextq[1 - dir] - pq[1 - dir]) / (nextq[dir] - pq[dir]) * (T((double)sign) * h[dir] - pq[dir]);
          pr[dir] = T((double)sign) * h[dir];
          pr += 2; nr++;
          if (nr & 8) { q = r; goto done; }
        }
        pq += 2;
      }
      q = r; r = (q == ret) ? buffer : ret; nq = nr;
    }
  }
done:
  if (q != ret) for (int i = 0; i < nr * 2; i++) ret[i] = q[i];
  return nr;
}

// dBoxBox; returns the number of contacts written to out (<= 8)
template <class T>
NB2_HD int collide_box_box(const V3<T>& size0, const Xf<T>& T0, const V3<T>& size1, const Xf<T>& T1, CR clip, ContactOutT<T>* out) {
  const double fudge = 1.05;
  const M3<T>&R1 = T0.R_, &R2 = T1.R_;
  const V3<T> p1 = T0.p, p2 = T1.p;
  const T A[3] = {size0.x * T(0.5), size0.y * T(0.5), size0.z * T(0.5)}, Bh[3] = {size1.x * T(0.5), size1.y * T(0.5), size1.z * T(0.5)};
  const V3<T> p = p2 - p1;
  const V3<T> ppv = mulT(R1, p);
  const T pp[3] = {ppv.x, ppv.y, ppv.z};
  T Rm[3][3], Q[3][3];
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { Rm[i][j] = dot(gcol3(R1, i), gcol3(R2, j)); Q[i][j] = nb2_abs(Rm[i][j]); }
  T s = T(-1e12), s2;
  int invert_normal = 0, code = 0, nbox = 0, ncol = -1;
  V3<T> normalC = mk3<T>(T(0.0), T(0.0), T(0.0));
#define NB2_TST(expr1, expr2, box, colj, cc) { const T e1 = (expr1); s2 = nb2_abs(e1) - (expr2); if (gval(s2) > gval(s)) { s = s2; nbox = box; ncol = colj; invert_normal = (gval(e1) < 0); code = (cc); } }
  NB2_TST(pp[0], (A[0] + Bh[0] * Q[0][0] + Bh[1] * Q[0][1] + Bh[2] * Q[0][2]), 1, 0, 1)
  NB2_TST(pp[1], (A[1] + Bh[0] * Q[1][0] + Bh[1] * Q[1][1] + Bh[2] * Q[1][2]), 1, 1, 2)
  NB2_TST(pp[2], (A[2] + Bh[0] * Q[2][0] + Bh[1] * Q[2][1] + Bh[2] * Q[2][2]), 1, 2, 3)
  NB2_TST(dot(gcol3(R2, 0), p), (A[0] * Q[0][0] + A[1] * Q[1][0] + A[2] * Q[2][0] + Bh[0]), 2, 0, 4)
  NB2_TST(dot(gcol3(R2, 1), p), (A[0] * Q[0][1] + A[1] * Q[1][1] + A[2] * Q[2][1] + Bh[1]), 2, 1, 5)
  NB2_TST(dot(gcol3(R2, 2), p), (A[0] * Q[0][2] + A[1] * Q[1][2] + A[2] * Q[2][2] + Bh[2]), 2, 2, 6)
#undef NB2_TST
#define NB2_TST2(expr1, expr2, n1, n2, n3, cc) { const T e1 = (expr1); s2 = nb2_abs(e1) - (expr2); const T N1 = (n1), N2 = (n2), N3 = (n3); const T l = nb2_sqrt(N1 * N1 + N2 * N2 + N3 * N3); \
    if (gval(l) > 0) { s2 = s2 / l; if (gval(s2) * fudge > gval(s)) { s = s2; ncol = -1; normalC = mk3<T>(N1 / l, N2 / l, N3 / l); invert_normal = (gval(e1) < 0); code = (cc); } } }
  const T Z = T(0.0);
  NB2_TST2(pp[2] * Rm[1][0] - pp[1] * Rm[2][0], (A[1] * Q[2][0] + A[2] * Q[1][0] + Bh[1] * Q[0][2] + Bh[2] * Q[0][1]), Z, -Rm[2][0], Rm[1][0], 7)
  NB2_TST2(pp[2] * Rm[1][1] - pp[1] * Rm[2][1], (A[1] * Q[2][1] + A[2] * Q[1][1] + Bh[0] * Q[0][2] + Bh[2] * Q[0][0]), Z, -Rm[2][1], Rm[1][1], 8)
  NB2_TST2(pp[2] * Rm[1][2] - pp[1] * Rm[2][2], (A[1] * Q[2][2] + A[2] * Q[1][2] + Bh[0] * Q[0][1] + Bh[1] * Q[0][0]), Z, -Rm[2][2], Rm[1][2], 9)
  NB2_TST2(pp[0] * Rm[2][0] - pp[2] * Rm[0][0], (A[0] * Q[2][0] + A[2] * Q[0][0] + Bh[1] * Q[1][2] + Bh[2] * Q[1][1]), Rm[2][0], Z, -Rm[0][0], 10)
  NB2_TST2(pp[0] * Rm[2][1] - pp[2] * Rm[0][1], (A[0] * Q[2][1] + A[2] * Q[0][1] + Bh[0] * Q[1][2] + Bh[2] * Q[1][0]), Rm[2][1], Z, -Rm[0][1], 11)
  NB2_TST2(pp[0] * Rm[2][2] - pp[2] * Rm[0][2], (A[0] * Q[2][2] + A[2] * Q[0][2] + Bh[0] * Q[1][1] + Bh[1] * Q[1][0]), Rm[2][2], Z, -Rm[0][2], 12)
  NB2_TST2(pp[1] * Rm[0][0] - pp[0] * Rm[1][0], (A[0] * Q[1][0] + A[1] * Q[0][0] + Bh[1] * Q[2][2] + Bh[2] * Q[2][1]), -Rm[1][0], Rm[0][0], Z, 13)
  NB2_TST2(pp[1] * Rm[0][1] - pp[0] * Rm[1][1], (A[0] * Q[1][1] + A[1] * Q[0][1] + Bh[0] * Q[2][2] + Bh[2] * Q[2][0]), -Rm[1][1], Rm[0][1], Z, 14)
  NB2_TST2(pp[1] * Rm[0][2] - pp[0] * Rm[1][2], (A[0] * Q[1][2] + A[1] * Q[0][2] + Bh[0] * Q[2][1] + Bh[1] * Q[2][0]), -Rm[1][2], Rm[0][2], Z, 15)
#undef NB2_TST2
  if (!code) return 0;
  if (gval(s) > 0.0) return 0;
  V3<T> normal;
  if (ncol >= 0) normal = gcol3(nbox == 1 ? R1 : R2, ncol);
  else { normal = mul(R1, normalC); normal = normal * (T(1.0) / nb2_sqrt(dot(normal, normal))); }
  if (invert_normal) normal = -normal;
  if (code > 6) {
    V3<T> pa = p1, pb = p2;
    for (int j = 0; j < 3; j++) { const double sg = (gval(dot(normal, gcol3(R1, j))) > -1e-10) ? 1.0 : -1.0; pa = pa + gcol3(R1, j) * (A[j] * T(sg)); }
    for (int j = 0; j < 3; j++) { const double sg = (gval(dot(normal, gcol3(R2, j))) > -1e-3) ? -1.0 : 1.0; pb = pb + gcol3(R2, j) * (Bh[j] * T(sg)); }
    const V3<T> ua = gcol3(R1, (code - 7) / 3), ub = gcol3(R2, (code - 7) % 3);
    const V3<T> dp = pb - pa;
    const T uaub = dot(ua, ub), q1 = dot(ua, dp), q2 = -dot(ub, dp);
    T d = T(1.0) - uaub * uaub, alpha = T(0.0), beta = T(0.0);
    if (gval(d) > 0.0) { d = T(1.0) / d; alpha = (q1 + uaub * q2) * d; beta = (uaub * q1 + q2) * d; }
    pa = pa + ua * alpha; pb = pb + ub * beta;
    const T pen = -s;
    if (gval(pen) > clip) return 0;
    out[0].point = (pa + pb) * T(0.5); out[0].normal = -normal; out[0].depth = pen; out[0].type = 3;
    return 1;
  }
  const M3<T>*Ra, *Rb; V3<T> pa, pb; const T *Sa, *Sb; bool flip;
  if (code <= 3) { Ra = &R1; Rb = &R2; pa = p1; pb = p2; Sa = A; Sb = Bh; flip = false; }
  else { Ra = &R2; Rb = &R1; pa = p2; pb = p1; Sa = Bh; Sb = A; flip = true; }
  const V3<T> normal2 = (code <= 3) ? normal : -normal;
  const V3<T> nr = mulT(*Rb, normal2);
  const double anr[3] = {fabs(gval(nr.x)), fabs(gval(nr.y)), fabs(gval(nr.z))};
  int lanr, a1, a2;
  if (anr[1] > anr[0]) { if (anr[1] > anr[2]) { a1 = 0; lanr = 1; a2 = 2; } else { a1 = 0; a2 = 1; lanr = 2; } }
  else { if (anr[0] > anr[2]) { lanr = 0; a1 = 1; a2 = 2; } else { a1 = 0; a2 = 1; lanr = 2; } }
  const V3<T> center = (gval(gget3(nr, lanr)) < 0) ? (pb - pa + gcol3(*Rb, lanr) * Sb[lanr]) : (pb - pa - gcol3(*Rb, lanr) * Sb[lanr]);
  const int codeN = (code <= 3) ? code - 1 : code - 4;
  int code1, code2;
  if (codeN == 0) { code1 = 1; code2 = 2; } else if (codeN == 1) { code1 = 0; code2 = 2; } else { code1 = 0; code2 = 1; }
  T quad[8];
  const T c1 = dot(center, gcol3(*Ra, code1)), c2 = dot(center, gcol3(*Ra, code2));
  T m11 = dot(gcol3(*Ra, code1), gcol3(*Rb, a1)), m12 = dot(gcol3(*Ra, code1), gcol3(*Rb, a2));
  T m21 = dot(gcol3(*Ra, code2), gcol3(*Rb, a1)), m22 = dot(gcol3(*Ra, code2), gcol3(*Rb, a2));
  {
    const T k1 = m11 * Sb[a1], k2 = m21 * Sb[a1], k3 = m12 * Sb[a2], k4 = m22 * Sb[a2];
    quad[0] = c1 - k1 - k3; quad[1] = c2 - k2 - k4; quad[2] = c1 - k1 + k3; quad[3] = c2 - k2 + k4;
    quad[4] = c1 + k1 + k3; quad[5] = c2 + k2 + k4; quad[6] = c1 + k1 - k3; quad[7] = c2 + k2 - k4;
  }
  const T rect[2] = {Sa[code1], Sa[code2]};
  T ret[16];
  const int n = intersect_rect_quad(rect, quad, ret);
  if (n < 1) return 0;
  const T det1 = T(1.0) / (m11 * m22 - m12 * m21);
  m11 = m11 * det1; m12 = m12 * det1; m21 = m21 * det1; m22 = m22 * det1;
  int cnum = 0;
  for (int j = 0; j < n; j++) {
    const T k1 = m22 * (ret[j * 2] - c1) - m12 * (ret[j * 2 + 1] - c2);
    const T k2 = -m21 * (ret[j * 2] - c1) + m11 * (ret[j * 2 + 1] - c2);
    const V3<T> pt = center + gcol3(*Rb, a1) * k1 + gcol3(*Rb, a2) * k2;
    const T dep = Sa[codeN] - dot(normal2, pt);
    if (gval(dep) >= 0) {
      ContactOutT<T>& c = out[cnum];
      c.point = pt + pa; c.normal = -normal; c.depth = dep;
      const bool onX = fabs(gval(ret[j * 2])) == gval(rect[0]), onY = fabs(gval(ret[j * 2 + 1])) == gval(rect[1]);
      if (onX && onY) {
        if (flip) { c.type = 2; c.point = c.point + c.normal * c.depth; } else { c.type = 1; c.point = c.point - c.normal * c.depth; }
      } else if (!onX && !onY) c.type = flip ? 1 : 2;
      else c.type = 3;
      cnum++;
    }
  }
  return cnum;
}

// ODE tangent basis (ContactConstraint.cpp:734-795, first frictional direction = UnitZ)
template <class T> NB2_HD void tangent_basis(const V3<T>& n, V3<T>* t1, V3<T>* t2) {
  V3<T> t = cross(mk3<T>(T(0.0), T(0.0), T(1.0)), n);
  if (gval(dot(t, t)) < 1e-12) { t = cross(mk3<T>(T(1.0), T(0.0), T(0.0)), n); if (gval(dot(t, t)) < 1e-12) { t = cross(mk3<T>(T(0.0), T(1.0), T(0.0)), n); if (gval(dot(t, t)) < 1e-12) t = cross(mk3<T>(T(0.0), T(0.0), T(1.0)), n); } }
  *t1 = t * (T(1.0) / nb2_sqrt(dot(t, t)));
  *t2 = cross(n, *t1);
}

}  // namespace nb2
