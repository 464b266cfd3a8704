// Contact geometry, scalar-generic: the same code runs on plain doubles (forward contact stage) and on small dual
// numbers carrying d/d(xi) for the 6 directions of a body-pose perturbation (backward: derivative of the contact
// point / normal with respect to the pose of the moving body, whatever the contact type).
// Restates  collideBoxSphere / collideSphereBox   dart/collision/dart/DARTCollide.cpp:1482-1653, 1655-1810
//           dBoxBox (ODE, Russell Smith; OBB test after Gottschalk) DARTCollide.cpp:764-1450, its rectangle clipper :513-580,
//           closest points of two lines :270-298
#pragma once
#include "nb2_math.cuh"

#if defined(__CUDACC__) && !defined(NB2_CW_NOINLINE)
#define NB2_HDG __host__ __device__ __forceinline__
#elif defined(__CUDACC__)
#define NB2_HDG __host__ __device__ __noinline__
#else
#define NB2_HDG inline
#endif

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
template <int N> NB2_HD DualT<N> operator/(const DualT<N>& a, const DualT<N>& b) { DualT<N> r; const double ib = nb2_rcp(b.v); r.v = a.v * ib; for (int i = 0; i < N; i++) r.d[i] = (a.d[i] - r.v * b.d[i]) * ib; return r; }
template <int N> NB2_HD DualT<N>& operator+=(DualT<N>& a, const DualT<N>& b) { a = a + b; return a; }
template <int N> NB2_HD DualT<N>& operator-=(DualT<N>& a, const DualT<N>& b) { a = a - b; return a; }
template <int N> NB2_HD DualT<N>& operator*=(DualT<N>& a, const DualT<N>& b) { a = a * b; return a; }
template <int N> NB2_HD DualT<N> gdiv(const DualT<N>& a, const DualT<N>& b) { return a / b; }
template <int N> NB2_HD DualT<N>& operator/=(DualT<N>& a, const DualT<N>& b) { a = a / b; return a; }
template <int N> NB2_HD DualT<N> nb2_sqrt(const DualT<N>& a) { DualT<N> r; r.v = nb2_sqrt(a.v); const double k = r.v > 0 ? 0.5 * nb2_rcp(r.v) : 0.0; for (int i = 0; i < N; i++) r.d[i] = k * a.d[i]; return r; }
template <int N> NB2_HD DualT<N> nb2_abs(const DualT<N>& a) { return a.v < 0 ? -a : a; }
NB2_HD double gval(double x) { return x; }
// a / b: fast division for plain doubles (nb2_math.cuh), the dual-number operator otherwise
NB2_HD double gdiv(double a, double b) { return nb2_div(a, b); }
template <int N> NB2_HD DualT<N> gdiv(const DualT<N>& a, const DualT<N>& b);
template <int N> NB2_HD double gval(const DualT<N>& x) { return x.v; }

template <class T> NB2_HD Xf<T> gxf_mul(const Xf<T>& A, const Xf<T>& B) { Xf<T> C; C.R_ = mul(A.R_, B.R_); C.p = mul(A.R_, B.p) + A.p; return C; }
template <class T> NB2_HD V3<T> gxf_apply(const Xf<T>& A, const V3<T>& x) { return mul(A.R_, x) + A.p; }
template <class T> NB2_HD V3<T> gxf_apply_inv(const Xf<T>& A, const V3<T>& x) { return mulT(A.R_, x - A.p); }
template <class T> NB2_HD V3<T> gcol3(const M3<T>& R, int j) { return j == 0 ? mk3<T>(R.m00, R.m10, R.m20) : (j == 1 ? mk3<T>(R.m01, R.m11, R.m21) : mk3<T>(R.m02, R.m12, R.m22)); }
template <class T> NB2_HD T gget3(const V3<T>& v, int k) { return k == 0 ? v.x : (k == 1 ? v.y : v.z); }
template <class T> NB2_HD void gset3(V3<T>& v, int k, const T& x) { if (k == 0) v.x = x; else if (k == 1) v.y = x; else v.z = x; }

template <class T> struct ContactOutT { V3<T> point, normal; T depth; int type; };

// contact types emitted (subset of collision::ContactType, dart/collision/Contact.hpp:50-80)
//   1 VERTEX_FACE  2 FACE_VERTEX  3 EDGE_EDGE  4 SPHERE_BOX  5 BOX_SPHERE  6 SPHERE_SPHERE  13 PIPE_SPHERE  14 SPHERE_PIPE  15 PIPE_PIPE

// ---- box vs sphere.  sphere_first=false: object 1 = box, object 2 = sphere (DARTCollide.cpp:1482-1653, normal = contact
// point - centre); sphere_first=true: object 1 = sphere (:1655-1810, normal = centre - contact point, halfspace ignored)
template <class T>
NB2_HDG int collide_box_sphere(const V3<T>& size0, const Xf<T>& T0, const T& r1, const Xf<T>& T1, CR clip, int halfspace, bool sphere_first,
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
  out->normal = (gval(mag) > 1e-6) ? n * gdiv(T(1.0), mag) : nface;
  return 1;
}

// ---- round shapes against each other: sphere-sphere (DARTCollide.cpp:1812-1882), capsule-capsule (:4183-4284), sphere-capsule /
// capsule-sphere (:4286-4420).  A capsule is its axis segment (local z, length h) swept by a sphere of radius r; the contact sits on the line
// between the two closest points, at the radius-weighted position; normal from object 2 towards object 1.
//   types: 6 SPHERE_SPHERE  13 PIPE_SPHERE  14 SPHERE_PIPE  15 PIPE_PIPE  (an axis parameter within 1e-8 of an end counts as that end's sphere)
template <class T> NB2_HDG int round_contact(const V3<T>& c0, const T& r0, const V3<T>& c1, const T& r1, CR clip, bool strict, bool centres_may_coincide,
                                             int type, ContactOutT<T>* out) {
  const V3<T> dv = c0 - c1;
  const T d2 = dot(dv, dv), rsum = r0 + r1;
  if (centres_may_coincide) {                                   // collideSphereSphere: squared test first, zero normal for coincident centres
    if (gval(d2) > gval(rsum) * gval(rsum)) return 0;
    out->point = (c0 * r1 + c1 * r0) * gdiv(T(1.0), rsum);
    out->type = type;
    if (gval(d2) < 1e-6) {                                      // DART_COLLISION_EPS
      if (gval(rsum) > clip) return 0;
      out->normal = mk3<T>(T(0.0), T(0.0), T(0.0)); out->depth = rsum; return 1;
    }
    const T dist = nb2_sqrt(d2);
    const T pen = rsum - dist;
    if (gval(pen) > clip) return 0;
    out->normal = dv * gdiv(T(1.0), dist); out->depth = pen;
    return 1;
  }
  const T dist = nb2_sqrt(d2);
  if (strict ? !(gval(dist) < gval(rsum)) : !(gval(dist) <= gval(rsum))) return 0;
  const T pen = rsum - dist;
  if (gval(pen) > clip) return 0;
  out->point = (c0 * r1 + c1 * r0) * gdiv(T(1.0), rsum);
  out->normal = dv * gdiv(T(1.0), dist);
  out->depth = pen; out->type = type;
  return 1;
}
// parameter in [0, 1] of the point of segment a -> b closest to p (dDistPointToSegment, DARTCollide.cpp:384-410)
template <class T> NB2_HDG T segment_param_of_point(const V3<T>& p, const V3<T>& a, const V3<T>& b) {
  const V3<T> v = b - a, w = p - a;
  const T c1 = dot(w, v);
  if (gval(c1) <= 0) return T(0.0);
  const T c2 = dot(v, v);
  if (gval(c2) <= gval(c1)) return T(1.0);
  return gdiv(c1, c2);
}
// parameters (alpha on p0 -> p1, beta on q0 -> q1) of the closest points of two segments (dSegmentsClosestApproach, DARTCollide.cpp:301-381),
// then clamped to [0, 1] as collideCapsuleCapsule does
template <class T> NB2_HDG void segment_segment_params(const V3<T>& p0, const V3<T>& p1, const V3<T>& q0, const V3<T>& q1, T* alpha, T* beta) {
  const V3<T> u = p1 - p0, v = q1 - q0, w = p0 - q0;
  const T a = dot(u, u), b = dot(u, v), c = dot(v, v), d = dot(u, w), e = dot(v, w);
  const T D = a * c - b * b;
  T sN, sD = D, tN, tD = D;
  if (gval(D) < 1e-15) { sN = T(0.0); sD = T(1.0); tN = e; tD = c; }  // almost parallel: start of the first segment
  else {
    sN = b * e - c * d; tN = a * e - b * d;
    if (gval(sN) < 0.0) { sN = T(0.0); tN = e; tD = c; }
    else if (gval(sN) > gval(sD)) { sN = sD; tN = e + b; tD = c; }
  }
  if (gval(tN) < 0.0) {
    tN = T(0.0);
    if (-gval(d) < 0.0) sN = T(0.0);
    else if (-gval(d) > gval(a)) sN = sD;
    else { sN = -d; sD = a; }
  } else if (gval(tN) > gval(tD)) {
    tN = tD;
    const T db = b - d;
    if (gval(db) < 0.0) sN = T(0.0);
    else if (gval(db) > gval(a)) sN = sD;
    else { sN = db; sD = a; }
  }
  T al = (fabs(gval(sN)) < 1e-15) ? T(0.0) : gdiv(sN, sD);
  T be = (fabs(gval(tN)) < 1e-15) ? T(0.0) : gdiv(tN, tD);
  if (gval(al) < 0) al = T(0.0);
  if (gval(al) > 1) al = T(1.0);
  if (gval(be) < 0) be = T(0.0);
  if (gval(be) > 1) be = T(1.0);
  *alpha = al; *beta = be;
}
NB2_HDG bool at_segment_end(double t) { return fabs(t) < 1e-8 || fabs(1.0 - t) < 1e-8; }
template <class T> NB2_HDG int collide_capsule_capsule(CR h0, const T& r0, const Xf<T>& T0, CR h1, const T& r1, const Xf<T>& T1, CR clip, ContactOutT<T>* out) {
  const V3<T> pa = gxf_apply(T0, mk3<T>(T(0.0), T(0.0), T(-0.5 * h0))), pb = gxf_apply(T0, mk3<T>(T(0.0), T(0.0), T(0.5 * h0)));
  const V3<T> ua = gxf_apply(T1, mk3<T>(T(0.0), T(0.0), T(-0.5 * h1))), ub = gxf_apply(T1, mk3<T>(T(0.0), T(0.0), T(0.5 * h1)));
  T al, be;
  segment_segment_params(pa, pb, ua, ub, &al, &be);
  const V3<T> c0 = pa + (pb - pa) * al, c1 = ua + (ub - ua) * be;
  const bool s0 = at_segment_end(gval(al)), s1 = at_segment_end(gval(be));
  return round_contact(c0, r0, c1, r1, clip, false, false, (s0 && s1) ? 6 : (s0 ? 14 : (s1 ? 13 : 15)), out);
}
// sphere_first: object 1 = sphere (collideSphereCapsule), else object 1 = capsule (collideCapsuleSphere)
template <class T> NB2_HDG int collide_sphere_capsule(const T& rs, const Xf<T>& Ts, CR h, const T& rc, const Xf<T>& Tc, CR clip, bool sphere_first, ContactOutT<T>* out) {
  const V3<T> ua = gxf_apply(Tc, mk3<T>(T(0.0), T(0.0), T(-0.5 * h))), ub = gxf_apply(Tc, mk3<T>(T(0.0), T(0.0), T(0.5 * h)));
  const T al = segment_param_of_point(Ts.p, ua, ub);
  const V3<T> cc = ua + (ub - ua) * al;
  const bool end = at_segment_end(gval(al));
  if (sphere_first) return round_contact(Ts.p, rs, cc, rc, clip, true, false, end ? 6 : 14, out);
  return round_contact(cc, rc, Ts.p, rs, clip, true, false, end ? 6 : 13, out);
}

// ---- box vs box (the algorithm of ODE's dBoxBox as used by the reference, DARTCollide.cpp:764-1450, after Gottschalk's OBB
// separating-axis test).  Three steps, written here as three functions:
//   sat_pick_axis      the 15 candidate axes (3 faces of each box, 9 edge x edge), smallest penetration wins; edge axes must
//                      beat the best face axis by the factor `fudge` (face contacts are preferred)
//   edge_edge_contact  closest points of the two touching edges
//   face_contact       the incident face of the other box, projected into the reference face and clipped against its rectangle
//                      (clip_quad_to_rect); every clip vertex that lies below the reference face becomes a contact
// The ORDER in which axes are tested and clip vertices are emitted fixes the order of the contacts and therefore the LCP row
// order, so it follows the reference; expressions keep the reference's operand order where a comparison depends on them.
template <class T> struct SatAxis {
  T depth_neg;        // s: largest separation (<= 0 when the boxes overlap)
  int code;           // 0 none, 1..3 face of box 1, 4..6 face of box 2, 7..15 edge i of box 1 x edge j of box 2 (7 + 3 i + j)
  bool invert;        // the axis points from box 2 to box 1: flip it
  V3<T> edge_normal;  // unit axis in box-1 coordinates (edge codes only)
};

// Sutherland-Hodgman: clip the quadrilateral quad[0..8) = (x0, y0, ..., x3, y3) against |x| <= h[0], |y| <= h[1], one half-plane at
// a time in the order -x, +x, -y, +y; vertices are emitted in traversal order, a crossing edge emits its intersection after the
// vertex it leaves from; at most 8 vertices (the walk stops when the 8th is written).  Returns the vertex count.
template <class T>
NB2_HD int clip_quad_to_rect(const T h[2], const T quad[8], T out[16]) {
  T bufA[16], bufB[16];
  for (int i = 0; i < 8; i++) bufA[i] = quad[i];
  T* src = bufA; T* dst = bufB;
  int n = 4;
  for (int plane = 0; plane < 4; plane++) {
    const int axis = plane >> 1, other = 1 - axis;
    const double sg = (plane & 1) ? 1.0 : -1.0;
    const T bound = T(sg) * h[axis];
    int k = 0;
    bool full = false;
    for (int i = 0; i < n && !full; i++) {
      const T* a = src + 2 * i;
      const T* nx = src + 2 * ((i + 1 == n) ? 0 : i + 1);
      const bool in_a = sg * gval(a[axis]) < gval(h[axis]), in_n = sg * gval(nx[axis]) < gval(h[axis]);
      if (in_a) { dst[2 * k] = a[0]; dst[2 * k + 1] = a[1]; k++; if (k == 8) { full = true; break; } }
      if (in_a != in_n) {
        dst[2 * k + other] = a[other] + gdiv(nx[other] - a[other], nx[axis] - a[axis]) * (bound - a[axis]);
        dst[2 * k + axis] = bound;
        k++;
        if (k == 8) full = true;
      }
    }
    T* t = src; src = dst; dst = t;
    n = k;
    if (full) break;
  }
  for (int i = 0; i < 2 * n; i++) out[i] = src[i];
  return n;
}

template <class T>
NB2_HD SatAxis<T> sat_pick_axis(const T A[3], const T Bh[3], const T pp[3], const V3<T>& p, const M3<T>& R2, const T Rm[3][3], const T Q[3][3]) {
  const double fudge = 1.05;
  SatAxis<T> ax;
  ax.depth_neg = T(-1e12); ax.code = 0; ax.invert = false; ax.edge_normal = mk3<T>(T(0.0), T(0.0), T(0.0));
  // faces of box 1 (axes of box 1), then faces of box 2
  for (int i = 0; i < 3; i++) {
    const T e1 = pp[i];
    const T s2 = nb2_abs(e1) - (A[i] + Bh[0] * Q[i][0] + Bh[1] * Q[i][1] + Bh[2] * Q[i][2]);
    if (gval(s2) > gval(ax.depth_neg)) { ax.depth_neg = s2; ax.invert = gval(e1) < 0; ax.code = 1 + i; }
  }
  for (int j = 0; j < 3; j++) {
    const T e1 = dot(gcol3(R2, j), p);
    const T s2 = nb2_abs(e1) - (A[0] * Q[0][j] + A[1] * Q[1][j] + A[2] * Q[2][j] + Bh[j]);
    if (gval(s2) > gval(ax.depth_neg)) { ax.depth_neg = s2; ax.invert = gval(e1) < 0; ax.code = 4 + j; }
  }
  // edge i of box 1 x edge j of box 2: axis = e_i x R2[:, j] in box-1 coordinates, i.e. components (0, -R[i2][j], R[i1][j]) rotated
  // into place (i1 = i + 1, i2 = i + 2 mod 3); a1 < a2 / b1 < b2 are the two axes other than i / j in ascending order
  for (int i = 0; i < 3; i++) {
    const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, a1 = i1 < i2 ? i1 : i2, a2 = i1 < i2 ? i2 : i1;
    for (int j = 0; j < 3; j++) {
      const int j1 = (j + 1) % 3, j2 = (j + 2) % 3, b1 = j1 < j2 ? j1 : j2, b2 = j1 < j2 ? j2 : j1;
      const T e1 = pp[i2] * Rm[i1][j] - pp[i1] * Rm[i2][j];
      T s2 = nb2_abs(e1) - (A[a1] * Q[a2][j] + A[a2] * Q[a1][j] + Bh[b1] * Q[i][b2] + Bh[b2] * Q[i][b1]);
      T nc[3];
      nc[i] = T(0.0); nc[i1] = -Rm[i2][j]; nc[i2] = Rm[i1][j];
      const T l = nb2_sqrt(nc[0] * nc[0] + nc[1] * nc[1] + nc[2] * nc[2]);
      if (gval(l) > 0) {
        s2 = gdiv(s2, l);
        if (gval(s2) * fudge > gval(ax.depth_neg)) {
          ax.depth_neg = s2; ax.invert = gval(e1) < 0; ax.code = 7 + 3 * i + j;
          ax.edge_normal = mk3<T>(gdiv(nc[0], l), gdiv(nc[1], l), gdiv(nc[2], l));
        }
      }
    }
  }
  return ax;
}

// dBoxBox; returns the number of contacts written to out (<= 8)
template <class T>
NB2_HDG int collide_box_box(const V3<T>& size0, const Xf<T>& T0, const V3<T>& size1, const Xf<T>& T1, CR clip, ContactOutT<T>* out) {
  const M3<T>&R1 = T0.R_, &R2 = T1.R_;
  const V3<T> p1 = T0.p, p2 = T1.p;
  const T A[3] = {size0.x * T(0.5), size0.y * T(0.5), size0.z * T(0.5)}, Bh[3] = {size1.x * T(0.5), size1.y * T(0.5), size1.z * T(0.5)};
  const V3<T> p = p2 - p1;
  const V3<T> ppv = mulT(R1, p);
  const T pp[3] = {ppv.x, ppv.y, ppv.z};
  T Rm[3][3], Q[3][3];
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { Rm[i][j] = dot(gcol3(R1, i), gcol3(R2, j)); Q[i][j] = nb2_abs(Rm[i][j]); }
  const SatAxis<T> ax = sat_pick_axis(A, Bh, pp, p, R2, Rm, Q);
  const int code = ax.code;
  if (!code) return 0;
  if (gval(ax.depth_neg) > 0.0) return 0;  // a separating axis exists
  V3<T> normal;
  if (code <= 3) normal = gcol3(R1, code - 1);
  else if (code <= 6) normal = gcol3(R2, code - 4);
  else { normal = mul(R1, ax.edge_normal); normal = normal * gdiv(T(1.0), nb2_sqrt(dot(normal, normal))); }
  if (ax.invert) normal = -normal;
  if (code > 6) {
    // ---- edge x edge: walk from each box centre to the touching edge, then closest points of the two lines
    V3<T> pa = p1, pb = p2;
    for (int j = 0; j < 3; j++) { const double sg = (gval(dot(normal, gcol3(R1, j))) > -1e-10) ? 1.0 : -1.0; pa = pa + gcol3(R1, j) * (A[j] * T(sg)); }
    for (int j = 0; j < 3; j++) { const double sg = (gval(dot(normal, gcol3(R2, j))) > -1e-3) ? -1.0 : 1.0; pb = pb + gcol3(R2, j) * (Bh[j] * T(sg)); }
    const V3<T> ua = gcol3(R1, (code - 7) / 3), ub = gcol3(R2, (code - 7) % 3);
    const V3<T> dp = pb - pa;
    const T uaub = dot(ua, ub), q1 = dot(ua, dp), q2 = -dot(ub, dp);
    T d = T(1.0) - uaub * uaub, alpha = T(0.0), beta = T(0.0);
    if (gval(d) > 0.0) { d = gdiv(T(1.0), d); alpha = (q1 + uaub * q2) * d; beta = (uaub * q1 + q2) * d; }
    pa = pa + ua * alpha; pb = pb + ub * beta;
    const T pen = -ax.depth_neg;
    if (gval(pen) > clip) return 0;
    out[0].point = (pa + pb) * T(0.5); out[0].normal = -normal; out[0].depth = pen; out[0].type = 3;
    return 1;
  }
  // ---- face contact.  Reference box a = the box owning the face, incident box b = the other one
  const bool flip = code > 3;
  const M3<T>& Ra = flip ? R2 : R1; const M3<T>& Rb = flip ? R1 : R2;
  const V3<T> pa = flip ? p2 : p1, pb = flip ? p1 : p2;
  const T* Sa = flip ? Bh : A; const T* Sb = flip ? A : Bh;
  const V3<T> normal2 = flip ? -normal : normal;
  const int refAxis = flip ? code - 4 : code - 1;
  // incident face of b: the one most anti-parallel to the reference normal
  const V3<T> nr = mulT(Rb, normal2);
  const double anr[3] = {fabs(gval(nr.x)), fabs(gval(nr.y)), fabs(gval(nr.z))};
  int lanr, a1, a2;
  if (anr[1] > anr[0]) { if (anr[1] > anr[2]) { a1 = 0; lanr = 1; a2 = 2; } else { a1 = 0; a2 = 1; lanr = 2; } }
  else { if (anr[0] > anr[2]) { lanr = 0; a1 = 1; a2 = 2; } else { a1 = 0; a2 = 1; lanr = 2; } }
  const V3<T> center = (gval(gget3(nr, lanr)) < 0) ? (pb - pa + gcol3(Rb, lanr) * Sb[lanr]) : (pb - pa - gcol3(Rb, lanr) * Sb[lanr]);
  const int c1i = (refAxis == 0) ? 1 : 0, c2i = (refAxis == 2) ? 1 : 2;  // the two in-plane axes of the reference face
  // incident face corners in the 2-D coordinates of the reference face
  const T c1 = dot(center, gcol3(Ra, c1i)), c2 = dot(center, gcol3(Ra, c2i));
  T m11 = dot(gcol3(Ra, c1i), gcol3(Rb, a1)), m12 = dot(gcol3(Ra, c1i), gcol3(Rb, a2));
  T m21 = dot(gcol3(Ra, c2i), gcol3(Rb, a1)), m22 = dot(gcol3(Ra, c2i), gcol3(Rb, a2));
  T quad[8];
  {
    const T k1 = m11 * Sb[a1], k2 = m21 * Sb[a1], k3 = m12 * Sb[a2], k4 = m22 * Sb[a2];
    quad[0] = c1 - k1 - k3; quad[1] = c2 - k2 - k4; quad[2] = c1 - k1 + k3; quad[3] = c2 - k2 + k4;
    quad[4] = c1 + k1 + k3; quad[5] = c2 + k2 + k4; quad[6] = c1 + k1 - k3; quad[7] = c2 + k2 - k4;
  }
  const T rect[2] = {Sa[c1i], Sa[c2i]};
  T ret[16];
  const int n = clip_quad_to_rect(rect, quad, ret);
  if (n < 1) return 0;
  // back to 3-D: invert the 2x2 projection, keep the vertices below the reference face
  const T det1 = gdiv(T(1.0), m11 * m22 - m12 * m21);
  m11 = m11 * det1; m12 = m12 * det1; m21 = m21 * det1; m22 = m22 * det1;
  int cnum = 0;
  for (int j = 0; j < n; j++) {
    const T k1 = m22 * (ret[j * 2] - c1) - m12 * (ret[j * 2 + 1] - c2);
    const T k2 = -m21 * (ret[j * 2] - c1) + m11 * (ret[j * 2 + 1] - c2);
    const V3<T> pt = center + gcol3(Rb, a1) * k1 + gcol3(Rb, a2) * k2;
    const T dep = Sa[refAxis] - dot(normal2, pt);
    if (gval(dep) >= 0) {
      ContactOutT<T>& c = out[cnum];
      c.point = pt + pa; c.normal = -normal; c.depth = dep;
      // a clip vertex on a corner of the rectangle is a vertex of the reference box, one strictly inside is a vertex of the incident
      // box, one on a single side is an edge-edge crossing (DARTCollide.cpp:1283-1379)
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
  *t1 = t * gdiv(T(1.0), nb2_sqrt(dot(t, t)));
  *t2 = cross(n, *t1);
}

}  // namespace nb2
