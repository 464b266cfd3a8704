// TEST INFRASTRUCTURE ONLY — contact generation and contact-constraint rows for the fp64 oracle.
// Restates (scalar-generic, so the same code runs on dual numbers for the Jacobians):
//   collide() pair loop / filter      dart/collision/dart/DARTCollisionDetector.cpp:150-175, dart/collision/CollisionFilter.cpp:105-152
//   collideBoxSphere / collideSphereBox   dart/collision/dart/DARTCollide.cpp:1482-1653, 1655-1810
//   collideBoxBox -> dBoxBox            DARTCollide.cpp:764-1450 (+ intersectRectQuad :513-580, dLineClosestApproach)
//   collideSphereSphere :1812-1882, collideCapsuleCapsule :4183-4284, collideSphereCapsule / collideCapsuleSphere :4286-4420
//   collideBoxCapsule / collideCapsuleBox DARTCollide.cpp:4422-4645 — the reference asks libccd's MPR (third-party, absent here)
//        which part of the capsule touches; on a box FACE that is the deeper end sphere, tested with the functions above
//        (:4462-4491).  This restatement picks the deeper end sphere geometrically and flags the "pipe" (side-on) case as
//        unsupported instead of calling MPR.  Documented deviation.
//   contact filtering                   dart/constraint/ConstraintSolver.cpp:576-601
//   ContactConstraint (rows, b, bounds) dart/constraint/ContactConstraint.cpp:66-230, 361-514, 687-695, 734-795
#pragma once
#include <vector>

#include "spatial.hpp"

namespace orc {

enum { SH_BOX = 0, SH_SPHERE = 1, SH_CAPSULE = 2 };
enum { CLIP_BOTH = 0, CLIP_TOP = 1, CLIP_BOTTOM = 2 };
// subset of collision::ContactType (dart/collision/Contact.hpp:50-80) that these generators emit
enum { CT_UNSUPPORTED = 0, CT_VERTEX_FACE = 1, CT_FACE_VERTEX = 2, CT_EDGE_EDGE = 3, CT_SPHERE_BOX = 4, CT_BOX_SPHERE = 5, CT_SPHERE_SPHERE = 6,
       CT_PIPE_SPHERE = 13, CT_SPHERE_PIPE = 14, CT_PIPE_PIPE = 15 };

template <class S> struct Contact {
  Vec3<S> point, normal;
  S depth;
  int bodyA, bodyB;   // raw body indices (object 1 / object 2)
  int shapeA, shapeB;
  int type;
};

template <class S> inline S sabs(const S& x) { return val(x) < 0 ? -x : x; }

// ---- box (o1) vs sphere (o2): normal points from the sphere towards the box (object 2 -> object 1)
template <class S>
inline void collide_box_sphere(const Vec3<S>& size0, const Iso<S>& T0, const S& r1, const Iso<S>& T1, double clip, int halfspace,
                               int bA, int bB, int sA, int sB, std::vector<Contact<S>>& out) {
  Vec3<S> half = size0 * S(0.5);
  bool inside = true;
  Vec3<S> c0 = T1.p;
  Vec3<S> p = apply(inverse(T0), c0);
  for (int k = 0; k < 3; k++) {
    if (val(p[k]) < -val(half[k])) { p[k] = -half[k]; inside = false; }
    if (val(p[k]) > val(half[k])) { p[k] = half[k]; inside = false; }
  }
  Contact<S> c; c.bodyA = bA; c.bodyB = bB; c.shapeA = sA; c.shapeB = sB; c.type = CT_BOX_SPHERE;
  auto nearest_face_normal = [&](S& mn) {
    mn = half[0] - sabs(p[0]); int idx = 0;
    S t = half[1] - sabs(p[1]); if (val(t) < val(mn)) { mn = t; idx = 1; }
    t = half[2] - sabs(p[2]); if (val(t) < val(mn)) { mn = t; idx = 2; }
    Vec3<S> n = v3<S>(S(0.0), S(0.0), S(0.0));
    n[idx] = S(val(p[idx]) > 0.0 ? -1.0 : 1.0);
    return mul(T0.R, n);
  };
  if (inside) {
    S mn; Vec3<S> n = nearest_face_normal(mn);
    S pen = mn + r1;
    if (val(pen) > clip) return;
    c.type = CT_FACE_VERTEX; c.point = c0; c.normal = n; c.depth = pen; out.push_back(c); return;
  }
  Vec3<S> cp = apply(T0, p);
  Vec3<S> n = cp - c0;
  S mag = sqrt(dot(n, n));
  S pen = r1 - mag;
  if (val(pen) > clip) return;
  if (halfspace == CLIP_BOTTOM && val(apply(inverse(T1), cp)[2]) >= 0) return;
  if (halfspace == CLIP_TOP && val(apply(inverse(T1), cp)[2]) <= 0) return;
  if (val(pen) < 0.0) return;
  if (val(mag) > 1e-6) { c.point = cp; c.normal = n * (S(1.0) / mag); c.depth = pen; out.push_back(c); }
  else { S mn; c.normal = nearest_face_normal(mn); c.point = cp; c.depth = pen; out.push_back(c); }
}

// ---- sphere (o1) vs box (o2): normal = sphere centre - contact point (object 2 -> object 1); halfspace ignored (:1664)
template <class S>
inline void collide_sphere_box(const S& r0, const Iso<S>& T0, const Vec3<S>& size1, const Iso<S>& T1, double clip,
                               int bA, int bB, int sA, int sB, std::vector<Contact<S>>& out) {
  Vec3<S> half = size1 * S(0.5);
  bool inside = true;
  Vec3<S> c0 = T0.p;
  Vec3<S> p = apply(inverse(T1), c0);
  for (int k = 0; k < 3; k++) {
    if (val(p[k]) < -val(half[k])) { p[k] = -half[k]; inside = false; }
    if (val(p[k]) > val(half[k])) { p[k] = half[k]; inside = false; }
  }
  Contact<S> c; c.bodyA = bA; c.bodyB = bB; c.shapeA = sA; c.shapeB = sB; c.type = CT_SPHERE_BOX;
  auto nearest_face_normal = [&](S& mn) {
    mn = half[0] - sabs(p[0]); int idx = 0;
    S t = half[1] - sabs(p[1]); if (val(t) < val(mn)) { mn = t; idx = 1; }
    t = half[2] - sabs(p[2]); if (val(t) < val(mn)) { mn = t; idx = 2; }
    Vec3<S> n = v3<S>(S(0.0), S(0.0), S(0.0));
    n[idx] = S(val(p[idx]) > 0.0 ? 1.0 : -1.0);
    return mul(T1.R, n);
  };
  if (inside) {
    S mn; Vec3<S> n = nearest_face_normal(mn);
    S pen = mn + r0;
    if (val(pen) > clip) return;
    c.type = CT_VERTEX_FACE; c.point = c0; c.normal = n; c.depth = pen; out.push_back(c); return;
  }
  Vec3<S> cp = apply(T1, p);
  Vec3<S> n = c0 - cp;
  S mag = sqrt(dot(n, n));
  S pen = r0 - mag;
  if (val(pen) > clip) return;
  if (val(pen) < 0.0) return;
  if (val(mag) > 1e-6) { c.point = cp; c.normal = n * (S(1.0) / mag); c.depth = pen; out.push_back(c); }
  else { S mn; c.normal = nearest_face_normal(mn); c.point = cp; c.depth = pen; out.push_back(c); }
}

// ---- sphere vs sphere (collideSphereSphere, DARTCollide.cpp:1812-1882)
template <class S>
inline void collide_sphere_sphere(const S& r0in, const Iso<S>& T0, const S& r1in, const Iso<S>& T1, double clip, int bA, int bB, int sA, int sB,
                                  std::vector<Contact<S>>& out) {
  S r0 = r0in, r1 = r1in;
  const S rsum = r0 + r1;
  Vec3<S> normal = T0.p - T1.p;
  S nsq = dot(normal, normal);
  if (val(nsq) > val(rsum) * val(rsum)) return;
  r0 = r0 / rsum; r1 = r1 / rsum;
  Contact<S> c; c.bodyA = bA; c.bodyB = bB; c.shapeA = sA; c.shapeB = sB; c.type = CT_SPHERE_SPHERE;
  c.point = T0.p * r1 + T1.p * r0;
  if (val(nsq) < 1e-6) {  // DART_COLLISION_EPS: coincident centres, zero normal (the constraint filter drops it)
    if (val(rsum) > clip) return;
    c.normal = v3<S>(S(0.0), S(0.0), S(0.0)); c.depth = rsum; out.push_back(c); return;
  }
  const S len = sqrt(nsq);
  const S pen = rsum - len;
  if (val(pen) > clip) return;
  c.normal = normal * (S(1.0) / len); c.depth = pen; out.push_back(c);
}
// dDistPointToSegment (DARTCollide.cpp:384-410): distance and segment parameter
template <class S> inline S dist_point_segment(const Vec3<S>& p, const Vec3<S>& ua, const Vec3<S>& ub, S& alpha) {
  const Vec3<S> v = ub - ua, w = p - ua;
  const S c1 = dot(w, v);
  if (val(c1) <= 0) { alpha = S(0.0); const Vec3<S> d = p - ua; return sqrt(dot(d, d)); }
  const S c2 = dot(v, v);
  if (val(c2) <= val(c1)) { alpha = S(1.0); const Vec3<S> d = p - ub; return sqrt(dot(d, d)); }
  alpha = c1 / c2;
  const Vec3<S> d = p - (ua + v * alpha);
  return sqrt(dot(d, d));
}
// dSegmentsClosestApproach (DARTCollide.cpp:301-381): segment 1 = pa -> pb (alpha), segment 2 = ua -> ub (beta)
template <class S> inline void segments_closest_approach(const Vec3<S>& pa, const Vec3<S>& ua, const Vec3<S>& pb, const Vec3<S>& ub, S& alpha, S& beta) {
  const Vec3<S> u = pb - pa, v = ub - ua, w = pa - ua;
  const S a = dot(u, u), b = dot(u, v), c = dot(v, v), d = dot(u, w), e = dot(v, w);
  const S D = a * c - b * b;
  S sN, sD = D, tN, tD = D;
  const double SMALL = 1e-15;
  if (val(D) < SMALL) { sN = S(0.0); sD = S(1.0); tN = e; tD = c; }
  else {
    sN = b * e - c * d; tN = a * e - b * d;
    if (val(sN) < 0.0) { sN = S(0.0); tN = e; tD = c; }
    else if (val(sN) > val(sD)) { sN = sD; tN = e + b; tD = c; }
  }
  if (val(tN) < 0.0) {
    tN = S(0.0);
    if (-val(d) < 0.0) sN = S(0.0);
    else if (-val(d) > val(a)) sN = sD;
    else { sN = -d; sD = a; }
  } else if (val(tN) > val(tD)) {
    tN = tD;
    if ((-val(d) + val(b)) < 0.0) sN = S(0.0);
    else if ((-val(d) + val(b)) > val(a)) sN = sD;
    else { sN = b - d; sD = a; }
  }
  alpha = (std::fabs(val(sN)) < SMALL) ? S(0.0) : sN / sD;
  beta = (std::fabs(val(tN)) < SMALL) ? S(0.0) : tN / tD;
}
template <class S> inline Vec3<S> capsule_end(const Iso<S>& T, double h, double sign) { return apply(T, v3<S>(S(0.0), S(0.0), S(sign * h / 2))); }
inline bool near_end(double t) { return std::fabs(t) < 1e-8 || std::fabs(1.0 - t) < 1e-8; }
// collideCapsuleCapsule (DARTCollide.cpp:4183-4284)
template <class S>
inline void collide_capsule_capsule(double h0, const S& r0in, const Iso<S>& T0, double h1, const S& r1in, const Iso<S>& T1, double clip,
                                    int bA, int bB, int sA, int sB, std::vector<Contact<S>>& out) {
  const Vec3<S> pa = capsule_end(T0, h0, -1.0), pb = capsule_end(T0, h0, 1.0), ua = capsule_end(T1, h1, -1.0), ub = capsule_end(T1, h1, 1.0);
  S alpha, beta;
  segments_closest_approach(pa, ua, pb, ub, alpha, beta);
  if (val(alpha) < 0) alpha = S(0.0);
  if (val(alpha) > 1) alpha = S(1.0);
  if (val(beta) < 0) beta = S(0.0);
  if (val(beta) > 1) beta = S(1.0);
  const Vec3<S> c0 = pa + (pb - pa) * alpha, c1 = ua + (ub - ua) * beta;
  const Vec3<S> dv = c0 - c1;
  const S dist = sqrt(dot(dv, dv)), rsum = r0in + r1in;
  if (!(val(dist) <= val(rsum))) return;
  Contact<S> c; c.bodyA = bA; c.bodyB = bB; c.shapeA = sA; c.shapeB = sB;
  c.depth = rsum - dist;
  if (val(c.depth) > clip) return;
  c.point = c0 * (r1in / rsum) + c1 * (r0in / rsum);
  c.normal = dv * (S(1.0) / dist);
  const bool s0 = near_end(val(alpha)), s1 = near_end(val(beta));
  c.type = (s0 && s1) ? CT_SPHERE_SPHERE : (s0 ? CT_SPHERE_PIPE : (s1 ? CT_PIPE_SPHERE : CT_PIPE_PIPE));
  out.push_back(c);
}
// collideSphereCapsule / collideCapsuleSphere (DARTCollide.cpp:4286-4420)
template <class S>
inline void collide_sphere_capsule(const S& rs, const Iso<S>& Ts, double h, const S& rc, const Iso<S>& Tc, double clip, bool sphere_first,
                                   int bA, int bB, int sA, int sB, std::vector<Contact<S>>& out) {
  const Vec3<S> ua = capsule_end(Tc, h, -1.0), ub = capsule_end(Tc, h, 1.0);
  S alpha;
  const S dist = dist_point_segment(Ts.p, ua, ub, alpha);
  const S rsum = rs + rc;
  if (!(val(dist) < val(rsum))) return;
  const Vec3<S> cc = ua + (ub - ua) * alpha;
  Contact<S> c; c.bodyA = bA; c.bodyB = bB; c.shapeA = sA; c.shapeB = sB;
  c.depth = rsum - dist;
  if (val(c.depth) > clip) return;
  c.point = Ts.p * (rc / rsum) + cc * (rs / rsum);
  const Vec3<S> dv = sphere_first ? Ts.p - cc : cc - Ts.p;
  c.normal = dv * (S(1.0) / sqrt(dot(dv, dv)));
  c.type = near_end(val(alpha)) ? CT_SPHERE_SPHERE : (sphere_first ? CT_SPHERE_PIPE : CT_PIPE_SPHERE);
  out.push_back(c);
}

// ---- intersectRectQuad (DARTCollide.cpp:513-580): clip quad p[8] against rect +-h; returns #points in ret[16]
template <class S> inline int intersect_rect_quad(const S h[2], S p[8], S ret[16]) {
  int nq = 4, nr = 0;
  S buffer[16];
  S* q = p; S* r = ret;
  for (int dir = 0; dir <= 1; dir++) {
    for (int sign = -1; sign <= 1; sign += 2) {
      S* pq = q; S* pr = r; nr = 0;
      for (int i = nq; i > 0; i--) {
        if (sign * val(pq[dir]) < val(h[dir])) {
          pr[0] = pq[0]; pr[1] = pq[1]; pr += 2; nr++;
          if (nr & 8) { q = r; goto done; }
        }
        S* nextq = (i > 1) ? pq + 2 : q;
        if ((sign * val(pq[dir]) < val(h[dir])) ^ (sign * val(nextq[dir]) < val(h[dir]))) {
          pr[1 - dir] = pq[1 - dir] + (nextq[1 - dir] - pq[1 - dir]) / (nextq[dir] - pq[dir]) * (S((double)sign) * h[dir] - pq[dir]);
          pr[dir] = S((double)sign) * h[dir];
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

// dLineClosestApproach (DARTCollide.cpp:270-298): closest points pa + alpha ua , pb + beta ub
template <class S>
inline void line_closest_approach(const Vec3<S>& pa, const Vec3<S>& ua, const Vec3<S>& pb, const Vec3<S>& ub, S* alpha, S* beta) {
  Vec3<S> p = pb - pa;
  S uaub = dot(ua, ub), q1 = dot(ua, p), q2 = -dot(ub, p);
  S d = S(1.0) - uaub * uaub;
  if (val(d) <= 0.0) { *alpha = S(0.0); *beta = S(0.0); }
  else { d = S(1.0) / d; *alpha = (q1 + uaub * q2) * d; *beta = (uaub * q1 + q2) * d; }
}

// ---- dBoxBox (DARTCollide.cpp:764-1450).  R columns are the box axes; A,B half sizes.
template <class S>
inline void collide_box_box(const Vec3<S>& size0, const Iso<S>& T0, const Vec3<S>& size1, const Iso<S>& T1, double clip,
                            int bA, int bB, int sA, int sB, std::vector<Contact<S>>& out) {
  const double fudge = 1.05;
  const Mat3<S>&R1 = T0.R, &R2 = T1.R;
  Vec3<S> p1 = T0.p, p2 = T1.p;
  S A[3] = {size0[0] * S(0.5), size0[1] * S(0.5), size0[2] * S(0.5)}, B[3] = {size1[0] * S(0.5), size1[1] * S(0.5), size1[2] * S(0.5)};
  auto col = [](const Mat3<S>& R, int j) { return v3<S>(R(0, j), R(1, j), R(2, j)); };
  Vec3<S> p = p2 - p1;
  Vec3<S> pp = mulT(R1, p);
  S Rm[3][3], Q[3][3];
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { Rm[i][j] = dot(col(R1, i), col(R2, j)); Q[i][j] = sabs(Rm[i][j]); }
  S s = S(-1e12), s2;
  int invert_normal = 0, code = 0;
  int normalR_box = 0, normalR_col = -1;  // normalR: column of R1 (box 1) or R2 (box 2)
  Vec3<S> normalC = v3<S>(S(0.0), S(0.0), S(0.0));
#define ORC_TST(expr1, expr2, box, colj, cc) { S e1 = (expr1); s2 = sabs(e1) - (expr2); if (val(s2) > val(s)) { s = s2; normalR_box = box; normalR_col = colj; invert_normal = (val(e1) < 0); code = (cc); } }
  ORC_TST(pp[0], (A[0] + B[0] * Q[0][0] + B[1] * Q[0][1] + B[2] * Q[0][2]), 1, 0, 1)
  ORC_TST(pp[1], (A[1] + B[0] * Q[1][0] + B[1] * Q[1][1] + B[2] * Q[1][2]), 1, 1, 2)
  ORC_TST(pp[2], (A[2] + B[0] * Q[2][0] + B[1] * Q[2][1] + B[2] * Q[2][2]), 1, 2, 3)
  ORC_TST(dot(col(R2, 0), p), (A[0] * Q[0][0] + A[1] * Q[1][0] + A[2] * Q[2][0] + B[0]), 2, 0, 4)
  ORC_TST(dot(col(R2, 1), p), (A[0] * Q[0][1] + A[1] * Q[1][1] + A[2] * Q[2][1] + B[1]), 2, 1, 5)
  ORC_TST(dot(col(R2, 2), p), (A[0] * Q[0][2] + A[1] * Q[1][2] + A[2] * Q[2][2] + B[2]), 2, 2, 6)
#undef ORC_TST
#define ORC_TST2(expr1, expr2, n1, n2, n3, cc) { S e1 = (expr1); s2 = sabs(e1) - (expr2); S N1 = (n1), N2 = (n2), N3 = (n3); S l = sqrt(N1 * N1 + N2 * N2 + N3 * N3); \
    if (val(l) > 0) { s2 = s2 / l; if (val(s2) * fudge > val(s)) { s = s2; normalR_col = -1; normalC = v3<S>(N1 / l, N2 / l, N3 / l); invert_normal = (val(e1) < 0); code = (cc); } } }
  const S Z = S(0.0);
  ORC_TST2(pp[2] * Rm[1][0] - pp[1] * Rm[2][0], (A[1] * Q[2][0] + A[2] * Q[1][0] + B[1] * Q[0][2] + B[2] * Q[0][1]), Z, -Rm[2][0], Rm[1][0], 7)
  ORC_TST2(pp[2] * Rm[1][1] - pp[1] * Rm[2][1], (A[1] * Q[2][1] + A[2] * Q[1][1] + B[0] * Q[0][2] + B[2] * Q[0][0]), Z, -Rm[2][1], Rm[1][1], 8)
  ORC_TST2(pp[2] * Rm[1][2] - pp[1] * Rm[2][2], (A[1] * Q[2][2] + A[2] * Q[1][2] + B[0] * Q[0][1] + B[1] * Q[0][0]), Z, -Rm[2][2], Rm[1][2], 9)
  ORC_TST2(pp[0] * Rm[2][0] - pp[2] * Rm[0][0], (A[0] * Q[2][0] + A[2] * Q[0][0] + B[1] * Q[1][2] + B[2] * Q[1][1]), Rm[2][0], Z, -Rm[0][0], 10)
  ORC_TST2(pp[0] * Rm[2][1] - pp[2] * Rm[0][1], (A[0] * Q[2][1] + A[2] * Q[0][1] + B[0] * Q[1][2] + B[2] * Q[1][0]), Rm[2][1], Z, -Rm[0][1], 11)
  ORC_TST2(pp[0] * Rm[2][2] - pp[2] * Rm[0][2], (A[0] * Q[2][2] + A[2] * Q[0][2] + B[0] * Q[1][1] + B[1] * Q[1][0]), Rm[2][2], Z, -Rm[0][2], 12)
  ORC_TST2(pp[1] * Rm[0][0] - pp[0] * Rm[1][0], (A[0] * Q[1][0] + A[1] * Q[0][0] + B[1] * Q[2][2] + B[2] * Q[2][1]), -Rm[1][0], Rm[0][0], Z, 13)
  ORC_TST2(pp[1] * Rm[0][1] - pp[0] * Rm[1][1], (A[0] * Q[1][1] + A[1] * Q[0][1] + B[0] * Q[2][2] + B[2] * Q[2][0]), -Rm[1][1], Rm[0][1], Z, 14)
  ORC_TST2(pp[1] * Rm[0][2] - pp[0] * Rm[1][2], (A[0] * Q[1][2] + A[1] * Q[0][2] + B[0] * Q[2][1] + B[1] * Q[2][0]), -Rm[1][2], Rm[0][2], Z, 15)
#undef ORC_TST2
  if (!code) return;
  if (val(s) > 0.0) return;
  Vec3<S> normal;
  if (normalR_col >= 0) normal = col(normalR_box == 1 ? R1 : R2, normalR_col);
  else { normal = mul(R1, normalC); normal = normal * (S(1.0) / sqrt(dot(normal, normal))); }
  if (invert_normal) normal = neg(normal);
  Contact<S> c; c.bodyA = bA; c.bodyB = bB; c.shapeA = sA; c.shapeB = sB;
  if (code > 6) {
    Vec3<S> pa = p1, pb = p2;
    for (int j = 0; j < 3; j++) { double sg = (val(dot(normal, col(R1, j))) > -1e-10) ? 1.0 : -1.0; pa = pa + col(R1, j) * (A[j] * S(sg)); }
    for (int j = 0; j < 3; j++) { double sg = (val(dot(normal, col(R2, j))) > -1e-3) ? -1.0 : 1.0; pb = pb + col(R2, j) * (B[j] * S(sg)); }
    Vec3<S> ua = col(R1, (code - 7) / 3), ub = col(R2, (code - 7) % 3);
    S alpha, beta;
    line_closest_approach(pa, ua, pb, ub, &alpha, &beta);
    pa = pa + ua * alpha; pb = pb + ub * beta;
    S pen = -s;
    if (val(pen) > clip) return;
    c.point = (pa + pb) * S(0.5); c.normal = neg(normal); c.depth = pen; c.type = CT_EDGE_EDGE;
    out.push_back(c);
    return;
  }
  const Mat3<S>*Ra, *Rb; Vec3<S> pa, pb; const S *Sa, *Sb; bool flip;
  if (code <= 3) { Ra = &R1; Rb = &R2; pa = p1; pb = p2; Sa = A; Sb = B; flip = false; }
  else { Ra = &R2; Rb = &R1; pa = p2; pb = p1; Sa = B; Sb = A; flip = true; }
  Vec3<S> normal2 = (code <= 3) ? normal : neg(normal);
  Vec3<S> nr = mulT(*Rb, normal2);
  double anr[3] = {std::fabs(val(nr[0])), std::fabs(val(nr[1])), std::fabs(val(nr[2]))};
  int lanr, a1, a2;
  if (anr[1] > anr[0]) { if (anr[1] > anr[2]) { a1 = 0; lanr = 1; a2 = 2; } else { a1 = 0; a2 = 1; lanr = 2; } }
  else { if (anr[0] > anr[2]) { lanr = 0; a1 = 1; a2 = 2; } else { a1 = 0; a2 = 1; lanr = 2; } }
  Vec3<S> center = (val(nr[lanr]) < 0) ? (pb - pa + col(*Rb, lanr) * Sb[lanr]) : (pb - pa - col(*Rb, lanr) * Sb[lanr]);
  int codeN = (code <= 3) ? code - 1 : code - 4, code1, code2;
  if (codeN == 0) { code1 = 1; code2 = 2; } else if (codeN == 1) { code1 = 0; code2 = 2; } else { code1 = 0; code2 = 1; }
  S quad[8];
  S c1 = dot(center, col(*Ra, code1)), c2 = dot(center, col(*Ra, code2));
  S m11 = dot(col(*Ra, code1), col(*Rb, a1)), m12 = dot(col(*Ra, code1), col(*Rb, a2));
  S m21 = dot(col(*Ra, code2), col(*Rb, a1)), m22 = dot(col(*Ra, code2), col(*Rb, a2));
  {
    S k1 = m11 * Sb[a1], k2 = m21 * Sb[a1], k3 = m12 * Sb[a2], k4 = m22 * Sb[a2];
    quad[0] = c1 - k1 - k3; quad[1] = c2 - k2 - k4; quad[2] = c1 - k1 + k3; quad[3] = c2 - k2 + k4;
    quad[4] = c1 + k1 + k3; quad[5] = c2 + k2 + k4; quad[6] = c1 + k1 - k3; quad[7] = c2 + k2 - k4;
  }
  S rect[2] = {Sa[code1], Sa[code2]};
  S ret[16];
  int n = intersect_rect_quad(rect, quad, ret);
  if (n < 1) return;
  Vec3<S> point[8]; S dep[8];
  S det1 = S(1.0) / (m11 * m22 - m12 * m21);
  m11 = m11 * det1; m12 = m12 * det1; m21 = m21 * det1; m22 = m22 * det1;
  int cnum = 0;
  for (int j = 0; j < n; j++) {
    S k1 = m22 * (ret[j * 2] - c1) - m12 * (ret[j * 2 + 1] - c2);
    S k2 = -m21 * (ret[j * 2] - c1) + m11 * (ret[j * 2 + 1] - c2);
    point[cnum] = center + col(*Rb, a1) * k1 + col(*Rb, a2) * k2;
    dep[cnum] = Sa[codeN] - dot(normal2, point[cnum]);
    if (val(dep[cnum]) >= 0) { ret[cnum * 2] = ret[j * 2]; ret[cnum * 2 + 1] = ret[j * 2 + 1]; cnum++; }
  }
  if (cnum < 1) return;
  for (int j = 0; j < cnum; j++) {
    c.point = point[j] + pa; c.normal = neg(normal); c.depth = dep[j];
    const bool onX = std::fabs(val(ret[j * 2])) == val(rect[0]), onY = std::fabs(val(ret[j * 2 + 1])) == val(rect[1]);
    if (onX && onY) {
      if (flip) { c.type = CT_FACE_VERTEX; c.point = c.point + c.normal * c.depth; }
      else { c.type = CT_VERTEX_FACE; c.point = c.point - c.normal * c.depth; }
    } else if (!onX && !onY) c.type = flip ? CT_VERTEX_FACE : CT_FACE_VERTEX;
    else c.type = CT_EDGE_EDGE;
    out.push_back(c);  // every clip point is emitted (the 4-point cull is commented out, :1384-1448)
  }
}

}  // namespace orc
