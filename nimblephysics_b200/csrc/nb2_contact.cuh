// Contact / boxed-LCP stage of one world, run by ONE thread in fp64 after the ABA kernel has produced the
// unconstrained step (v* = v + dt qdd) and streamed U, psi, joint transforms to the saved stream.
//
// Restates, per world (reference file:line):
//   contact generation vs static colliders   dart/collision/dart/DARTCollide.cpp:764-1450 (dBoxBox), :1482-1810 (box/sphere),
//                                            :4422-4645 (capsule end spheres on a box face; see oracle/contact.hpp for the MPR note)
//   contact filtering                        dart/constraint/ConstraintSolver.cpp:576-601
//   rows, b, bounds                          dart/constraint/ContactConstraint.cpp:66-230, 361-514, 687-695, 734-795
//   A by impulse tests                       dart/constraint/BoxedLcpConstraintSolver.cpp:190-349 (impulse-ABA: BodyNode.cpp:2117-2138,
//                                            2188-2215, GenericJoint.hpp:2482-2498, 2607-2613, 2713-2725) — here reusing the forward's U, psi
//   solve chain                              BoxedLcpConstraintSolver.cpp:352-789; LCPUtils.cpp:12-140; PgsBoxedLcpSolver.cpp:79-278
//   classification / standardisation         dart/neural/ConstrainedGroupGradientMatrices.cpp:482-872, 218-339
//   impulses, velocity update                ContactConstraint.cpp:630-684, Skeleton.cpp:13571-13595
// New: the minimum-norm least-squares solves (Eigen completeOrthogonalDecomposition in the reference) are done with a
// rank-revealing pivoted Cholesky (symmetric case) / normal equations (sliding-friction case) instead of a QR/SVD.
#pragma once
#include <type_traits>
#include "nb2_dantzig.cuh"
#include "nb2_dyn.cuh"
#include "nb2_geom.cuh"

#include "../../include/nb2.h"  // NB2_MAX_CONTACTS, NB2_MAX_ROWS

#define NB2_MAX_SHAPES 24
#define NB2_MAX_PAIRS 64

// status bits (per world)
#define NB2_ST_SHORTCIRCUIT 1
#define NB2_ST_DANTZIG 2
#define NB2_ST_DANTZIG_FAILED 4
#define NB2_ST_PGS 8
#define NB2_ST_FRICTION_DROPPED 16
#define NB2_ST_NAN 32
#define NB2_ST_NOT_STANDARDIZED 64
#define NB2_ST_UNSUPPORTED_GEOMETRY 128
#define NB2_ST_CONTACT_OVERFLOW 256
#define NB2_ST_MERGED 512
#define NB2_ST_BOUNCE 1024  // a restitution (bounce) or penetration-correction term raised some b_i: the backward of such a step is not implemented (NaN, loud)  // LCPUtils::reduce merged near-identical columns before a solver ran

// ConstraintMapping (dart/neural/ConstrainedGroupGradientMatrices.hpp:33-39)
#define NB2_MAP_NOT_CLAMPING (-1)
#define NB2_MAP_CLAMPING (-2)
#define NB2_MAP_ILLEGAL (-3)

struct Nb2ContactDev {
  int nshapes, npairs, pen_correction, pad;
  double clip_depth, fallback_cfm;
  int16_t shape_body[NB2_MAX_SHAPES];       // canonical body index, -1 = static (world-fixed)
  int16_t shape_type[NB2_MAX_SHAPES];       // 0 box, 1 sphere, 2 capsule
  int16_t shape_orig_body[NB2_MAX_SHAPES];  // reference BodyNode index (reported with the contacts)
  int16_t pair_a[NB2_MAX_PAIRS], pair_b[NB2_MAX_PAIRS];  // collision pairs in the reference's enumeration order
  double shape_dims[NB2_MAX_SHAPES][3];
  double shape_T[NB2_MAX_SHAPES][12];       // shape frame -> canonical body frame (or world)
  double shape_mu[NB2_MAX_SHAPES], shape_rest[NB2_MAX_SHAPES];
};

namespace nb2 {

// The threads that cooperate on one world (the GROUP: `nl` adjacent lanes of a warp, this one is number `cl`).  The serial
// parts of the contact stage are executed REDUNDANTLY by every thread of the group — in SIMT a warp instruction costs the
// same whether one or eight lanes are active, and identical inputs keep the group in lock step — so that the dense inner
// products can be split across the lanes and summed with a butterfly of shuffles (every lane receives the same bits,
// which keeps all data-dependent branches uniform).  Host builds and one-thread-per-world launches use the default (1 lane).
struct Grp {
  int cl = 0, nl = 1;
  unsigned mask = 0;
  NB2_HD CR sum(CR v) const {
#ifdef __CUDA_ARCH__
    for (int off = nl >> 1; off > 0; off >>= 1) v += __shfl_xor_sync(mask, v, off);
#endif
    return v;
  }
};

#define NB2_CONTACT_LANES 8  // threads that may cooperate on one world in the contact kernel
template <int ST>
struct ContactWsT {  // per-world fp64 workspace; lane-interleaved on the device (ST = 32), contiguous on the host (ST = 1)
  typedef SP<CR, ST> PD; typedef SP<int, ST> PI; typedef SP<unsigned char, ST> PB;
  PD W, T, V, pI, uI, dqd, vstar;
  PD cpoint, cnormal, cdepth, cmu, crest;
  PI cbodyA, cbodyB, ctype, cshapeA, cshapeB;
  PD JA, JB, b, lo, hi, x, x0, rest, colnorm;
  PI findex, mapping, clampIdx, ubIdx;
  PD A, Aw, L, Q, Q2;
  PD v1, v2, v3, v4, v5, v6, v7, v8;
  PI i1, i2;
  PB st8;
  Grp grp;  // the threads cooperating on this world (set by the kernel; default: one)
  PD lbuf;  // NB2_CONTACT_LANES private sweep buffers (pI, V: 6 nb each; uI, dqd: ndof each) for the parallel impulse tests
  PI meta;  // m, nc, status carried between the phases of world_contact
};
NB2_HD size_t contact_rec_doubles(int ndof) { return 2 + 2 * (size_t)NB2_MAX_ROWS + ndof + (size_t)NB2_MAX_ROWS * NB2_MAX_ROWS; }
// doubles per WORLD (the kernel allocates 32x this per warp)
NB2_HD size_t contact_ws_doubles(int nb, int ndof) {
  const int MC = NB2_MAX_CONTACTS, MR = NB2_MAX_ROWS;
  return (size_t)nb * 36 + 3 * ndof + MC * 9 + 5 * ((MC + 1) / 2) + 2 * MR * 6 + 7 * MR + 4 * ((MR + 1) / 2) + 5 * (size_t)MR * MR + 8 * MR
         + 2 * ((MR + 1) / 2) + (2 * MR + 7) / 8 + 4 + (size_t)NB2_CONTACT_LANES * (12 * nb + 2 * ndof) + 2;
}
// `block`: start of the warp's (device) or world's (host) block; `lane`: 0 on the host
template <int ST>
NB2_HD ContactWsT<ST> carve_ws(CR* block, int lane, int nb, int ndof) {
  const int MC = NB2_MAX_CONTACTS, MR = NB2_MAX_ROWS;
  ContactWsT<ST> w;
  size_t off = 0;  // in doubles per world
  auto D = [&](size_t cnt) { SP<CR, ST> r; r.p = block + off * ST + lane; off += cnt; return r; };
  auto I = [&](size_t cnt) { SP<int, ST> r; r.p = (int*)(block + off * ST) + lane; off += (cnt + 1) / 2; return r; };
  auto Bt = [&](size_t cnt) { SP<unsigned char, ST> r; r.p = (unsigned char*)(block + off * ST) + lane; off += (cnt + 7) / 8; return r; };
  w.W = D(nb * 12); w.T = D(nb * 12); w.V = D(nb * 6); w.pI = D(nb * 6); w.uI = D(ndof); w.dqd = D(ndof); w.vstar = D(ndof);
  w.cpoint = D(MC * 3); w.cnormal = D(MC * 3); w.cdepth = D(MC); w.cmu = D(MC); w.crest = D(MC);
  w.cbodyA = I(MC); w.cbodyB = I(MC); w.ctype = I(MC); w.cshapeA = I(MC); w.cshapeB = I(MC);
  w.JA = D(MR * 6); w.JB = D(MR * 6); w.b = D(MR); w.lo = D(MR); w.hi = D(MR); w.x = D(MR); w.x0 = D(MR); w.rest = D(MR); w.colnorm = D(MR);
  w.findex = I(MR); w.mapping = I(MR); w.clampIdx = I(MR); w.ubIdx = I(MR);
  w.A = D((size_t)MR * MR); w.Aw = D((size_t)MR * MR); w.L = D((size_t)MR * MR); w.Q = D((size_t)MR * MR); w.Q2 = D((size_t)MR * MR);
  w.v1 = D(MR); w.v2 = D(MR); w.v3 = D(MR); w.v4 = D(MR); w.v5 = D(MR); w.v6 = D(MR); w.v7 = D(MR); w.v8 = D(MR);
  w.i1 = I(MR); w.i2 = I(MR);
  w.st8 = Bt(2 * MR);
  w.lbuf = D((size_t)NB2_CONTACT_LANES * (12 * nb + 2 * ndof));
  w.meta = I(4);
  return w;
}

// ------------------------------------------------------------------ small helpers on raw arrays
template <class P> NB2_HD Xf<CR> xf_from12(P t) {
  Xf<CR> T; T.R_.m00 = t[0]; T.R_.m01 = t[1]; T.R_.m02 = t[2]; T.R_.m10 = t[3]; T.R_.m11 = t[4]; T.R_.m12 = t[5];
  T.R_.m20 = t[6]; T.R_.m21 = t[7]; T.R_.m22 = t[8]; T.p = mk3<CR>(t[9], t[10], t[11]); return T;
}
template <class P> NB2_HD void xf_to12(P t, const Xf<CR>& T) {
  t[0] = T.R_.m00; t[1] = T.R_.m01; t[2] = T.R_.m02; t[3] = T.R_.m10; t[4] = T.R_.m11; t[5] = T.R_.m12;
  t[6] = T.R_.m20; t[7] = T.R_.m21; t[8] = T.R_.m22; t[9] = T.p.x; t[10] = T.p.y; t[11] = T.p.z;
}
NB2_HD Xf<CR> xf_mul(const Xf<CR>& A, const Xf<CR>& B) { Xf<CR> C; C.R_ = mul(A.R_, B.R_); C.p = mul(A.R_, B.p) + A.p; return C; }
NB2_HD V3<CR> xf_apply(const Xf<CR>& A, const V3<CR>& x) { return mul(A.R_, x) + A.p; }
NB2_HD V3<CR> xf_apply_inv(const Xf<CR>& A, const V3<CR>& x) { return mulT(A.R_, x - A.p); }
NB2_HD V3<CR> col3(const M3<CR>& R, int j) { return j == 0 ? mk3<CR>(R.m00, R.m10, R.m20) : (j == 1 ? mk3<CR>(R.m01, R.m11, R.m21) : mk3<CR>(R.m02, R.m12, R.m22)); }
NB2_HD CR get3(const V3<CR>& v, int k) { return k == 0 ? v.x : (k == 1 ? v.y : v.z); }
NB2_HD void set3(V3<CR>& v, int k, CR x) { if (k == 0) v.x = x; else if (k == 1) v.y = x; else v.z = x; }
template <class P> NB2_HD V6<CR> ldv6(P p) { V6<CR> v; v.a = mk3<CR>(p[0], p[1], p[2]); v.l = mk3<CR>(p[3], p[4], p[5]); return v; }
template <class P> NB2_HD void stv6(P p, const V6<CR>& v) { p[0] = v.a.x; p[1] = v.a.y; p[2] = v.a.z; p[3] = v.l.x; p[4] = v.l.y; p[5] = v.l.z; }

typedef ContactOutT<CR> ContactOut;

// ------------------------------------------------------------------ joint transform of body i from the saved stream
NB2_HD Xf<CR> saved_xf(const Nb2ModelDev<CR>& M, int i, const float* st, const CR* sv, size_t B) {
  const int jt = M.jtype[i];
  const CR* s = sv + (size_t)(i * 21) * B;
  if (jt == NB2_JT_REV) return xf_rev(M, i, s[19 * B], s[20 * B]);
  if (jt == NB2_JT_PRIS) return xf_pris(M, i, (CR)st[M.dof_off[i]]);
  CR t12[12];
  for (int k = 0; k < 12; k++) t12[k] = sv[(size_t)(M.nb * 21 + M.free_idx[i] * 33 + 21 + k) * B];
  return xf_from12(t12);
}

// impulse-ABA with the forward's U, psi: body impulses in ws.pI (holding -imp on entry, consumed) -> ws.dqd (joint velocity
// changes) and ws.V (spatial velocity changes).  Only bodies in `mask` are visited: for the impulse TESTS that build A the
// mask holds the ancestors of the contact bodies (every other body has zero bias impulse and its velocity change is never
// read); the final impulse application visits every body.
template <int ST>
NB2_HD void impulse_response(const Nb2ModelDev<CR>& M, const CR* sv, size_t B, const ContactWsT<ST>& ws, unsigned long long mask) {
  const int nb = M.nb;
  for (int i = nb - 1; i >= 0; i--) {
    if (!((mask >> i) & 1ull)) continue;
    const int jt = M.jtype[i], p = M.parent[i], o = M.dof_off[i];
    const CR* s = sv + (size_t)(i * 21) * B;
    const V6<CR> pI = ldv6(ws.pI + 6 * i);
    V6<CR> beta;
    if (jt != NB2_JT_FREE) {
      const CR u = -S_dot(jt, pI);
      ws.uI[o] = u;
      if (p >= 0) beta = pI + sv_ld6<CR>(s, B, 12) * (s[18 * B] * u);
    } else {
      stv6(ws.uI + o, zero6<CR>() - pI);
      if (p >= 0) beta = zero6<CR>();  // pI + I (I^-1 u) = 0: a 6-dof joint absorbs the whole impulse
    }
    if (p >= 0) {
      const V6<CR> pc = dAdInvT(xf_from12(ws.T + 12 * i), beta);
      auto pp = ws.pI + 6 * p;
      pp[0] += pc.a.x; pp[1] += pc.a.y; pp[2] += pc.a.z; pp[3] += pc.l.x; pp[4] += pc.l.y; pp[5] += pc.l.z;
    }
  }
  for (int i = 0; i < nb; i++) {
    if (!((mask >> i) & 1ull)) continue;
    const int jt = M.jtype[i], p = M.parent[i], o = M.dof_off[i];
    const CR* s = sv + (size_t)(i * 21) * B;
    V6<CR> dV = (p >= 0) ? AdInvT(xf_from12(ws.T + 12 * i), ldv6(ws.V + 6 * p)) : zero6<CR>();
    if (jt != NB2_JT_FREE) {
      const CR d = s[18 * B] * (ws.uI[o] - dot(sv_ld6<CR>(s, B, 12), dV));
      ws.dqd[o] = d;
      if (jt == NB2_JT_REV) dV.a.z += d; else dV.l.z += d;
    } else {
      CR i21[21];
      for (int k = 0; k < 21; k++) i21[k] = sv[(size_t)(M.nb * 21 + M.free_idx[i] * 33 + k) * B];
      const V6<CR> d = mul(ldSI<CR, 1>(i21), ldv6(ws.uI + o)) - dV;
      stv6(ws.dqd + o, d);
      dV = dV + d;
    }
    stv6(ws.V + 6 * i, dV);
  }
}

// sum_{j != skip} row[j] * x[j].  One thread: four independent partial sums (the Gauss-Seidel row update is a chain of
// dependent fp64 FMAs otherwise).  A group: lane l takes the columns j = l (mod nl).  The reference sums left to right
// (PgsBoxedLcpSolver.cpp:126-140); the difference is rounding only.
template <class PA, class PX>
NB2_HD CR row_dot_skip(int m, PA row, PX x, int skip, const Grp& g = Grp()) {
  if (g.nl > 1) {
    CR a = 0;
    for (int j = g.cl; j < m; j += g.nl) a += (j == skip) ? CR(0) : row[j] * x[j];
    return g.sum(a);
  }
  CR a0 = 0, a1 = 0, a2 = 0, a3 = 0;
  int j = 0;
  for (; j + 3 < m; j += 4) {
    a0 += (j == skip) ? CR(0) : row[j] * x[j];
    a1 += (j + 1 == skip) ? CR(0) : row[j + 1] * x[j + 1];
    a2 += (j + 2 == skip) ? CR(0) : row[j + 2] * x[j + 2];
    a3 += (j + 3 == skip) ? CR(0) : row[j + 3] * x[j + 3];
  }
  for (; j < m; j++) a0 += (j == skip) ? CR(0) : row[j] * x[j];
  return (a0 + a1) + (a2 + a3);
}

// ------------------------------------------------------------------ dense helpers (m x m, stride m)
template <class PA, class PX, class PB_, class PH, class PL, class PF>
NB2_HD bool lcp_valid(int m, PA A, PX x, PB_ b, PH hi, PL lo, PF fi, bool ignoreFriction, const Grp& g = Grp()) {
  for (int i = 0; i < m; i++) {
    const CR v = row_dot_skip(m, A + i * m, x, -1, g) - b[i];
    CR up = hi[i], low = lo[i];
    if (fi[i] != -1) { if (ignoreFriction) { if (x[i] != 0) return false; continue; } up *= x[fi[i]]; low *= x[fi[i]]; }
    const CR tol = 1e-5;
    if (nb2_abs(low) < tol && nb2_abs(up) < tol && nb2_abs(x[i]) < tol) {}
    else if (nb2_abs(x[i] - low) < tol) { if (v < -tol) return false; }
    else if (nb2_abs(x[i] - up) < tol) { if (v > tol) return false; }
    else if (x[i] > low && x[i] < up) { if (nb2_abs(v) > tol) return false; }
    else return false;
  }
  return true;
}

// dst[i] = src[i], i < n.  The GPU issues in order: a copy loop "load, store, load, store" keeps ONE load in flight (each store
// waits for its load), so the loads of a batch are issued before its stores.  dst and src must not overlap.
template <class PDst, class PSrc>
NB2_HD void copy_n(PDst dst, PSrc src, int n) {
  int i = 0;
  for (; i + 7 < n; i += 8) {
    CR t[8];
#pragma unroll
    for (int u = 0; u < 8; u++) t[u] = src[i + u];
#pragma unroll
    for (int u = 0; u < 8; u++) dst[i + u] = t[u];
  }
  for (; i < n; i++) dst[i] = src[i];
}

// minimum-norm least squares x = Q^+ rhs for an n x n matrix.  symmetric PSD Q: rank-revealing pivoted Cholesky
// Q = P L L^T P^T (L: n x r) and Q^+ = L (L^T L)^-2 L^T ; general Q: x = (Q^T Q)^+ Q^T rhs through the same routine.
// work: G (n*n, destroyed copy), Lf (n*n), t1..t3 (n), perm (n)
template <class PQ, class PR, class PX, class PG, class PL, class PT1, class PT2, class PP>
NB2_HD void pinv_psd(int n, PQ Qin, PR rhs, PX x, PG G, PL Lf, PT1 t1, PT2 t2, PP perm) {
  copy_n(G, Qin, n * n);
  for (int i = 0; i < n; i++) perm[i] = i;
  CR dmax0 = 0;
  auto dg = t2;  // running diagonal of the Schur complement: dg[i] = G[i][i] - sum_{j<k} Lf[i][j]^2 (same subtraction order as a fresh sum)
  for (int i = 0; i < n; i++) { const CR d = G[i * n + i]; dg[i] = d; dmax0 = d > dmax0 ? d : dmax0; }
  const CR tol = dmax0 * 1e-12;
  int r = 0;
  // pivoted Cholesky; Lf[row * n + k]
  for (int k = 0; k < n; k++) {
    int piv = -1; CR best = tol;
    for (int i = k; i < n; i++) { const CR d = dg[perm[i]]; if (d > best) { best = d; piv = i; } }
    if (piv < 0) break;
    { const int t = perm[k]; perm[k] = perm[piv]; perm[piv] = t; }
    const int pk = perm[k];
    const CR lkk = nb2_sqrt(best);
    Lf[pk * n + k] = lkk;
    for (int i = k + 1; i < n; i++) {
      const int pi = perm[i];
      CR s = G[pi * n + pk];
#pragma unroll 4
      for (int j = 0; j < k; j++) s -= Lf[pi * n + j] * Lf[pk * n + j];
      const CR l = s / lkk;
      Lf[pi * n + k] = l;
      dg[pi] -= l * l;
    }
    r++;
  }
  for (int i = 0; i < n; i++) x[i] = 0;
  if (r == 0) return;
  if (r == n) {  // full rank: Q^+ = Q^-1 = P L^-T L^-1 P^T, two triangular solves (rows of L in pivot order)
    for (int i = 0; i < n; i++) {
      const int pi = perm[i];
      CR s = rhs[pi];
#pragma unroll 4
      for (int k = 0; k < i; k++) s -= Lf[pi * n + k] * t1[k];
      t1[i] = s / Lf[pi * n + i];
    }
    for (int i = n - 1; i >= 0; i--) {
      CR s = t1[i];
#pragma unroll 4
      for (int k = i + 1; k < n; k++) s -= Lf[perm[k] * n + i] * t2[k];
      t2[i] = s / Lf[perm[i] * n + i];
    }
    for (int i = 0; i < n; i++) x[perm[i]] = t2[i];
    return;
  }
  // rows of L for indices not yet pivoted at step k are valid for columns < r; rows of pivoted indices have zeros above: fill
  for (int k = 0; k < r; k++) for (int j = k + 1; j < r; j++) Lf[perm[k] * n + j] = 0;
  // M = L^T L (r x r) in G ; y = L^T rhs
  for (int a = 0; a < r; a++) {
    CR ya = 0;
#pragma unroll 4
    for (int i = 0; i < n; i++) ya += Lf[i * n + a] * rhs[i];
    t1[a] = ya;
    for (int c = a; c < r; c++) {
      CR s = 0;
#pragma unroll 4
      for (int i = 0; i < n; i++) s += Lf[i * n + a] * Lf[i * n + c];
      G[a * n + c] = s; G[c * n + a] = s;
    }
  }
  // z = M^-2 y  via Cholesky of M (SPD r x r), two solves
  for (int j = 0; j < r; j++) {
    CR d = G[j * n + j];
#pragma unroll 4
    for (int k = 0; k < j; k++) d -= G[j * n + k] * G[j * n + k];
    d = nb2_sqrt(d); G[j * n + j] = d;
    for (int i = j + 1; i < r; i++) {
      CR s = G[i * n + j];
#pragma unroll 4
      for (int k = 0; k < j; k++) s -= G[i * n + k] * G[j * n + k];
      G[i * n + j] = s / d;
    }
  }
  for (int rep = 0; rep < 2; rep++) {
    for (int i = 0; i < r; i++) {
      CR s = t1[i];
#pragma unroll 4
      for (int k = 0; k < i; k++) s -= G[i * n + k] * t2[k];
      t2[i] = s / G[i * n + i];
    }
    for (int i = r - 1; i >= 0; i--) {
      CR s = t2[i];
#pragma unroll 4
      for (int k = i + 1; k < r; k++) s -= G[k * n + i] * t1[k];
      t1[i] = s / G[i * n + i];
    }
  }
  for (int i = 0; i < n; i++) {
    CR s = 0;
#pragma unroll 4
    for (int a = 0; a < r; a++) s += Lf[i * n + a] * t1[a];
    x[i] = s;
  }
}

// classification (constructMatrices) + standardisation; x is updated in place when the standardised solution is valid.
// returns true when the results are standardised.
template <int ST, class PD>
NB2_HD bool classify_once(int m, PD A, PD x, PD b, PD lo, PD hi, SP<int, ST> fi, PD colnorm,
                          bool ignoreFriction, const ContactWsT<ST>& ws, bool* again) {
  *again = false;
  auto mapping = ws.mapping; auto clampIdx = ws.clampIdx; auto ubIdx = ws.ubIdx;
  int nCl = 0, nUb = 0;
  for (int j = 0; j < m; j++) { mapping[j] = fi[j]; clampIdx[j] = -1; ubIdx[j] = -1; }
  for (int j = 0; j < m; j++) {
    if (colnorm[j] < 1e-9) { mapping[j] = NB2_MAP_NOT_CLAMPING; continue; }
    CR up = hi[j], low = lo[j];
    const int fp = fi[j];
    if (fp != -1) { up *= x[fp]; low *= x[fp]; }
    if (nb2_abs(x[j]) < 1e-6) {
      if (fp != -1) {
        if (nb2_abs(x[fp]) < 1e-6) mapping[j] = NB2_MAP_NOT_CLAMPING;
        else if (ignoreFriction) mapping[j] = NB2_MAP_NOT_CLAMPING;
        else { mapping[j] = NB2_MAP_CLAMPING; clampIdx[j] = nCl++; }
      } else mapping[j] = NB2_MAP_NOT_CLAMPING;
      continue;
    }
    const CR tie = 1e-5;
    if ((x[j] > low + tie && x[j] < up - tie) || (low - x[j] > 1e-2 || x[j] - up > 1e-2)) { mapping[j] = NB2_MAP_CLAMPING; clampIdx[j] = nCl++; }
    else if (fp != -1 && nb2_abs(x[fp]) > 1e-9 && colnorm[fp] > 1e-9 && ((fp > j) || mapping[fp] == NB2_MAP_CLAMPING)) { mapping[j] = fp; ubIdx[j] = nUb++; }
    else mapping[j] = NB2_MAP_NOT_CLAMPING;
  }
  // ---- opportunisticallyStandardizeResults
  if (nCl == 0) {
    for (int i = 0; i < m; i++) ws.v1[i] = 0;
    if (lcp_valid(m, A, ws.v1, b, hi, lo, fi, ignoreFriction, ws.grp)) { for (int i = 0; i < m; i++) x[i] = 0; return true; }
    return false;
  }
  auto cl = ws.i1; auto ub = ws.i2;
  for (int j = 0; j < m; j++) { if (clampIdx[j] >= 0) cl[clampIdx[j]] = j; if (ubIdx[j] >= 0) ub[ubIdx[j]] = j; }
  // E(u, clampIdx[fp]) = hi or lo of the row; Q = A[cl,cl] + A[cl,ub] E
  auto Q = ws.Q; auto bc = ws.v2; auto orig = ws.v3; auto fc = ws.v4;
  for (int r = 0; r < nCl; r++) {
    bc[r] = b[cl[r]]; orig[r] = x[cl[r]];
    const int rowr = cl[r] * m;
    for (int c0 = 0; c0 < nCl; c0 += 4) {  // batched gathers (see copy_n)
      CR t[4];
#pragma unroll
      for (int u = 0; u < 4; u++) if (c0 + u < nCl) t[u] = A[rowr + cl[c0 + u]];
#pragma unroll
      for (int u = 0; u < 4; u++) if (c0 + u < nCl) Q[r * nCl + c0 + u] = t[u];
    }
  }
  for (int u = 0; u < nUb; u++) {
    const int j = ub[u], fp = mapping[j];
    const CR up = x[fp] * hi[j], low = x[fp] * lo[j];
    const CR e = (nb2_abs(x[j] - up) < nb2_abs(x[j] - low)) ? hi[j] : lo[j];
    const int c = clampIdx[fp];
    for (int r = 0; r < nCl; r++) Q[r * nCl + c] += A[cl[r] * m + j] * e;
  }
  if (nUb == 0) pinv_psd(nCl, Q, bc, fc, ws.Aw, ws.L, ws.v5, ws.v6, ws.i2);
  else {
    // general Q: x = (Q^T Q)^+ Q^T b
    auto QtQ = ws.Q2;
    auto Qtb = ws.v7;
    for (int a = 0; a < nCl; a++) {
      CR s = 0; for (int r = 0; r < nCl; r++) s += Q[r * nCl + a] * bc[r];
      Qtb[a] = s;
      for (int c = 0; c < nCl; c++) { CR t = 0; for (int r = 0; r < nCl; r++) t += Q[r * nCl + a] * Q[r * nCl + c]; QtQ[a * nCl + c] = t; }
    }
    pinv_psd(nCl, QtQ, Qtb, fc, ws.Aw, ws.L, ws.v5, ws.v6, ws.i2);
    // ws.i2 (ub list) was clobbered by the permutation: rebuild
    for (int j = 0; j < m; j++) if (ubIdx[j] >= 0) ub[ubIdx[j]] = j;
  }
  bool anyNewlyNotClamping = false;
  auto nx = ws.v8;
  for (int i = 0; i < m; i++) {
    nx[i] = 0;
    if (clampIdx[i] != -1) {
      nx[i] = fc[clampIdx[i]];
      if (nb2_abs(fc[clampIdx[i]]) < 1e-6 && nb2_abs(x[i]) > 1e-6 && fi[i] == -1) anyNewlyNotClamping = true;
    }
    if (ubIdx[i] != -1) {
      const int fp = fi[i];
      const CR origMult = orig[clampIdx[fp]] / x[i];
      const CR clean = (nb2_abs(origMult - hi[i]) < nb2_abs(origMult - lo[i])) ? hi[i] : lo[i];
      nx[i] = fc[clampIdx[fp]] * clean;
    }
  }
  if (lcp_valid(m, A, nx, b, hi, lo, fi, ignoreFriction, ws.grp)) {
    for (int i = 0; i < m; i++) x[i] = nx[i];
    *again = anyNewlyNotClamping;  // a previously clamping normal row dropped to ~0: re-classify (:283-331)
    return true;
  }
  return false;
}
template <int ST, class PD>
NB2_HD bool classify_and_standardize(int m, PD A, PD x, PD b, PD lo, PD hi, SP<int, ST> fi, PD colnorm,
                                     bool ignoreFriction, const ContactWsT<ST>& ws) {
  bool ok = false, again = false;
  for (int it = 0; it < 6; it++) {
    ok = classify_once<ST>(m, A, x, b, lo, hi, fi, colnorm, ignoreFriction, ws, &again);
    if (!ok || !again) break;
  }
  return ok;
}

#ifdef __CUDA_ARCH__
// pgs_solve for a group of 8 threads: lane l keeps x[l], x[l + 8], ... in registers and owns the same columns of A, so a
// row update is <= 6 loads + FMAs per lane, one butterfly sum, and two broadcasts (old x_i, x of the friction row's normal);
// nothing on the dependent chain goes through memory.  Same sweeps, tolerances and clamping as the one-thread version below.
template <class PA, class PX, class PB_, class PL, class PH, class PF, class PS>
__device__ bool pgs_solve_group8(int m, PA A, PX x, PB_ b, PL lo, PH hi, PF fi, PS skip, const Grp& g) {
  constexpr int KX = NB2_MAX_ROWS / 8;
  const CR dxTol = 1e-6, relTol = 1e-3, epsDiv = 1e-9;
  const int base = __ffs(g.mask) - 1;
  CR xr[KX];
#pragma unroll
  for (int k = 0; k < KX; k++) { const int j = g.cl + 8 * k; xr[k] = (j < m) ? x[j] : 0.0; }
  auto xget = [&](int j) {
    const int kk = j >> 3;
    CR v = 0;
#pragma unroll
    for (int k = 0; k < KX; k++) if (k == kk) v = xr[k];
    return __shfl_sync(g.mask, v, base + (j & 7));
  };
  auto xset = [&](int j, CR v) {
    if ((j & 7) != g.cl) return;
    const int kk = j >> 3;
#pragma unroll
    for (int k = 0; k < KX; k++) if (k == kk) xr[k] = v;
  };
  auto rowdot = [&](int i) {
    CR a = 0;
#pragma unroll
    for (int k = 0; k < KX; k++) { const int j = g.cl + 8 * k; if (j < m && j != i) a += A[i * m + j] * xr[k]; }
    return g.sum(a);
  };
  auto finish = [&](bool result) {
#pragma unroll
    for (int k = 0; k < KX; k++) { const int j = g.cl + 8 * k; if (j < m) x[j] = xr[k]; }
    __syncwarp(g.mask);  // every lane wrote its own columns of x: make them visible to the whole group
    return result;
  };
  bool term = true;
  for (int i = 0; i < m; i++) {
    const CR aii = A[i * m + i];
    if (aii < epsDiv) { xset(i, 0.0); skip[i] = 1; continue; }
    skip[i] = 0;
    const CR old_x = xget(i);
    const CR nx = (b[i] - rowdot(i)) / aii;
    CR hi_t = hi[i], lo_t = lo[i];
    const int f = fi[i];
    if (f >= 0) { hi_t = hi[i] * xget(f); lo_t = -hi_t; }
    const CR xi = nx > hi_t ? hi_t : (nx < lo_t ? lo_t : nx);
    xset(i, xi);
    if (term && nb2_abs(xi - old_x) > dxTol) term = false;
  }
  if (term) return finish(true);
  for (int i = 0; i < m; i++) if (!skip[i]) {
    const CR dm = 1.0 / A[i * m + i];
    const CR bi = b[i] * dm;
    __syncwarp(g.mask);            // every lane has read A[i][i] and b[i] before anyone rescales them
    b[i] = bi;                      // (same value from every lane)
    for (int j = g.cl; j < m; j += 8) A[i * m + j] *= dm;  // each lane rescales the columns it owns
  }
  for (int iter = 1; iter < 30; iter++) {
    term = true;
    for (int i = 0; i < m; i++) {
      if (skip[i]) continue;
      const CR old_x = xget(i);
      const CR nx = b[i] - rowdot(i);
      CR hi_t = hi[i], lo_t = lo[i];
      const int f = fi[i];
      if (f >= 0) { hi_t = hi[i] * xget(f); lo_t = -hi_t; }
      const CR xi = nx > hi_t ? hi_t : (nx < lo_t ? lo_t : nx);
      xset(i, xi);
      if (term && nb2_abs(xi) > epsDiv) { if (nb2_abs((xi - old_x) / xi) > relTol) term = false; }
    }
    if (term) break;
  }
  return finish(term);
}
#endif

// PgsBoxedLcpSolver::solve with Option(30, 1e-6, 1e-3, 1e-9, false); A (m x m) and b are clobbered
template <class PA, class PX, class PB_, class PL, class PH, class PF, class PS>
NB2_HD bool pgs_solve(int m, PA A, PX x, PB_ b, PL lo, PH hi, PF fi, PS skip, const Grp& g = Grp()) {
#ifdef __CUDA_ARCH__
  if (g.nl == 8) return pgs_solve_group8(m, A, x, b, lo, hi, fi, skip, g);
#endif
  const CR dxTol = 1e-6, relTol = 1e-3, epsDiv = 1e-9;
  bool term = true;
  for (int i = 0; i < m; i++) {
    skip[i] = 0;
    if (A[i * m + i] < epsDiv) { x[i] = 0.0; skip[i] = 1; continue; }
    const CR old_x = x[i];
    CR nx = b[i] - row_dot_skip(m, A + i * m, x, i, g);
    nx /= A[i * m + i];
    CR hi_t = hi[i], lo_t = lo[i];
    if (fi[i] >= 0) { hi_t = hi[i] * x[fi[i]]; lo_t = -hi_t; }
    x[i] = nx > hi_t ? hi_t : (nx < lo_t ? lo_t : nx);
    if (term && nb2_abs(x[i] - old_x) > dxTol) term = false;
  }
  if (term) return true;
  for (int i = 0; i < m; i++) if (!skip[i]) { const CR dm = 1.0 / A[i * m + i]; b[i] *= dm; for (int j = 0; j < m; j++) A[i * m + j] *= dm; }
  for (int iter = 1; iter < 30; iter++) {
    term = true;
    for (int i = 0; i < m; i++) {
      if (skip[i]) continue;
      const CR old_x = x[i];
      const CR nx = b[i] - row_dot_skip(m, A + i * m, x, i, g);
      CR hi_t = hi[i], lo_t = lo[i];
      if (fi[i] >= 0) { hi_t = hi[i] * x[fi[i]]; lo_t = -hi_t; }
      x[i] = nx > hi_t ? hi_t : (nx < lo_t ? lo_t : nx);
      if (term && nb2_abs(x[i]) > epsDiv) { if (nb2_abs((x[i] - old_x) / x[i]) > relTol) term = false; }
    }
    if (term) break;
  }
  return term;
}

// LCPUtils::reduce (LCPUtils.cpp:144-201) + mergeLCPColumns (:346-444): near-identical columns (same bounds, findex,
// |b_a - b_b| < 1e-4, ||A_a - A_b||^2 < 1e-4) are merged one pair at a time — column a doubled, row/column b dropped —
// until none is left; `map` (n0 x n, row-major with the CURRENT n as stride) accumulates the un-merge map x = map * x_r.
// Everything is compacted in place (destination index <= source index in row-major order). Returns the reduced size.
template <class PD, class PI>
NB2_HD int lcp_reduce(int n0, PD A, PD x, PD b, PD lo, PD hi, PI fi, PD map) {
  int n = n0;
  for (int i = 0; i < n0; i++) for (int j = 0; j < n0; j++) map[i * n0 + j] = (i == j) ? 1.0 : 0.0;
  while (true) {
    int ca = -1, cb = -1;
    for (int a = 0; a < n - 1 && ca < 0; a++) for (int c = a + 1; c < n; c++) {
      if (fi[a] != fi[c] || hi[a] != hi[c] || lo[a] != lo[c] || !(nb2_abs(b[a] - b[c]) < 1e-4)) continue;
      CR d2 = 0;
      for (int r = 0; r < n; r++) { const CR d = A[r * n + a] - A[r * n + c]; d2 += d * d; }
      if (d2 < 1e-4) { ca = a; cb = c; break; }
    }
    if (ca < 0) break;
    for (int r = 0, ri = 0; r < n; r++) {
      if (r == cb) continue;
      for (int i = 0, ci = 0; i < n; i++) { if (i == cb) continue; A[ri * (n - 1) + ci] = A[r * n + i] * (i == ca ? 2.0 : 1.0); ci++; }
      ri++;
    }
    for (int i = 0, ni = 0; i < n; i++) {
      if (i == cb) continue;
      x[ni] = x[i]; b[ni] = b[i]; lo[ni] = lo[i]; hi[ni] = hi[i];
      const int f = fi[i];
      fi[ni] = (f < cb) ? f : (f == cb ? ca : f - 1);
      ni++;
    }
    for (int r = 0; r < n0; r++) {
      const CR vb = map[r * n + cb];
      for (int i = 0, ni = 0; i < n; i++) { if (i == cb) continue; map[r * (n - 1) + ni] = map[r * n + i]; ni++; }
      map[r * (n - 1) + ca] += vb;
    }
    n--;
  }
  return n;
}

// The solve chain of BoxedLcpConstraintSolver::solveLcp (:352-789) on the problem held in the workspace (A, b, lo, hi,
// findex of size m): warm start -> short-circuit classification -> [reduce] Dantzig -> cfm + [reduce] PGS -> friction drop
// -> classification / standardisation.  x_cached: last step's solution when it has the same size, else nullptr
// (LCPUtils::guessSolution).  Leaves x in ws.x and the labels in ws.mapping; returns the status bits.
template <int ST>
NB2_HD int lcp_chain_ws(int m, const ContactWsT<ST>& ws, CR fallback_cfm, const CR* x_cached) {
  int status = 0;
  auto A = ws.A;
  for (int c = 0; c < m; c++) { CR sn = 0; for (int r = 0; r < m; r++) sn += A[r * m + c] * A[r * m + c]; ws.colnorm[c] = sn; }
  auto b = ws.b; auto lo = ws.lo; auto hi = ws.hi; auto fi = ws.findex; auto x = ws.x; auto x0 = ws.x0;
  // ---- warm start: cached solution if it has the same size, else LCPUtils::guessSolution (LCPUtils.cpp:86-140)
  if (x_cached) { for (int i = 0; i < m; i++) x0[i] = x_cached[i]; }
  else {
    int ng = 0;
    for (int i = 0; i < m; i++) { x0[i] = 0; if (fi[i] == -1) { if (b[i] > 0) ws.i1[ng++] = i; } else ws.i1[ng++] = i; }
    if (ng > 0) {
      for (int r = 0; r < ng; r++) { ws.v1[r] = b[ws.i1[r]]; for (int c = 0; c < ng; c++) ws.Q[r * ng + c] = A[ws.i1[r] * m + ws.i1[c]]; }
      pinv_psd(ng, ws.Q, ws.v1, ws.v2, ws.Aw, ws.L, ws.v5, ws.v6, ws.i2);
      for (int r = 0; r < ng; r++) x0[ws.i1[r]] = ws.v2[r];
    }
  }
  for (int i = 0; i < m; i++) x[i] = x0[i];
  // ---- solve chain (BoxedLcpConstraintSolver.cpp:352-789)
  bool success = classify_and_standardize<ST>(m, A, x, b, lo, hi, fi, ws.colnorm, false, ws);
  const bool shortCircuit = success;
  bool ignoredFriction = false;
  if (success) status |= NB2_ST_SHORTCIRCUIT;
  else {
    status |= NB2_ST_DANTZIG;
    copy_n(ws.Aw, A, m * m);
    for (int i = 0; i < m; i++) { ws.v1[i] = b[i]; ws.v2[i] = lo[i]; ws.v3[i] = hi[i]; ws.i1[i] = fi[i]; ws.v4[i] = x[i]; }
    const int mr = lcp_reduce(m, ws.Aw, ws.v4, ws.v1, ws.v2, ws.v3, ws.i1, ws.Q2);  // :596 reduce before the Dantzig solve
    if (mr < m) status |= NB2_ST_MERGED;
    DantzigWorkT<SP<CR, ST>, SP<int, ST>, SP<unsigned char, ST>> W;
    W.A = ws.Aw; W.x = ws.v4; W.b = ws.v1; W.w = ws.v5; W.lo = ws.v2; W.hi = ws.v3; W.L = ws.L; W.d = ws.v6; W.delta_x = ws.v7; W.delta_w = ws.v8;
    W.Dell = ws.Q; W.ell = ws.Q + mr; W.tmp = ws.Q + 2 * mr; W.findex = ws.i1; W.p = ws.i2; W.C = ws.clampIdx; W.state = ws.st8;
    const int rc = dantzig_solve(W, mr, true);
    success = (rc == 1);
    if (success) {
      for (int i = 0; i < m; i++) { CR v = 0; for (int c = 0; c < mr; c++) v += ws.Q2[i * mr + c] * ws.v4[c]; x[i] = v; }  // x = mapOut * x_reduced
      if (!lcp_valid(m, A, x, b, hi, lo, fi, false, ws.grp)) success = false;
    }
    if (!success) status |= NB2_ST_DANTZIG_FAILED;
  }
  { bool nan = false; for (int i = 0; i < m; i++) if (x[i] != x[i]) nan = true; if (nan) { success = false; for (int i = 0; i < m; i++) x[i] = 0; status |= NB2_ST_NAN; } }
  if (!success) {
    for (int i = 0; i < m; i++) A[i * m + i] += fallback_cfm;  // :539-547 (both backups get the cfm; colnorms were taken before)
    status |= NB2_ST_PGS;
    copy_n(ws.Aw, A, m * m);
    for (int i = 0; i < m; i++) { ws.v1[i] = b[i]; ws.v2[i] = lo[i]; ws.v3[i] = hi[i]; ws.i1[i] = fi[i]; ws.v4[i] = x0[i]; }
    const int mr = lcp_reduce(m, ws.Aw, ws.v4, ws.v1, ws.v2, ws.v3, ws.i1, ws.Q2);  // :551-557 the backup problem is reduced too
    if (mr < m) status |= NB2_ST_MERGED;
    success = pgs_solve(mr, ws.Aw, ws.v4, ws.v1, ws.v2, ws.v3, ws.i1, ws.st8, ws.grp);
    if (success) {
      for (int i = 0; i < m; i++) { CR v = 0; for (int c = 0; c < mr; c++) v += ws.Q2[i * mr + c] * ws.v4[c]; x[i] = v; }
      if (!lcp_valid(m, A, x, b, hi, lo, fi, false, ws.grp)) success = false;
    }
  }
  if (!success) {
    ignoredFriction = true;
    status |= NB2_ST_FRICTION_DROPPED;
    int k = 0;
    for (int i = 0; i < m; i++) if (fi[i] == -1) ws.i1[k++] = i;
    for (int r = 0; r < k; r++) { ws.v1[r] = b[ws.i1[r]]; ws.v2[r] = lo[ws.i1[r]]; ws.v3[r] = hi[ws.i1[r]]; ws.v4[r] = 0; ws.i2[r] = -1; for (int c = 0; c < k; c++) ws.Aw[r * k + c] = A[ws.i1[r] * m + ws.i1[c]]; }
    pgs_solve(k, ws.Aw, ws.v4, ws.v1, ws.v2, ws.v3, ws.i2, ws.st8, ws.grp);
    for (int i = 0; i < m; i++) x[i] = 0;
    for (int r = 0; r < k; r++) x[ws.i1[r]] = ws.v4[r];
  }
  { bool nan = false; for (int i = 0; i < m; i++) if (x[i] != x[i]) nan = true; if (nan) { for (int i = 0; i < m; i++) x[i] = 0; status |= NB2_ST_NAN; } }
  if (!shortCircuit) {
    for (int i = 0; i < m; i++) ws.v7[i] = x[i];
    // classify works on x in place and only keeps the standardised x when valid
    if (!classify_and_standardize<ST>(m, A, x, b, lo, hi, fi, ws.colnorm, ignoredFriction, ws)) { status |= NB2_ST_NOT_STANDARDIZED; }
  }
  return status;
}

// =====================================================================================================
// the contact stage of one world.  `out` holds [q+ ; v*] on entry (written by the ABA kernel) and [q+ ; v+] on exit.
// x_io: cached LCP solution (NB2_MAX_ROWS doubles), m_io: its size (-1 none) -> new solution / size.
// =====================================================================================================
template <int ST>
NB2_HD void contact_phase0(const Nb2ModelDev<CR>& M, const Nb2ContactDev& C, const float* st, float* out, const CR* sv, size_t B,
                           CR* wsblock, int lane, CR* x_io, int* m_io, int* labels_out, int* status_out, int* nc_out, float* cinfo_out,
                           CR* crec) {
  const int nb = M.nb, n = M.ndof;
  const ContactWsT<ST> ws = carve_ws<ST>(wsblock, lane, nb, n);
  const int kQdd = nb * 21 + M.nfree * 33;
  const CR dt = M.dt;
  int status = 0;
  // ---- world transforms and body velocities at v* = v + dt qdd
  for (int i = 0; i < nb; i++) {
    const int jt = M.jtype[i], p = M.parent[i], o = M.dof_off[i];
    const Xf<CR> T = saved_xf(M, i, st, sv, B);
    const Xf<CR> W = (p >= 0) ? xf_mul(xf_from12(ws.W + 12 * p), T) : T;
    xf_to12(ws.W + 12 * i, W);
    xf_to12(ws.T + 12 * i, T);
    V6<CR> V = (p >= 0) ? AdInvT(T, ldv6(ws.V + 6 * p)) : zero6<CR>();
    if (jt != NB2_JT_FREE) {
      const CR vs = (CR)st[n + o] + dt * sv[(size_t)(kQdd + o) * B];
      ws.uI[o] = vs;  // v* kept in uI until the impulse sweeps start
      if (jt == NB2_JT_REV) V.a.z += vs; else V.l.z += vs;
    } else {
      V6<CR> vs;
      vs.a = mk3<CR>((CR)st[n + o] + dt * sv[(size_t)(kQdd + o) * B], (CR)st[n + o + 1] + dt * sv[(size_t)(kQdd + o + 1) * B], (CR)st[n + o + 2] + dt * sv[(size_t)(kQdd + o + 2) * B]);
      vs.l = mk3<CR>((CR)st[n + o + 3] + dt * sv[(size_t)(kQdd + o + 3) * B], (CR)st[n + o + 4] + dt * sv[(size_t)(kQdd + o + 4) * B], (CR)st[n + o + 5] + dt * sv[(size_t)(kQdd + o + 5) * B]);
      stv6(ws.uI + o, vs);
      V = V + vs;
    }
    stv6(ws.V + 6 * i, V);
  }
  for (int d = 0; d < n; d++) ws.vstar[d] = ws.uI[d];

  // ---- contact generation over the precomputed pair list (reference order)
  int nc = 0;
  for (int pi = 0; pi < C.npairs; pi++) {
    const int sa = C.pair_a[pi], sb = C.pair_b[pi];
    const int ba = C.shape_body[sa], bb = C.shape_body[sb];
    const Xf<CR> Ta = (ba >= 0) ? xf_mul(xf_from12(ws.W + 12 * ba), xf_from12(C.shape_T[sa])) : xf_from12(C.shape_T[sa]);
    const Xf<CR> Tb = (bb >= 0) ? xf_mul(xf_from12(ws.W + 12 * bb), xf_from12(C.shape_T[sb])) : xf_from12(C.shape_T[sb]);
    const int ta = C.shape_type[sa], tb = C.shape_type[sb];
    const V3<CR> da = mk3<CR>(C.shape_dims[sa][0], C.shape_dims[sa][1], C.shape_dims[sa][2]);
    const V3<CR> db = mk3<CR>(C.shape_dims[sb][0], C.shape_dims[sb][1], C.shape_dims[sb][2]);
    ContactOut co[8];
    int k = 0;
    if (ta == 0 && tb == 0) k = collide_box_box(da, Ta, db, Tb, C.clip_depth, co);
    else if (ta == 0 && tb == 1) k = collide_box_sphere(da, Ta, db.x, Tb, C.clip_depth, 0, false, co);
    else if (ta == 1 && tb == 0) k = collide_box_sphere(db, Tb, da.x, Ta, C.clip_depth, 0, true, co);
    else if ((ta == 0 && tb == 2) || (ta == 2 && tb == 0)) {
      const bool boxFirst = (ta == 0);
      const Xf<CR>& Tc = boxFirst ? Tb : Ta; const Xf<CR>& Tbx = boxFirst ? Ta : Tb;
      const V3<CR> bdim = boxFirst ? da : db;
      const CR r = boxFirst ? db.x : da.x, h = boxFirst ? db.y : da.y;
      CR dep[2]; Xf<CR> Tend[2];
      for (int e = 0; e < 2; e++) {
        Tend[e] = Tc; Tend[e].p = xf_apply(Tc, mk3<CR>(0, 0, e == 0 ? h / 2 : -h / 2));
        const V3<CR> pl = xf_apply_inv(Tbx, Tend[e].p);
        V3<CR> q = pl; bool inside = true;
        for (int kk = 0; kk < 3; kk++) { const CR hk = 0.5 * get3(bdim, kk), v = get3(q, kk); if (v < -hk) { set3(q, kk, -hk); inside = false; } if (v > hk) { set3(q, kk, hk); inside = false; } }
        if (inside) { CR mn = 1e300; for (int kk = 0; kk < 3; kk++) { const CR v = 0.5 * get3(bdim, kk) - nb2_abs(get3(pl, kk)); mn = v < mn ? v : mn; } dep[e] = mn + r; }
        else { const V3<CR> dd = pl - q; dep[e] = r - nb2_sqrt(dot(dd, dd)); }
      }
      if ((dep[0] > dep[1] ? dep[0] : dep[1]) >= 0) {
        if (nb2_abs(dep[0] - dep[1]) < 1e-9) status |= NB2_ST_UNSUPPORTED_GEOMETRY;
        else {
          const int e = dep[0] > dep[1] ? 0 : 1;
          k = collide_box_sphere(bdim, Tbx, r, Tend[e], C.clip_depth, e == 0 ? 1 : 2, !boxFirst, co);
        }
      }
    } else status |= NB2_ST_UNSUPPORTED_GEOMETRY;
    for (int c = 0; c < k; c++) {
      if (dot(co[c].normal, co[c].normal) < 1e-12) continue;
      if (co[c].depth < 0.0 || co[c].depth > C.clip_depth) continue;
      if (ba < 0 && bb < 0) continue;
      if (nc >= NB2_MAX_CONTACTS) { status |= NB2_ST_CONTACT_OVERFLOW; continue; }
      ws.cpoint[3 * nc] = co[c].point.x; ws.cpoint[3 * nc + 1] = co[c].point.y; ws.cpoint[3 * nc + 2] = co[c].point.z;
      ws.cnormal[3 * nc] = co[c].normal.x; ws.cnormal[3 * nc + 1] = co[c].normal.y; ws.cnormal[3 * nc + 2] = co[c].normal.z;
      ws.cdepth[nc] = co[c].depth; ws.cbodyA[nc] = ba; ws.cbodyB[nc] = bb; ws.ctype[nc] = co[c].type; ws.cshapeA[nc] = sa; ws.cshapeB[nc] = sb;
      ws.cmu[nc] = C.shape_mu[sa] < C.shape_mu[sb] ? C.shape_mu[sa] : C.shape_mu[sb];
      ws.crest[nc] = C.shape_rest[sa] * C.shape_rest[sb];
      nc++;
    }
  }
  *nc_out = nc;
  if (cinfo_out) for (int c = 0; c < nc; c++) {
    float* o = cinfo_out + 10 * c;
    o[0] = (float)ws.cpoint[3 * c]; o[1] = (float)ws.cpoint[3 * c + 1]; o[2] = (float)ws.cpoint[3 * c + 2];
    o[3] = (float)ws.cnormal[3 * c]; o[4] = (float)ws.cnormal[3 * c + 1]; o[5] = (float)ws.cnormal[3 * c + 2];
    o[6] = (float)ws.cdepth[c]; o[7] = (float)C.shape_orig_body[ws.cshapeA[c]]; o[8] = (float)C.shape_orig_body[ws.cshapeB[c]]; o[9] = (float)ws.ctype[c];
  }
  // ---- rows
  int m = 0;
  for (int c = 0; c < nc; c++) {
    const CR mu = ws.cmu[c], e = ws.crest[c];
    const bool fric = mu > 1e-3, bounce = e > 1e-3;
    const V3<CR> nrm = mk3<CR>(ws.cnormal[3 * c], ws.cnormal[3 * c + 1], ws.cnormal[3 * c + 2]);
    const V3<CR> pt = mk3<CR>(ws.cpoint[3 * c], ws.cpoint[3 * c + 1], ws.cpoint[3 * c + 2]);
    V3<CR> dirs[3]; dirs[0] = nrm;
    if (fric) {  // getTangentBasisMatrixODE with first frictional direction = UnitZ (ContactConstraint.cpp:734-795)
      V3<CR> t = cross(mk3<CR>(0, 0, 1), nrm);
      if (dot(t, t) < 1e-12) { t = cross(mk3<CR>(1, 0, 0), nrm); if (dot(t, t) < 1e-12) { t = cross(mk3<CR>(0, 1, 0), nrm); if (dot(t, t) < 1e-12) t = cross(mk3<CR>(0, 0, 1), nrm); } }
      dirs[1] = t * (CR(1) / nb2_sqrt(dot(t, t)));
      dirs[2] = cross(nrm, dirs[1]);
    }
    const int dim = fric ? 3 : 1, off = m;
    const int ba = ws.cbodyA[c], bb = ws.cbodyB[c];
    for (int k = 0; k < dim; k++) {
      CR rel = 0;
      V6<CR> JA = zero6<CR>(), JB = zero6<CR>();
      if (ba >= 0) { const Xf<CR> W = xf_from12(ws.W + 12 * ba); const V3<CR> pA = xf_apply_inv(W, pt), dA = mulT(W.R_, dirs[k]); JA.a = cross(pA, dA); JA.l = dA; rel -= dot(JA, ldv6(ws.V + 6 * ba)); }
      if (bb >= 0) { const Xf<CR> W = xf_from12(ws.W + 12 * bb); const V3<CR> pB = xf_apply_inv(W, pt), dB = mulT(W.R_, -dirs[k]); JB.a = cross(pB, dB); JB.l = dB; rel -= dot(JB, ldv6(ws.V + 6 * bb)); }
      stv6(ws.JA + 6 * m, JA); stv6(ws.JB + 6 * m, JB);
      ws.b[m] = rel; ws.rest[m] = (k == 0 && bounce) ? e : 0.0;
      ws.i1[m] = c;  // row -> contact (i1 is free until classification)
      if (k == 0) { ws.lo[m] = 0.0; ws.hi[m] = HUGE_VAL; ws.findex[m] = -1; } else { ws.lo[m] = -mu; ws.hi[m] = mu; ws.findex[m] = off; }
      m++;
    }
    CR bv = ws.cdepth[c];
    if (bv < 0) bv = 0; else { bv *= 0.01 * (1.0 / dt); if (bv > 1e-3) bv = 1e-3; }
    if (!C.pen_correction) bv = 0;
    else if (bv > 0) status |= NB2_ST_BOUNCE;  // depth-dependent ERP velocity: not modelled by the backward either (loud)
    if (bounce) { const CR rv = ws.b[off] * e; if (rv > 1e-1) { if (rv > bv) { bv = rv; if (bv > 1e2) bv = 1e2; status |= NB2_ST_BOUNCE; } } }
    ws.b[off] += bv;
  }
  ws.meta[0] = m; ws.meta[1] = nc; ws.meta[2] = status;
  if (m == 0) { *m_io = 0; *status_out = status; if (crec) crec[0] = 0; return; }  // out already holds v*
  // row -> contact map must survive classification (which uses i1): copy to st8 region as bytes
  auto rowc = ws.st8 + NB2_MAX_ROWS;
  for (int r = 0; r < m; r++) rowc[r] = (unsigned char)ws.i1[r];
}

// phase 1 — A by impulse tests (upper blocks measured, lower mirrored in phase 2; BoxedLcpConstraintSolver.cpp:293-314).
// The m impulse tests are independent: lane `cl` of `nl` cooperating threads takes rows cl, cl + nl, ... with private
// sweep buffers; every lane writes only its own rows of A.
template <int ST>
NB2_HD void contact_phase1(const Nb2ModelDev<CR>& M, const CR* sv, size_t B, CR* wsblock, int lane, int cl, int nl) {
  const int nb = M.nb, n = M.ndof;
  const ContactWsT<ST> ws = carve_ws<ST>(wsblock, lane, nb, n);
  const int m = ws.meta[0], nc = ws.meta[1];
  if (m == 0) return;
  auto rowc = ws.st8 + NB2_MAX_ROWS;
  auto A = ws.A;
  unsigned long long mask = 0ull;  // ancestors (and self) of every contact body
  for (int c = 0; c < nc; c++) {
    for (int bdy = ws.cbodyA[c]; bdy >= 0; bdy = M.parent[bdy]) mask |= (1ull << bdy);
    for (int bdy = ws.cbodyB[c]; bdy >= 0; bdy = M.parent[bdy]) mask |= (1ull << bdy);
  }
  ContactWsT<ST> wl = ws;  // same world, this lane's sweep buffers
  {
    auto base = ws.lbuf + (cl % NB2_CONTACT_LANES) * (12 * nb + 2 * n);
    wl.pI = base; wl.V = base + 6 * nb; wl.uI = base + 12 * nb; wl.dqd = base + 12 * nb + n;
  }
  for (int r = cl; r < m; r += nl) {
    const int c = rowc[r];
    for (int i = 0; i < nb; i++) if ((mask >> i) & 1ull) for (int k = 0; k < 6; k++) wl.pI[6 * i + k] = 0;
    if (ws.cbodyA[c] >= 0) { auto J = ws.JA + 6 * r; auto p = wl.pI + 6 * ws.cbodyA[c]; for (int k = 0; k < 6; k++) p[k] -= J[k]; }
    if (ws.cbodyB[c] >= 0) { auto J = ws.JB + 6 * r; auto p = wl.pI + 6 * ws.cbodyB[c]; for (int k = 0; k < 6; k++) p[k] -= J[k]; }
    impulse_response<ST>(M, sv, B, wl, mask);
    for (int s2 = 0; s2 < m; s2++) {
      const int cj = rowc[s2];
      if (cj < c) continue;  // mirrored from row s2 in phase 2
      CR a = 0;
      if (ws.cbodyA[cj] >= 0) a += dot(ldv6(ws.JA + 6 * s2), ldv6(wl.V + 6 * ws.cbodyA[cj]));
      if (ws.cbodyB[cj] >= 0) a += dot(ldv6(ws.JB + 6 * s2), ldv6(wl.V + 6 * ws.cbodyB[cj]));
      A[r * m + s2] = a;
    }
  }
}

// phase 2 — mirror A, warm start, solve chain, classification, impulse application (one thread)
template <int ST>
NB2_HD void contact_phase2(const Nb2ModelDev<CR>& M, const Nb2ContactDev& C, const float* st, float* out, const CR* sv, size_t B,
                           CR* wsblock, int lane, CR* x_io, int* m_io, int* labels_out, int* status_out, CR* crec, const Grp& grp = Grp()) {
  const int nb = M.nb, n = M.ndof;
  ContactWsT<ST> ws = carve_ws<ST>(wsblock, lane, nb, n);
  ws.grp = grp;
  const int m = ws.meta[0];
  int status = ws.meta[2];
  if (m == 0) return;
  auto rowc = ws.st8 + NB2_MAX_ROWS;
  auto A = ws.A;
  for (int r = 0; r < m; r++) { const int cr = rowc[r]; for (int s2 = 0; s2 < r; s2++) if (rowc[s2] < cr) A[r * m + s2] = A[s2 * m + r]; }
  status |= lcp_chain_ws<ST>(m, ws, C.fallback_cfm, (*m_io == m) ? x_io : nullptr);
  auto x = ws.x;
  for (int i = 0; i < m; i++) { x_io[i] = x[i]; labels_out[i] = ws.mapping[i]; }
  // ---- apply the impulses and update the velocities (ContactConstraint.cpp:630-684, Skeleton.cpp:13571-13595)
  for (int i = 0; i < nb * 6; i++) ws.pI[i] = 0;
  for (int r = 0; r < m; r++) {
    const int c = rowc[r];
    if (ws.cbodyA[c] >= 0) { auto J = ws.JA + 6 * r; auto p = ws.pI + 6 * ws.cbodyA[c]; for (int k = 0; k < 6; k++) p[k] -= J[k] * x[r]; }
    if (ws.cbodyB[c] >= 0) { auto J = ws.JB + 6 * r; auto p = ws.pI + 6 * ws.cbodyB[c]; for (int k = 0; k < 6; k++) p[k] -= J[k] * x[r]; }
  }
  impulse_response<ST>(M, sv, B, ws, ~0ull);
  for (int d = 0; d < n; d++) out[n + d] = (float)(ws.vstar[d] + ws.dqd[d]);
  *m_io = m; *status_out = status;
  if (crec) {  // what the backward pass needs: sizes, labels, impulses, the velocity change they caused, the LCP matrix
    crec[0] = (CR)m; crec[1] = (CR)status;
    for (int i = 0; i < m; i++) { crec[2 + i] = (CR)ws.mapping[i]; crec[2 + NB2_MAX_ROWS + i] = x[i]; }
    CR* cd = crec + 2 + 2 * NB2_MAX_ROWS;
    for (int d = 0; d < n; d++) cd[d] = ws.dqd[d];
    CR* cA = cd + n;
    copy_n(cA, A, m * m);
  }
}


// single-thread form (host emulation, tests): the three phases back to back
template <int ST>
NB2_HD void world_contact(const Nb2ModelDev<CR>& M, const Nb2ContactDev& C, const float* st, float* out, const CR* sv, size_t B,
                          CR* wsblock, int lane, CR* x_io, int* m_io, int* labels_out, int* status_out, int* nc_out, float* cinfo_out,
                          CR* crec, int emulate_lanes = 1) {
  contact_phase0<ST>(M, C, st, out, sv, B, wsblock, lane, x_io, m_io, labels_out, status_out, nc_out, cinfo_out, crec);
  for (int cl = emulate_lanes - 1; cl >= 0; cl--) contact_phase1<ST>(M, sv, B, wsblock, lane, cl, emulate_lanes);
  contact_phase2<ST>(M, C, st, out, sv, B, wsblock, lane, x_io, m_io, labels_out, status_out, crec);
}

// =====================================================================================================
// backward through the contact stage (classification frozen at the forward solution), adjoint form.
// With  P = A_c + A_ub E,  f = Q^+ b_c,  v+ = v* + M^-1 P f  (BackpropSnapshot.cpp:980-1107 materialises the Jacobians of
// this map); for an incoming g = dL/dv+ :
//     lambda = M^-1 g ;  fbar = P^T lambda ;  mu = Q^-T fbar ;  nu = M^-1 A_c mu ;  w = lambda - nu
//     dL/dv* = g - A_c mu  (so the ABA part is back-propagated with w in place of lambda and the REALISED acceleration
//     (v+ - v)/dt in place of the unconstrained one) ;  dL/dtau = dt w
//     contact-Jacobian part:  d/dq of  Phi(q) = sum_r  f_r J_r(q) w  -  mu_r J_r(q) v+   at fixed w, v+  (upper-bound rows: f_r := x_r, mu_r := 0),
//     split into (i) the motion of the contact frame with the pose of the moving body — the contact generator re-run on
//     dual numbers for the 6 pose directions, any contact type — and (ii) the kinematic chain (reverse velocity recursion).
// This function runs between the lambda sweeps (B1/B2) and the reverse RNEA sweep (B3) of world_backward and prepares:
//     scr: lambda -> w, W_i -> W_i(w);   ws: per-body injections, realised accelerations, v+ fields.
// =====================================================================================================
template <int ST>
struct BwdContactView {  // views into the contact workspace used by world_backward<double, ST, true>
  SP<CR, ST> Aacc;   // [nb][6] spatial accelerations for the realised joint accelerations
  SP<CR, ST> Uplus;  // [nb][6] spatial velocities for v+
  SP<CR, ST> aeff;   // [n]
  SP<CR, ST> vplus;  // [n]
  SP<CR, ST> inj;    // [nb][24]  Uw_bar(6) Up_bar(6) G(6) H(6), already scaled by -1/dt except H
  SP<CR, ST> JcTmu;  // [n] out
  int active;       // 0: this world had no contact rows (plain contact-free backward)
  int error;        // structure mismatch / unsupported pair
};

template <int ST>
NB2_HD BwdContactView<ST> contact_backward_prepare(const Nb2ModelDev<CR>& M, const Nb2ContactDev& C, const float* st, const CR* sv, size_t B,
                                                   CR* wsblock, int lane, const CR* crec, CR* scr, int oLam, int oBody) {
  const int nb = M.nb, n = M.ndof;
  const ContactWsT<ST> ws = carve_ws<ST>(wsblock, lane, nb, n);
  BwdContactView<ST> cv;
  cv.Aacc = ws.pI; cv.Uplus = ws.V; cv.aeff = ws.uI; cv.vplus = ws.vstar; cv.inj = ws.Q2; cv.JcTmu = ws.dqd; cv.active = 0; cv.error = 0;
  const int m = (int)crec[0];
  if (m <= 0) return cv;
  cv.active = 1;
  // restitution: b depends on v* through (1 + e) J v*, which needs a second multiplier field in the reverse sweep
  // (BackpropSnapshot::getBounceApproximationJacobian, BackpropSnapshot.cpp:1131-1226) — not implemented: fail loudly
  if (((int)crec[1]) & NB2_ST_BOUNCE) { cv.error = 5; return cv; }
  const CR dt = M.dt;
  const int kQdd = nb * 21 + M.nfree * 33;
  const CR* mapping = crec + 2; const CR* xr = crec + 2 + NB2_MAX_ROWS; const CR* dqd_imp = crec + 2 + 2 * NB2_MAX_ROWS; const CR* Arec = dqd_imp + n;
  // ---- transforms
  for (int i = 0; i < nb; i++) {
    const int p = M.parent[i];
    const Xf<CR> T = saved_xf(M, i, st, sv, B);
    xf_to12(ws.T + 12 * i, T);
    xf_to12(ws.W + 12 * i, (p >= 0) ? xf_mul(xf_from12(ws.W + 12 * p), T) : T);
  }
  // ---- contact rows re-generated on dual numbers.  Row r acts on up to two bodies: wrench F_A on the body of shape A and
  // F_B on the body of shape B (ContactConstraint.cpp:66-230; a static side has no term).  The generator runs in two
  // passes: pass 0 records the wrenches (needed for mu), pass 1 — after mu, w and v+ are known — differentiates them with
  // respect to the pose of one moving body at a time (6 twist directions) and accumulates the contact-frame wrench G.
  typedef DualT<6> D6;
  auto rowFA = ws.JA; auto rowFB = ws.JB; auto rowbA = ws.i1; auto rowbB = ws.findex; auto rowmu = ws.v1;
  auto coefWr = ws.v7; auto coefVr = ws.v8;  // per-row weights of the w- and v+-fields (filled before pass 1)
  auto inj = ws.Q2;
  auto Aacc = ws.pI; auto Uplus = ws.V;      // valid in pass 1 (after the impulse sweep for nu released these buffers)
  auto lift = [](const Xf<CR>& X) { Xf<D6> o; const CR* r = &X.R_.m00; D6* q = &o.R_.m00; for (int i = 0; i < 9; i++) q[i] = D6(r[i]); o.p.x = D6(X.p.x); o.p.y = D6(X.p.y); o.p.z = D6(X.p.z); return o; };
  auto dual_pose = [&](int dyn) {  // world transform of body `dyn` perturbed by a body twist:  W (I + [xi_w]x , xi_v)
    const Xf<CR> Wd = xf_from12(ws.W + 12 * dyn);
    Xf<D6> WD;
    {
      const CR* r = &Wd.R_.m00; D6* o = &WD.R_.m00;
      for (int i = 0; i < 9; i++) o[i] = D6(r[i]);
      // dR/dxi_w[k] = R skew(e_k):  columns: R[:,a] x ... ; (R [e_k]x)[:,j] = R (e_k x e_j)
      // k=0: e0 x e1 = e2, e0 x e2 = -e1 ; k=1: e1 x e0 = -e2, e1 x e2 = e0 ; k=2: e2 x e0 = e1, e2 x e1 = -e0
      const V3<CR> c0 = col3(Wd.R_, 0), c1 = col3(Wd.R_, 1), c2 = col3(Wd.R_, 2);
      // column 1 gets +c2 for k=0 ; column 2 gets -c1 for k=0
      WD.R_.m01.d[0] = c2.x; WD.R_.m11.d[0] = c2.y; WD.R_.m21.d[0] = c2.z;
      WD.R_.m02.d[0] = -c1.x; WD.R_.m12.d[0] = -c1.y; WD.R_.m22.d[0] = -c1.z;
      WD.R_.m00.d[1] = -c2.x; WD.R_.m10.d[1] = -c2.y; WD.R_.m20.d[1] = -c2.z;
      WD.R_.m02.d[1] = c0.x; WD.R_.m12.d[1] = c0.y; WD.R_.m22.d[1] = c0.z;
      WD.R_.m00.d[2] = c1.x; WD.R_.m10.d[2] = c1.y; WD.R_.m20.d[2] = c1.z;
      WD.R_.m01.d[2] = -c0.x; WD.R_.m11.d[2] = -c0.y; WD.R_.m21.d[2] = -c0.z;
      WD.p.x = D6(Wd.p.x); WD.p.y = D6(Wd.p.y); WD.p.z = D6(Wd.p.z);
      WD.p.x.d[3] = c0.x; WD.p.y.d[3] = c0.y; WD.p.z.d[3] = c0.z;
      WD.p.x.d[4] = c1.x; WD.p.y.d[4] = c1.y; WD.p.z.d[4] = c1.z;
      WD.p.x.d[5] = c2.x; WD.p.y.d[5] = c2.y; WD.p.z.d[5] = c2.z;
    }
    return WD;
  };
  unsigned long long pairs_with_rows = 0ull;  // NB2_MAX_PAIRS <= 64
  // S = CR: values only (pass 0) ; S = D6: derivatives with respect to one moving body's pose at a time (pass 1)
  auto rows_pass = [&](auto tag) {
    typedef decltype(tag) S;
    constexpr bool DUALS = !std::is_same<S, CR>::value;
    constexpr int pass = DUALS ? 1 : 0;
    auto liftS = [](const Xf<CR>& X) { Xf<S> o; const CR* r = &X.R_.m00; S* q = &o.R_.m00; for (int i = 0; i < 9; i++) q[i] = S(r[i]); o.p.x = S(X.p.x); o.p.y = S(X.p.y); o.p.z = S(X.p.z); return o; };
    auto poseS = [&](int body, bool vary) {
      if constexpr (DUALS) { if (vary) return dual_pose(body); }
      return liftS(xf_from12(ws.W + 12 * body));
    };
    int m2 = 0;
    for (int pi = 0; pi < C.npairs && !cv.error; pi++) {
      if (DUALS && !((pairs_with_rows >> pi) & 1ull)) continue;  // the plain pass found no contact row for this pair
      const int sa = C.pair_a[pi], sb = C.pair_b[pi];
      const int ba = C.shape_body[sa], bb = C.shape_body[sb];
      const int nvary = (pass == 1 && ba >= 0 && bb >= 0) ? 2 : 1;
      const int m2_pair = m2;
      for (int v = 0; v < nvary; v++) {
        m2 = m2_pair;
        const bool varyA = (ba >= 0) && (v == 0);          // which body's pose carries the dual part in this run
        const int dyn = varyA ? ba : bb;
        const Xf<S> WDa = (ba >= 0) ? poseS(ba, varyA) : Xf<S>();
        const Xf<S> WDb = (bb >= 0) ? poseS(bb, !varyA) : Xf<S>();
        const Xf<S> Ta = (ba >= 0) ? gxf_mul(WDa, liftS(xf_from12(C.shape_T[sa]))) : liftS(xf_from12(C.shape_T[sa]));
        const Xf<S> Tb = (bb >= 0) ? gxf_mul(WDb, liftS(xf_from12(C.shape_T[sb]))) : liftS(xf_from12(C.shape_T[sb]));
        const int ta = C.shape_type[sa], tb = C.shape_type[sb];
        const V3<S> da = mk3<S>(S(C.shape_dims[sa][0]), S(C.shape_dims[sa][1]), S(C.shape_dims[sa][2]));
        const V3<S> db = mk3<S>(S(C.shape_dims[sb][0]), S(C.shape_dims[sb][1]), S(C.shape_dims[sb][2]));
        ContactOutT<S> co[8];
        int k = 0;
        if (ta == 0 && tb == 0) k = collide_box_box(da, Ta, db, Tb, C.clip_depth, co);
        else if (ta == 0 && tb == 1) k = collide_box_sphere(da, Ta, db.x, Tb, C.clip_depth, 0, false, co);
        else if (ta == 1 && tb == 0) k = collide_box_sphere(db, Tb, da.x, Ta, C.clip_depth, 0, true, co);
        else if ((ta == 0 && tb == 2) || (ta == 2 && tb == 0)) {
          const bool boxFirst = (ta == 0);
          const Xf<S>& Tc = boxFirst ? Tb : Ta; const Xf<S>& Tbx = boxFirst ? Ta : Tb;
          const V3<S> bdim = boxFirst ? da : db;
          const S r = boxFirst ? db.x : da.x; const CR h = boxFirst ? C.shape_dims[sb][1] : C.shape_dims[sa][1];
          CR dep[2]; Xf<S> Tend[2];
          for (int e = 0; e < 2; e++) {
            Tend[e] = Tc; Tend[e].p = gxf_apply(Tc, mk3<S>(S(0.0), S(0.0), S(e == 0 ? h / 2 : -h / 2)));
            const V3<S> pld = gxf_apply_inv(Tbx, Tend[e].p);
            const V3<CR> pl = mk3<CR>(gval(pld.x), gval(pld.y), gval(pld.z));
            V3<CR> q = pl; bool inside = true;
            for (int kk = 0; kk < 3; kk++) { const CR hk = 0.5 * gval(gget3(bdim, kk)), v = get3(q, kk); if (v < -hk) { set3(q, kk, -hk); inside = false; } if (v > hk) { set3(q, kk, hk); inside = false; } }
            if (inside) { CR mn = 1e300; for (int kk = 0; kk < 3; kk++) { const CR v = 0.5 * gval(gget3(bdim, kk)) - nb2_abs(get3(pl, kk)); mn = v < mn ? v : mn; } dep[e] = mn + gval(r); }
            else { const V3<CR> dd = pl - q; dep[e] = gval(r) - nb2_sqrt(dot(dd, dd)); }
          }
          if ((dep[0] > dep[1] ? dep[0] : dep[1]) >= 0 && nb2_abs(dep[0] - dep[1]) >= 1e-9) {
            const int e = dep[0] > dep[1] ? 0 : 1;
            k = collide_box_sphere(bdim, Tbx, r, Tend[e], C.clip_depth, e == 0 ? 1 : 2, !boxFirst, co);
          }
        }
        for (int c = 0; c < k; c++) {
          const V3<CR> nv = mk3<CR>(gval(co[c].normal.x), gval(co[c].normal.y), gval(co[c].normal.z));
          if (dot(nv, nv) < 1e-12) continue;
          if (gval(co[c].depth) < 0.0 || gval(co[c].depth) > C.clip_depth) continue;
          const CR mu = C.shape_mu[sa] < C.shape_mu[sb] ? C.shape_mu[sa] : C.shape_mu[sb];
          const bool fric = mu > 1e-3;
          V3<S> dirs[3]; dirs[0] = co[c].normal;
          if (fric) tangent_basis<S>(co[c].normal, &dirs[1], &dirs[2]);
          const int dim = fric ? 3 : 1;
          V3<S> pA, pB;
          if (ba >= 0) pA = gxf_apply_inv(WDa, co[c].point);
          if (bb >= 0) pB = gxf_apply_inv(WDb, co[c].point);
          for (int kk = 0; kk < dim; kk++) {
            if (m2 >= NB2_MAX_ROWS) { cv.error = 2; break; }
            S FA[6], FB[6];
            if (ba >= 0) { const V3<S> dA = mulT(WDa.R_, dirs[kk]); const V3<S> mo = cross(pA, dA); FA[0] = mo.x; FA[1] = mo.y; FA[2] = mo.z; FA[3] = dA.x; FA[4] = dA.y; FA[5] = dA.z; }
            if (bb >= 0) { const V3<S> dB = mulT(WDb.R_, -dirs[kk]); const V3<S> mo = cross(pB, dB); FB[0] = mo.x; FB[1] = mo.y; FB[2] = mo.z; FB[3] = dB.x; FB[4] = dB.y; FB[5] = dB.z; }
            if constexpr (!DUALS) {
              for (int j = 0; j < 6; j++) { rowFA[6 * m2 + j] = (ba >= 0) ? gval(FA[j]) : 0.0; rowFB[6 * m2 + j] = (bb >= 0) ? gval(FB[j]) : 0.0; }
              rowbA[m2] = ba; rowbB[m2] = bb; rowmu[m2] = mu;
              pairs_with_rows |= (1ull << pi);
            } else if (coefWr[m2] != 0.0 || coefVr[m2] != 0.0) {
              // G_dyn += sum_terms (dF_term/dxi_dyn)^T (coefW * field_w(body_term) + coefV * field_v+(body_term)), scaled by -1/dt
              auto gj = inj + 24 * dyn + 12;
              for (int side = 0; side < 2; side++) {
                const int body = side == 0 ? ba : bb;
                if (body < 0) continue;
                const S* F = side == 0 ? FA : FB;
                const V6<CR> Ww = ld6<CR, ST>(scr + (size_t)(oBody + 7 * body + 1) * ST);
                const V6<CR> Up = ldv6(Uplus + 6 * body);
                const CR fw[6] = {Ww.a.x, Ww.a.y, Ww.a.z, Ww.l.x, Ww.l.y, Ww.l.z}, fu[6] = {Up.a.x, Up.a.y, Up.a.z, Up.l.x, Up.l.y, Up.l.z};
                for (int kx = 0; kx < 6; kx++) {
                  CR g = 0;
                  for (int jx = 0; jx < 6; jx++) g += F[jx].d[kx] * (coefWr[m2] * fw[jx] + coefVr[m2] * fu[jx]);
                  gj[kx] += (-1.0 / dt) * g;
                }
              }
            }
            m2++;
          }
        }
      }
    }
    return m2;
  };
  const int m2 = rows_pass(CR());
  if (m2 != m) cv.error = cv.error ? cv.error : 3;
  if (cv.error) return cv;
  // ---- clamping / upper-bound sets from the saved labels
  auto clampIdx = ws.clampIdx; auto cl = ws.i2; auto ubl = ws.ubIdx;  // ubl: list of ub rows
  int nCl = 0, nUb = 0;
  for (int j = 0; j < m; j++) { clampIdx[j] = -1; if ((int)mapping[j] == NB2_MAP_CLAMPING) { clampIdx[j] = nCl; cl[nCl++] = j; } }
  for (int j = 0; j < m; j++) if ((int)mapping[j] >= 0) ubl[nUb++] = j;
  auto fbar = ws.v2; auto mu_c = ws.v3; auto Eu = ws.v4;
  // W_body(lambda) read from the strided scratch
  auto ldW = [&](int body) { return ld6<CR, ST>(scr + (size_t)(oBody + 7 * body + 1) * ST); };
  auto rowdot = [&](int j) {  // J_j lambda-field: wrench of row j against the field induced on its (one or two) bodies
    CR a = 0;
    if (rowbA[j] >= 0) a += dot(ldv6(rowFA + 6 * j), ldW(rowbA[j]));
    if (rowbB[j] >= 0) a += dot(ldv6(rowFB + 6 * j), ldW(rowbB[j]));
    return a;
  };
  for (int r = 0; r < nCl; r++) fbar[r] = rowdot(cl[r]);
  for (int u = 0; u < nUb; u++) {
    const int j = ubl[u], fp = (int)mapping[j];
    const CR up = xr[fp] * rowmu[j], low = -xr[fp] * rowmu[j];
    Eu[u] = (nb2_abs(xr[j] - up) < nb2_abs(xr[j] - low)) ? rowmu[j] : -rowmu[j];
    fbar[clampIdx[fp]] += Eu[u] * rowdot(j);
  }
  for (int r = 0; r < nCl; r++) mu_c[r] = 0;
  if (nCl > 0) {
    auto Q = ws.Q;
    for (int r = 0; r < nCl; r++) {
      const int rowr = cl[r] * m;
      for (int c0 = 0; c0 < nCl; c0 += 4) {  // batched: the gathers of four entries are issued before their stores
        CR t[4];
#pragma unroll
        for (int u = 0; u < 4; u++) if (c0 + u < nCl) t[u] = Arec[rowr + cl[c0 + u]];
#pragma unroll
        for (int u = 0; u < 4; u++) if (c0 + u < nCl) Q[r * nCl + c0 + u] = t[u];
      }
    }
    for (int u = 0; u < nUb; u++) { const int j = ubl[u], c = clampIdx[(int)mapping[j]]; for (int r = 0; r < nCl; r++) Q[r * nCl + c] += Arec[cl[r] * m + j] * Eu[u]; }
    if (nUb == 0) pinv_psd(nCl, Q, fbar, mu_c, ws.Aw, ws.L, ws.v5, ws.v6, ws.mapping);
    else {  // Q^T mu = fbar  ->  mu = (Q Q^T)^+ Q fbar
      auto QQt = ws.Aw + (size_t)NB2_MAX_ROWS * NB2_MAX_ROWS / 2;  // nCl <= 33 guaranteed below
      if (2 * nCl * nCl > NB2_MAX_ROWS * NB2_MAX_ROWS / 2 * 2) { cv.error = 4; return cv; }
      auto Qf = ws.v7;
      for (int a = 0; a < nCl; a++) {
        CR sacc = 0; for (int c = 0; c < nCl; c++) sacc += Q[a * nCl + c] * fbar[c];
        Qf[a] = sacc;
        for (int c = 0; c < nCl; c++) {
          CR t = 0;
#pragma unroll 4
          for (int kx = 0; kx < nCl; kx++) t += Q[a * nCl + kx] * Q[c * nCl + kx];
          QQt[a * nCl + c] = t;
        }
      }
      pinv_psd(nCl, QQt, Qf, mu_c, ws.Aw, ws.L, ws.v5, ws.v6, ws.mapping);
    }
  }
  // ---- nu = M^-1 A_c mu  (one impulse sweep) ; w = lambda - nu ; W_i(w)
  for (int i = 0; i < nb * 6; i++) ws.pI[i] = 0;
  for (int r = 0; r < nCl; r++) {
    const int j = cl[r];
    if (rowbA[j] >= 0) { auto F = rowFA + 6 * j; auto p = ws.pI + 6 * rowbA[j]; for (int kx = 0; kx < 6; kx++) p[kx] -= F[kx] * mu_c[r]; }
    if (rowbB[j] >= 0) { auto F = rowFB + 6 * j; auto p = ws.pI + 6 * rowbB[j]; for (int kx = 0; kx < 6; kx++) p[kx] -= F[kx] * mu_c[r]; }
  }
  impulse_response<ST>(M, sv, B, ws, ~0ull);
  for (int d = 0; d < n; d++) scr[(size_t)(oLam + d) * ST] -= ws.dqd[d];
  for (int i = 0; i < nb; i++) {
    const V6<CR> Wn = ld6<CR, ST>(scr + (size_t)(oBody + 7 * i + 1) * ST) - ldv6(ws.V + 6 * i);
    st6<CR, ST>(scr + (size_t)(oBody + 7 * i + 1) * ST, Wn);
  }
  // ---- realised accelerations, v+, and the fields they induce (the sweep buffers pI / V are free now: Aacc | Uplus;
  // ws.W keeps the world transforms for the second generator pass)
  auto aeff = ws.uI; auto vplus = ws.vstar;
  for (int d = 0; d < n; d++) { aeff[d] = sv[(size_t)(kQdd + d) * B] + dqd_imp[d] / dt; vplus[d] = (CR)st[n + d] + dt * aeff[d]; }
  V6<CR> A0; A0.a = zero3<CR>(); A0.l = mk3<CR>(-M.gravity[0], -M.gravity[1], -M.gravity[2]);
  for (int i = 0; i < nb; i++) {
    const int jt = M.jtype[i], p = M.parent[i], o = M.dof_off[i];
    const Xf<CR> T = xf_from12(ws.T + 12 * i);
    const V6<CR> V = sv_ld6<CR>(sv + (size_t)(i * 21) * B, B, 0);
    V6<CR> Ai = AdInvT(T, (p >= 0) ? ldv6(Aacc + 6 * p) : A0);
    V6<CR> Ui = (p >= 0) ? AdInvT(T, ldv6(Uplus + 6 * p)) : zero6<CR>();
    if (jt != NB2_JT_FREE) {
      const V6<CR> Sv = S_times<CR>(jt, (CR)st[n + o]);
      Ai = Ai + S_times<CR>(jt, aeff[o]) + ad(V, Sv);
      Ui = Ui + S_times<CR>(jt, vplus[o]);
    } else {
      V6<CR> Sv; Sv.a = mk3<CR>((CR)st[n + o], (CR)st[n + o + 1], (CR)st[n + o + 2]); Sv.l = mk3<CR>((CR)st[n + o + 3], (CR)st[n + o + 4], (CR)st[n + o + 5]);
      Ai = Ai + ldv6(aeff + o) + ad(V, Sv);
      Ui = Ui + ldv6(vplus + o);
    }
    stv6(Aacc + 6 * i, Ai); stv6(Uplus + 6 * i, Ui);
  }
  // ---- per-body injections for the reverse sweep: Uw_bar, Up_bar, G (all scaled by -1/dt: they join the (dID/dq)^T w
  // accumulator that is multiplied by -dt at the end) and H (plain: A_c mu propagated to joint space)
  for (int i = 0; i < nb * 24; i++) inj[i] = 0;
  const CR kap = -1.0 / dt;
  for (int j = 0; j < m; j++) {
    CR coefW = 0, coefV = 0, coefH = 0;
    if (clampIdx[j] >= 0) { coefW = xr[j]; coefV = -mu_c[clampIdx[j]]; coefH = mu_c[clampIdx[j]]; }
    else if ((int)mapping[j] >= 0) coefW = xr[j];
    coefWr[j] = coefW; coefVr[j] = coefV;
    if (coefW == 0.0 && coefV == 0.0 && coefH == 0.0) continue;
    for (int side = 0; side < 2; side++) {
      const int body = side == 0 ? rowbA[j] : rowbB[j];
      if (body < 0) continue;
      auto F = (side == 0 ? rowFA : rowFB) + 6 * j;
      auto bj = inj + 24 * body;
      for (int kx = 0; kx < 6; kx++) {
        bj[kx] += kap * coefW * F[kx];
        bj[6 + kx] += kap * coefV * F[kx];
        bj[18 + kx] += coefH * F[kx];
      }
    }
  }
  rows_pass(D6());  // contact-frame part G: derivatives of every wrench w.r.t. the pose of each moving body
  return cv;
}

template <int ST>
NB2_HD BwdContactData<ST> contact_backward_hook(const Nb2ModelDev<double>& M, const BwdContactHook& H, const float* st, const double* sv, size_t B,
                                            double* scr, int oLam, int oBody) {
  const BwdContactView<ST> v = contact_backward_prepare<ST>(M, *(const Nb2ContactDev*)H.model_contact, st, sv, B, H.ws, H.lane, H.crec, scr, oLam, oBody);
  BwdContactData<ST> d;
  d.Aacc.p = v.Aacc.p; d.Uplus.p = v.Uplus.p; d.aeff.p = v.aeff.p; d.vplus.p = v.vplus.p; d.inj.p = v.inj.p; d.JcTmu.p = v.JcTmu.p; d.active = v.active; d.error = v.error;
  return d;
}

}  // namespace nb2
