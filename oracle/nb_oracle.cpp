// =====================================================================================
// TEST INFRASTRUCTURE ONLY.  fp64 CPU oracle for the differentiable timestep.
//
// A plain-C++ (no Eigen) restatement of the reference's algorithm for one world:
//   World::step                      dart/simulation/World.cpp:221-254, :307-333
//   Skeleton::computeForwardDynamics dart/dynamics/Skeleton.cpp:13296-13314 (ABA)
//   BodyNode::update{ArtInertia,BiasForce,AccelerationFD}  dart/dynamics/BodyNode.cpp:2046-2185
//   GenericJoint::*Dynamic           dart/dynamics/detail/GenericJoint.hpp:2168-2185, 2276-2301,
//                                    2395-2421, 2554-2571, 2656-2676
//   joint kinematics                 RevoluteJoint.cpp:141-152,203-211; PrismaticJoint.cpp:139-145,188-196;
//                                    FreeJoint.cpp:65-81,922-929,1027-1061 (DART_USE_IDENTITY_JACOBIAN build)
//   BackpropSnapshot::backprop       dart/neural/BackpropSnapshot.cpp:121-194, clip :425-479, action map :404-417
//
// The five step Jacobians (posPos, velPos, posVel, velVel, forceVel; BackpropSnapshot.cpp:159-178) are
// obtained here by forward-mode automatic differentiation (dual numbers) of the restated forward step,
// i.e. they are the exact derivatives the reference's analytic formulas compute (the reference verifies its
// formulas against finite differences of the same step at 1e-8, unittests/GradientTestUtils.hpp:637-680).
//
// Parity status: the reference cannot be built in this container (Eigen/ccd/assimp absent), and its test-suite
// holds no golden numeric vectors for this path => value-level parity is UNPINNED; the oracle is pinned by the
// reference's own property tests instead (analytic-vs-FD consistency, tests/test_oracle.py) and, for the LCP
// stage, bit-comparison with the reference's ODE dSolveLCP compiled from /root/reference (oracle/Makefile).
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this library.
// =====================================================================================
#include <cstdint>
#include <cstdio>
#include <vector>

#include "spatial.hpp"

namespace orc {

enum { WELD = 0, REVOLUTE = 1, PRISMATIC = 2, FREE = 3 };

struct Model {
  int nb = 0, ndof = 0;
  std::vector<int> parent, jtype, dof_off, mobile;
  std::vector<double> axis, Tpj, Tcj, mass, com, moment;
  std::vector<double> damping, spring, rest, pos_lo, pos_hi, vel_lo, vel_hi, force_lo, force_hi;
  double gravity[3] = {0, 0, -9.81};
  double dt = 1e-3;
  std::vector<int> action_map;
};

template <class S> static Iso<S> iso_from12(const double* t) {
  Iso<S> T;
  for (int i = 0; i < 9; i++) T.R.m[i] = S(t[i]);
  for (int i = 0; i < 3; i++) T.p[i] = S(t[9 + i]);
  return T;
}

// spatial inertia tensor (Inertia::computeSpatialTensor, dart/dynamics/Inertia.cpp:1368-1383)
template <class S> static Mat6<S> spatial_tensor(const Model& M, int i) {
  const double* mo = &M.moment[6 * i];
  double m = M.mass[i];
  Vec3<double> c = v3<double>(M.com[3 * i], M.com[3 * i + 1], M.com[3 * i + 2]);
  Mat3<double> C = skew(c), CT = transpose(C), CCt = mul(C, CT);
  double I[9] = {mo[0], mo[3], mo[4], mo[3], mo[1], mo[5], mo[4], mo[5], mo[2]};
  Mat6<S> G = zero66<S>();
  for (int r = 0; r < 3; r++) for (int cc = 0; cc < 3; cc++) {
    G(r, cc) = S(I[3 * r + cc] + m * CCt(r, cc));
    G(3 + r, cc) = S(m * CT(r, cc));
    G(r, 3 + cc) = S(m * C(r, cc));
    G(3 + r, 3 + cc) = S(r == cc ? m : 0.0);
  }
  return G;
}

template <class S> struct BodyState {
  Iso<S> T;       // parent body <- this body  (Joint::getRelativeTransform)
  Iso<S> W;       // world <- this body
  Mat6<S> G;      // spatial inertia
  Vec6<S> Scol[6];  // joint Jacobian columns (body frame)
  int k = 0;      // dofs
  Vec6<S> V, eta; // spatial velocity, partial acceleration
  Mat6<S> AI;     // articulated inertia
  Vec6<S> pA;     // bias force
  S psi[36];      // inverse projected articulated inertia (k x k)
  S u[6];         // total force
  Vec6<S> A;      // spatial acceleration
};

template <class S> static void invert_spd(const S* Min, int k, S* out) {
  // Gauss-Jordan with partial pivoting on a small k x k matrix (k in {1,6});
  // the reference uses math::inverse<ConfigSpaceT> (1/x for R1, LDLT-based inverse for SE3)
  S a[36], b[36];
  for (int i = 0; i < k * k; i++) a[i] = Min[i];
  for (int i = 0; i < k; i++) for (int j = 0; j < k; j++) b[i * k + j] = S(i == j ? 1.0 : 0.0);
  for (int c = 0; c < k; c++) {
    int piv = c; double best = std::fabs(val(a[c * k + c]));
    for (int r = c + 1; r < k; r++) if (std::fabs(val(a[r * k + c])) > best) { best = std::fabs(val(a[r * k + c])); piv = r; }
    if (piv != c) for (int j = 0; j < k; j++) { S t = a[c * k + j]; a[c * k + j] = a[piv * k + j]; a[piv * k + j] = t; t = b[c * k + j]; b[c * k + j] = b[piv * k + j]; b[piv * k + j] = t; }
    S inv = S(1.0) / a[c * k + c];
    for (int j = 0; j < k; j++) { a[c * k + j] = a[c * k + j] * inv; b[c * k + j] = b[c * k + j] * inv; }
    for (int r = 0; r < k; r++) if (r != c) {
      S f = a[r * k + c];
      for (int j = 0; j < k; j++) { a[r * k + j] = a[r * k + j] - f * a[c * k + j]; b[r * k + j] = b[r * k + j] - f * b[c * k + j]; }
    }
  }
  for (int i = 0; i < k * k; i++) out[i] = b[i];
}

// FreeJoint::integratePositionsExplicit, identity-Jacobian branch (FreeJoint.cpp:922-929)
template <class S> static void free_integrate(const S* q, const S* v, double dt, S* out) {
  Mat3<S> R = expMapRot(v3<S>(q[0], q[1], q[2]));
  Mat3<S> E = expMapRot(v3<S>(v[0] * dt, v[1] * dt, v[2] * dt));
  Vec3<S> w = logMap(mul(R, E));
  Vec3<S> p = v3<S>(q[3], q[4], q[5]) + mul(R, v3<S>(v[3] * dt, v[4] * dt, v[5] * dt));
  for (int i = 0; i < 3; i++) { out[i] = w[i]; out[3 + i] = p[i]; }
}

// One contact-free World::step.  q,v,tau: [ndof] -> qn, vn.  (World.cpp:221-254,307-333)
template <class S>
static void step_nocontact(const Model& M, const S* q, const S* v, const S* tau, S* qn, S* vn, S* qdd_out = nullptr) {
  const int nb = M.nb;
  // per-thread reusable workspace: a fresh ~1 MB vector per call would hit mmap/munmap and serialise threads
  static thread_local std::vector<BodyState<S>> B;
  if ((int)B.size() < nb) B.resize(nb);
  const double dt = M.dt;
  // ---- kinematics, root -> leaf (Frame.cpp:144-160, GenericJoint.hpp:1803-1823)
  for (int i = 0; i < nb; i++) {
    BodyState<S>& b = B[i];
    const int o = M.dof_off[i];
    Iso<S> Tpj = iso_from12<S>(&M.Tpj[12 * i]), Tcj = iso_from12<S>(&M.Tcj[12 * i]);
    Vec3<S> ax = v3<S>(S(M.axis[3 * i]), S(M.axis[3 * i + 1]), S(M.axis[3 * i + 2]));
    Iso<S> Q = iso_identity<S>();
    b.k = 0;
    switch (M.jtype[i]) {
      case REVOLUTE: {  // RevoluteJoint.cpp:203-211 (T), :141-152 (S = AdTAngular(T_cj, axis))
        b.k = 1;
        Q.R = expMapRot(ax * q[o]);
        b.Scol[0] = AdT(Tcj, v6(ax, v3<S>(S(0.0), S(0.0), S(0.0))));
        break;
      }
      case PRISMATIC: {  // PrismaticJoint.cpp:188-196, :139-145
        b.k = 1;
        Q.p = ax * q[o];
        b.Scol[0] = AdT(Tcj, v6(v3<S>(S(0.0), S(0.0), S(0.0)), ax));
        break;
      }
      case FREE: {  // FreeJoint.cpp:74-81, 1027-1061: T = Tpj * [expMapRot(q0..2), q3..5] * Tcj^-1, S = Ad[Tcj]
        b.k = 6;
        Q.R = expMapRot(v3<S>(q[o], q[o + 1], q[o + 2]));
        Q.p = v3<S>(q[o + 3], q[o + 4], q[o + 5]);
        for (int c = 0; c < 6; c++) { Vec6<S> e = zero6<S>(); e[c] = S(1.0); b.Scol[c] = AdT(Tcj, e); }
        break;
      }
      default: break;  // WELD
    }
    b.T = mul(mul(Tpj, Q), inverse(Tcj));
    const int p = M.parent[i];
    b.W = (p >= 0) ? mul(B[p].W, b.T) : b.T;
    Vec6<S> Sv = zero6<S>();
    for (int c = 0; c < b.k; c++) Sv = Sv + b.Scol[c] * v[o + c];
    b.V = (p >= 0) ? AdInvT(b.T, B[p].V) + Sv : Sv;
    b.eta = ad(b.V, Sv);  // dS = 0 for these joint types in this build
    b.G = spatial_tensor<S>(M, i);
  }
  // ---- articulated inertia + bias force, leaf -> root (BodyNode.cpp:2046-2114)
  static thread_local std::vector<std::vector<int>> kids;
  if ((int)kids.size() < nb) kids.resize(nb);
  for (int i = 0; i < nb; i++) kids[i].clear();
  for (int i = 0; i < nb; i++) if (M.parent[i] >= 0) kids[M.parent[i]].push_back(i);
  Vec3<S> g = v3<S>(S(M.gravity[0]), S(M.gravity[1]), S(M.gravity[2]));
  for (int i = nb - 1; i >= 0; i--) {
    BodyState<S>& b = B[i];
    if (!M.mobile[i]) continue;
    b.AI = b.G;
    for (int c : kids[i]) {
      BodyState<S>& ch = B[c];
      Mat6<S> PI = ch.AI;
      if (ch.k > 0) {  // GenericJoint.hpp:2168-2185 ; weld child: no projection (ZeroDofJoint.cpp:828-835)
        Vec6<S> AIS[6];
        for (int a = 0; a < ch.k; a++) AIS[a] = mul(ch.AI, ch.Scol[a]);
        for (int r = 0; r < 6; r++) for (int cc = 0; cc < 6; cc++) {
          S acc = S(0.0);
          for (int a = 0; a < ch.k; a++) for (int e = 0; e < ch.k; e++) acc = acc + AIS[a][r] * ch.psi[a * ch.k + e] * AIS[e][cc];
          PI(r, cc) = PI(r, cc) - acc;
        }
      }
      b.AI = b.AI + transformInertia(inverse(ch.T), PI);
    }
    if (b.k > 0) {  // updateInvProjArtInertiaDynamic (GenericJoint.hpp:2276-2301)
      S proj[36];
      for (int a = 0; a < b.k; a++) { Vec6<S> AIa = mul(b.AI, b.Scol[a]); for (int e = 0; e < b.k; e++) proj[e * b.k + a] = dot(b.Scol[e], AIa); }
      invert_spd(proj, b.k, b.psi);
    }
    // bias force (BodyNode.cpp:2076-2104): -dad(V, G V) - Fext - G * AdInvRLinear(W, g)
    Vec6<S> Fg = mul(b.G, v6(v3<S>(S(0.0), S(0.0), S(0.0)), mulT(b.W.R, g)));
    b.pA = zero6<S>() - dad(b.V, mul(b.G, b.V)) - Fg;
    for (int c : kids[i]) {
      BodyState<S>& ch = B[c];
      Vec6<S> inner = ch.eta;  // addChildBiasForceToDynamic (GenericJoint.hpp:2395-2421)
      if (ch.k > 0) {
        for (int a = 0; a < ch.k; a++) { S s = S(0.0); for (int e = 0; e < ch.k; e++) s = s + ch.psi[a * ch.k + e] * ch.u[e]; inner = inner + ch.Scol[a] * s; }
      }
      Vec6<S> beta = ch.pA + mul(ch.AI, inner);
      b.pA = b.pA + dAdInvT(ch.T, beta);
    }
    // total force (GenericJoint.hpp:2554-2571): tau - K (q - q0 + v dt) - D v - S^T (AI eta + pA)
    Vec6<S> bodyForce = mul(b.AI, b.eta) + b.pA;
    const int o = M.dof_off[i];
    for (int a = 0; a < b.k; a++) {
      S spring = S(-M.spring[o + a]) * (q[o + a] - S(M.rest[o + a]) + v[o + a] * dt);
      S damp = S(-M.damping[o + a]) * v[o + a];
      b.u[a] = tau[o + a] + spring + damp - dot(b.Scol[a], bodyForce);
    }
  }
  // ---- accelerations, root -> leaf (BodyNode.cpp:2159-2185, GenericJoint.hpp:2656-2676, Frame.cpp:254-271)
  std::vector<S> qdd(M.ndof, S(0.0));
  for (int i = 0; i < nb; i++) {
    BodyState<S>& b = B[i];
    if (!M.mobile[i]) { b.A = zero6<S>(); continue; }
    const int p = M.parent[i], o = M.dof_off[i];
    Vec6<S> Ap = (p >= 0) ? AdInvT(b.T, B[p].A) : zero6<S>();
    Vec6<S> AIAp = mul(b.AI, Ap);
    Vec6<S> Sa = zero6<S>();
    for (int a = 0; a < b.k; a++) {
      S acc = S(0.0);
      for (int e = 0; e < b.k; e++) acc = acc + b.psi[a * b.k + e] * (b.u[e] - dot(b.Scol[e], AIAp));
      qdd[o + a] = acc;
      Sa = Sa + b.Scol[a] * acc;
    }
    b.A = Ap + Sa + b.eta;
  }
  // ---- integrate: v+ = v + dt qdd (GenericJoint.hpp:1410-1414); q+ uses the PRE-step velocity (World.cpp:307-322)
  for (int i = 0; i < nb; i++) {
    const int o = M.dof_off[i];
    const int k = (M.jtype[i] == FREE) ? 6 : (M.jtype[i] == WELD ? 0 : 1);
    if (M.jtype[i] == FREE) free_integrate(&q[o], &v[o], dt, &qn[o]);
    else for (int a = 0; a < k; a++) qn[o + a] = q[o + a] + v[o + a] * dt;
    for (int a = 0; a < k; a++) vn[o + a] = M.mobile[i] ? v[o + a] + qdd[o + a] * dt : v[o + a];
  }
  if (qdd_out) for (int i = 0; i < M.ndof; i++) qdd_out[i] = qdd[i];
}

// J = d[qn; vn] / d[q; v; tau]  row-major [2n x 3n]
static void step_jacobian(const Model& M, const double* q, const double* v, const double* tau, double* J) {
  constexpr int N = 12;
  typedef Dual<N> D;
  const int n = M.ndof, cols = 3 * n;
  std::vector<D> dq(n), dv(n), dtau(n), qn(n), vn(n);
  for (int c0 = 0; c0 < cols; c0 += N) {
    for (int i = 0; i < n; i++) { dq[i] = D(q[i]); dv[i] = D(v[i]); dtau[i] = D(tau[i]); }
    for (int k = 0; k < N && c0 + k < cols; k++) {
      int c = c0 + k;
      if (c < n) dq[c].d[k] = 1.0; else if (c < 2 * n) dv[c - n].d[k] = 1.0; else dtau[c - 2 * n].d[k] = 1.0;
    }
    step_nocontact<D>(M, dq.data(), dv.data(), dtau.data(), qn.data(), vn.data());
    for (int k = 0; k < N && c0 + k < cols; k++) {
      int c = c0 + k;
      for (int r = 0; r < n; r++) { J[(size_t)r * cols + c] = qn[r].d[k]; J[(size_t)(n + r) * cols + c] = vn[r].d[k]; }
    }
  }
}

}  // namespace orc

// =====================================================================================
// C interface (ctypes)
// =====================================================================================
using orc::Model;
extern "C" {

void* orc_model_create(int nb, int ndof, const int* parent, const int* jtype, const int* dof_off, const int* mobile,
                       const double* axis, const double* Tpj, const double* Tcj, const double* mass, const double* com,
                       const double* moment, const double* damping, const double* spring, const double* rest,
                       const double* pos_lo, const double* pos_hi, const double* vel_lo, const double* vel_hi,
                       const double* force_lo, const double* force_hi, const double* gravity, double dt, int na,
                       const int* action_map) {
  Model* M = new Model();
  M->nb = nb; M->ndof = ndof;
  M->parent.assign(parent, parent + nb); M->jtype.assign(jtype, jtype + nb);
  M->dof_off.assign(dof_off, dof_off + nb); M->mobile.assign(mobile, mobile + nb);
  M->axis.assign(axis, axis + 3 * nb); M->Tpj.assign(Tpj, Tpj + 12 * nb); M->Tcj.assign(Tcj, Tcj + 12 * nb);
  M->mass.assign(mass, mass + nb); M->com.assign(com, com + 3 * nb); M->moment.assign(moment, moment + 6 * nb);
  M->damping.assign(damping, damping + ndof); M->spring.assign(spring, spring + ndof); M->rest.assign(rest, rest + ndof);
  M->pos_lo.assign(pos_lo, pos_lo + ndof); M->pos_hi.assign(pos_hi, pos_hi + ndof);
  M->vel_lo.assign(vel_lo, vel_lo + ndof); M->vel_hi.assign(vel_hi, vel_hi + ndof);
  M->force_lo.assign(force_lo, force_lo + ndof); M->force_hi.assign(force_hi, force_hi + ndof);
  for (int i = 0; i < 3; i++) M->gravity[i] = gravity[i];
  M->dt = dt;
  M->action_map.assign(action_map, action_map + na);
  return M;
}

void orc_model_destroy(void* h) { delete (Model*)h; }

// state = [q; v] (2n), action (na) -> next_state (2n).   World::setState/setAction/step/getState
void orc_step(void* h, const double* state, const double* action, double* next_state, double* qdd /*nullable*/) {
  const Model& M = *(Model*)h;
  const int n = M.ndof;
  std::vector<double> tau(n, 0.0);
  for (size_t i = 0; i < M.action_map.size(); i++) tau[M.action_map[i]] = action[i];  // World.cpp:2061-2086
  orc::step_nocontact<double>(M, state, state + n, tau.data(), next_state, next_state + n, qdd);
}

// float-precision run of the same restatement (used only to *measure* fp32 sensitivity in tests)
void orc_step_f32(void* h, const double* state, const double* action, double* next_state) {
  const Model& M = *(Model*)h;
  const int n = M.ndof;
  std::vector<float> q(n), v(n), tau(n, 0.f), qn(n), vn(n);
  for (int i = 0; i < n; i++) { q[i] = (float)state[i]; v[i] = (float)state[n + i]; }
  for (size_t i = 0; i < M.action_map.size(); i++) tau[M.action_map[i]] = (float)action[i];
  orc::step_nocontact<float>(M, q.data(), v.data(), tau.data(), qn.data(), vn.data());
  for (int i = 0; i < n; i++) { next_state[i] = qn[i]; next_state[n + i] = vn[i]; }
}

// Jacobians wrt full tau (n columns), row-major: J[2n x 3n] = d[q+;v+]/d[q;v;tau]
void orc_jacobian(void* h, const double* state, const double* action, double* J) {
  const Model& M = *(Model*)h;
  const int n = M.ndof;
  std::vector<double> tau(n, 0.0);
  for (size_t i = 0; i < M.action_map.size(); i++) tau[M.action_map[i]] = action[i];
  orc::step_jacobian(M, state, state + n, tau.data(), J);
}

// BackpropSnapshot::backpropState: grad_next_state (2n) -> grad_state (2n), grad_action (na)
void orc_backprop(void* h, const double* state, const double* action, const double* grad_next, double* grad_state,
                  double* grad_action) {
  const Model& M = *(Model*)h;
  const int n = M.ndof, cols = 3 * n;
  std::vector<double> tau(n, 0.0);
  for (size_t i = 0; i < M.action_map.size(); i++) tau[M.action_map[i]] = action[i];
  std::vector<double> J((size_t)2 * n * cols);
  orc::step_jacobian(M, state, state + n, tau.data(), J.data());
  std::vector<double> g(cols, 0.0);
  for (int r = 0; r < 2 * n; r++) for (int c = 0; c < cols; c++) g[c] += J[(size_t)r * cols + c] * grad_next[r];
  // clipLossGradientsToBounds (BackpropSnapshot.cpp:425-479): exact-equality tests against the pre-step state
  for (int j = 0; j < n; j++) {
    double qj = state[j], vj = state[n + j], fj = tau[j];
    if (qj == M.pos_lo[j] && g[j] > 0) g[j] = 0;
    if (qj == M.pos_hi[j] && g[j] < 0) g[j] = 0;
    if (vj == M.vel_lo[j] && g[n + j] > 0) g[n + j] = 0;
    if (vj == M.vel_hi[j] && g[n + j] < 0) g[n + j] = 0;
    if (fj == M.force_lo[j] && g[2 * n + j] > 0) g[2 * n + j] = 0;
    if (fj == M.force_hi[j] && g[2 * n + j] < 0) g[2 * n + j] = 0;
  }
  for (int j = 0; j < 2 * n; j++) grad_state[j] = g[j];
  for (size_t i = 0; i < M.action_map.size(); i++) grad_action[i] = g[2 * n + M.action_map[i]];  // :404-417
}

}  // extern "C"
