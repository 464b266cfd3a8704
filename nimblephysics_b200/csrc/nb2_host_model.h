// nb2_model_desc (C ABI, doubles) -> Nb2ModelDev<R> (kernel parameter block).  Host only.
#pragma once
#include <string>

#include "../../include/nb2.h"
#include "nb2_model.h"
#include "nb2_cw.cuh"

template <class R>
static inline bool nb2_fill_model(const nb2_model_desc& d, Nb2ModelDev<R>& M, std::string& err) {
  if (d.nb <= 0 || d.nb > NB2_MAX_BODIES) { err = "model has " + std::to_string(d.nb) + " moving bodies; compiled limit is " + std::to_string(NB2_MAX_BODIES); return false; }
  if (d.ndof <= 0 || d.ndof > NB2_MAX_DOFS) { err = "model has " + std::to_string(d.ndof) + " dofs; compiled limit is " + std::to_string(NB2_MAX_DOFS); return false; }
  if (d.na < 0 || d.na > d.ndof) { err = "bad action map size"; return false; }
  M.nb = d.nb; M.ndof = d.ndof; M.na = d.na; M.nslots = d.nslots;
  M.pad2 = 0; M.pad3 = 0;
  auto magic = [](int x) { return x > 0 ? (unsigned)((0x100000000ull + (unsigned)x - 1) / (unsigned)x) : 0u; };
  M.magic_n2 = magic(2 * d.ndof); M.magic_n = magic(d.ndof); M.magic_na = magic(d.na);
  M.dt = (R)d.dt;
  for (int k = 0; k < 3; k++) M.gravity[k] = (R)d.gravity[k];
  int nfree = 0, ndof = 0;
  for (int i = 0; i < NB2_MAX_BODIES; i++) {
    M.parent[i] = -1; M.jtype[i] = 0; M.dof_off[i] = 0; M.flags[i] = 0; M.slot_self[i] = -1; M.slot_parent[i] = -1; M.slot_count[i] = 0; M.free_idx[i] = -1;
    for (int k = 0; k < 12; k++) M.Xtree[i][k] = R(0);
    for (int k = 0; k < 10; k++) M.inertia[i][k] = R(0);
  }
  for (int i = 0; i < d.nb; i++) {
    const int jt = d.jtype[i];
    if (jt != NB2_JT_REV && jt != NB2_JT_PRIS && jt != NB2_JT_FREE) { err = "unsupported canonical joint type"; return false; }
    if (d.parent[i] >= i) { err = "bodies must be numbered parents-first"; return false; }
    if (d.dof_off[i] != ndof) { /* canonical order may permute bodies relative to dof order: allowed */ }
    M.parent[i] = (int16_t)d.parent[i]; M.jtype[i] = (int16_t)jt; M.dof_off[i] = (int16_t)d.dof_off[i];
    M.flags[i] = (int16_t)d.flags[i]; M.slot_self[i] = (int16_t)d.slot_self[i]; M.slot_parent[i] = (int16_t)d.slot_parent[i];
    M.slot_count[i] = (int16_t)d.slot_count[i];
    if ((d.flags[i] & NB2_F_HANDOFF) && d.parent[i] != i - 1) { err = "handoff flag on a body whose parent is not i-1"; return false; }
    if (jt == NB2_JT_FREE) M.free_idx[i] = (int16_t)nfree++;
    ndof += (jt == NB2_JT_FREE) ? 6 : 1;
    for (int k = 0; k < 12; k++) M.Xtree[i][k] = (R)d.Xtree[12 * i + k];
    for (int k = 0; k < 10; k++) M.inertia[i][k] = (R)d.inertia[10 * i + k];
  }
  if (ndof != d.ndof) { err = "sum of joint dofs does not match ndof"; return false; }
  // ---- schedule: parse and validate (a wrong schedule would be a silent data race on the device)
  {
    const int K = d.lanes;
    if (K != 1 && K != 2 && K != 4 && K != 8) { err = "lanes must be 1, 2, 4 or 8"; return false; }
    M.lanes = K;
    for (int r = 0; r < NB2_MAX_RANGES; r++) M.trunk_lo[r] = M.trunk_hi[r] = 0;
    for (int l = 0; l < NB2_MAX_LANES; l++) { M.limb_n[l] = 0; for (int r = 0; r < NB2_MAX_RANGES; r++) M.limb_lo[l][r] = M.limb_hi[l][r] = 0; }
    int owner[NB2_MAX_BODIES], rng[NB2_MAX_BODIES];  // -1 trunk, else lane ; range id
    for (int i = 0; i < d.nb; i++) { owner[i] = -2; rng[i] = -1; }
    int pos = 0, rid = 0;
    auto take = [&](int& v) { if (pos >= d.nsched) return false; v = d.sched[pos++]; return true; };
    auto take_ranges = [&](int own, int16_t* lo, int16_t* hi, int& cnt) {
      if (!take(cnt) || cnt < 0 || cnt > NB2_MAX_RANGES) return false;
      int prev = 0;
      for (int r = 0; r < cnt; r++) {
        int a, b;
        if (!take(a) || !take(b) || a < prev || b <= a || b > d.nb) return false;
        lo[r] = (int16_t)a; hi[r] = (int16_t)b; prev = b;
        for (int i = a; i < b; i++) { if (owner[i] != -2) return false; owner[i] = own; rng[i] = rid; }
        rid++;
      }
      return true;
    };
    int tn = 0;
    if (!take_ranges(-1, M.trunk_lo, M.trunk_hi, tn)) { err = "malformed trunk schedule"; return false; }
    M.trunk_n = tn;
    for (int l = 0; l < K; l++) {
      int ln = 0;
      if (!take_ranges(l, M.limb_lo[l], M.limb_hi[l], ln)) { err = "malformed limb schedule"; return false; }
      M.limb_n[l] = (int16_t)ln;
    }
    for (int i = 0; i < d.nb; i++) {
      if (owner[i] == -2) { err = "schedule does not cover every body"; return false; }
      const int p = d.parent[i];
      if (p >= 0 && owner[p] != -1 && owner[p] != owner[i]) { err = "schedule: parent of a body is swept by another lane"; return false; }
      if ((d.flags[i] & NB2_F_HANDOFF) && rng[p] != rng[i]) { err = "handoff flag across schedule ranges"; return false; }
      if (p >= 0 && !(d.flags[i] & NB2_F_HANDOFF)) {
        const int sp = d.slot_parent[i];
        if (sp < d.slot_self[p] || sp >= d.slot_self[p] + d.slot_count[p] || sp >= d.nslots) { err = "child slot outside its parent's slots"; return false; }
      }
    }
  }
  M.nfree = nfree;
  for (int j = 0; j < NB2_MAX_DOFS; j++) {
    M.damping[j] = M.spring[j] = M.rest[j] = R(0);
    M.pos_lo[j] = M.vel_lo[j] = M.force_lo[j] = -__builtin_inff();
    M.pos_hi[j] = M.vel_hi[j] = M.force_hi[j] = __builtin_inff();
    M.action_map[j] = 0; M.act_of_dof[j] = -1;
  }
  for (int j = 0; j < d.ndof; j++) {
    M.damping[j] = (R)d.damping[j]; M.spring[j] = (R)d.spring[j]; M.rest[j] = (R)d.rest[j];
    M.pos_lo[j] = (float)d.pos_lo[j]; M.pos_hi[j] = (float)d.pos_hi[j];
    M.vel_lo[j] = (float)d.vel_lo[j]; M.vel_hi[j] = (float)d.vel_hi[j];
    M.force_lo[j] = (float)d.force_lo[j]; M.force_hi[j] = (float)d.force_hi[j];
  }
  for (int i = 0; i < d.na; i++) {
    if (d.action_map[i] < 0 || d.action_map[i] >= d.ndof) { err = "action map entry out of range"; return false; }
    M.action_map[i] = (int16_t)d.action_map[i];
    if (M.act_of_dof[d.action_map[i]] >= 0) { err = "the action map lists dof " + std::to_string(d.action_map[i]) + " twice"; return false; }
    M.act_of_dof[d.action_map[i]] = (int16_t)i;
  }
  return true;
}

static inline bool nb2_fill_contact(const nb2_model_desc& d, Nb2ContactDev& C, std::string& err) {
  if (d.nshapes < 0 || d.nshapes > NB2_MAX_SHAPES) { err = "model has " + std::to_string(d.nshapes) + " collision shapes; compiled limit is " + std::to_string(NB2_MAX_SHAPES); return false; }
  if (d.npairs < 0 || d.npairs > NB2_MAX_PAIRS) { err = "model has " + std::to_string(d.npairs) + " collision pairs; compiled limit is " + std::to_string(NB2_MAX_PAIRS); return false; }
  C.nshapes = d.nshapes; C.npairs = d.npairs; C.pen_correction = d.penetration_correction; C.pad_ = 0;
  C.clip_depth = d.contact_clipping_depth; C.fallback_cfm = d.fallback_cfm;
  for (int k = 0; k < NB2_MAX_CB; k++) C.cb_body[k] = -1;
  for (int s = 0; s < NB2_MAX_SHAPES; s++) {
    C.shape_body[s] = -1; C.shape_type[s] = 0; C.shape_orig_body[s] = -1; C.shape_mu[s] = 0; C.shape_rest[s] = 0;
    for (int k = 0; k < 3; k++) C.shape_dims[s][k] = 0;
    for (int k = 0; k < 12; k++) C.shape_T[s][k] = 0;
  }
  for (int s = 0; s < d.nshapes; s++) {
    if (d.shape_body[s] >= d.nb) { err = "shape attached to a body that does not exist"; return false; }
    if (d.shape_type[s] < 0 || d.shape_type[s] > 2) { err = "unsupported shape type"; return false; }
    C.shape_body[s] = (int16_t)d.shape_body[s]; C.shape_type[s] = (int16_t)d.shape_type[s]; C.shape_orig_body[s] = (int16_t)d.shape_orig_body[s];
    C.shape_mu[s] = d.shape_mu[s]; C.shape_rest[s] = d.shape_rest[s];
    for (int k = 0; k < 3; k++) C.shape_dims[s][k] = d.shape_dims[3 * s + k];
    for (int k = 0; k < 12; k++) C.shape_T[s][k] = d.shape_T[12 * s + k];
  }
  for (int p = 0; p < NB2_MAX_PAIRS; p++) { C.pair_a[p] = 0; C.pair_b[p] = 0; }
  for (int p = 0; p < d.npairs; p++) {
    if (d.pair_a[p] < 0 || d.pair_a[p] >= d.nshapes || d.pair_b[p] < 0 || d.pair_b[p] >= d.nshapes) { err = "collision pair references a missing shape"; return false; }
    C.pair_a[p] = (int16_t)d.pair_a[p]; C.pair_b[p] = (int16_t)d.pair_b[p];
  }
  // tree tables of the warp-cooperative contact stage: ancestor sets (bit masks: bodies are numbered parents-first, nb <= 64),
  // depths, and the list of collision bodies (moving bodies carrying at least one shape)
  for (int i = 0; i < NB2_MAX_BODIES; i++) { C.cb_of_body[i] = -1; C.cdof0[i] = 0; C.anc_mask[i] = 0ull; }
  for (int i = 0; i < d.nb; i++) {
    const int p = d.parent[i];
    C.anc_mask[i] = (1ull << i) | (p >= 0 ? C.anc_mask[p] : 0ull);
    C.cdof0[i] = (int16_t)(p >= 0 ? C.cdof0[p] + (d.jtype[p] == NB2_JT_FREE ? 6 : 1) : 0);
  }
  C.ncb = 0; C.max_chain_dofs = 1;
  for (int s = 0; s < d.nshapes; s++) {
    const int bdy = d.shape_body[s];
    if (bdy < 0 || C.cb_of_body[bdy] >= 0) continue;
    C.cb_of_body[bdy] = (int16_t)C.ncb; C.cb_body[C.ncb] = (int16_t)bdy; C.ncb++;
    const int cd = C.cdof0[bdy] + (d.jtype[bdy] == NB2_JT_FREE ? 6 : 1);
    if (cd > C.max_chain_dofs) C.max_chain_dofs = cd;
  }
  // joints with enforced position limits: their rows load the child body and (reaction) the parent body, which therefore join the list
  C.nlim = 0; C.pad2_ = 0;
  for (int l = 0; l < NB2_MAX_LIMITS; l++) C.lim_body[l] = -1;
  if (d.nlimits < 0 || d.nlimits > NB2_MAX_LIMITS) { err = "model has " + std::to_string(d.nlimits) + " joints with enforced limits; compiled limit is " + std::to_string(NB2_MAX_LIMITS); return false; }
  for (int l = 0; l < d.nlimits; l++) {
    const int i = d.limit_body ? d.limit_body[l] : -1;
    if (i < 0 || i >= d.nb || (d.jtype[i] != NB2_JT_REV && d.jtype[i] != NB2_JT_PRIS)) { err = "limit_body[" + std::to_string(l) + "] is not a revolute / prismatic body"; return false; }
    C.lim_body[C.nlim++] = (int16_t)i;
    for (int side = 0; side < 2; side++) {
      const int bdy = side ? d.parent[i] : i;
      if (bdy < 0 || C.cb_of_body[bdy] >= 0) continue;
      if (C.ncb >= NB2_MAX_CB) { err = "too many bodies take part in contacts / joint limits"; return false; }
      C.cb_of_body[bdy] = (int16_t)C.ncb; C.cb_body[C.ncb] = (int16_t)bdy; C.ncb++;
      const int cd = C.cdof0[bdy] + (d.jtype[bdy] == NB2_JT_FREE ? 6 : 1);
      if (cd > C.max_chain_dofs) C.max_chain_dofs = cd;
    }
  }
  // every pair this stage can generate contacts for must be of a supported shape combination: reject the others at model
  // creation instead of flagging them world by world at run time
  for (int p = 0; p < d.npairs; p++) {
    const int ta = d.shape_type[d.pair_a[p]], tb = d.shape_type[d.pair_b[p]];
    const bool ok = ta >= 0 && ta <= 2 && tb >= 0 && tb <= 2;  // box, sphere, capsule in any combination
    if (!ok) { err = "collision pair " + std::to_string(p) + ": shape types (" + std::to_string(ta) + ", " + std::to_string(tb) + ") have no contact generator (supported: box, sphere, capsule)"; return false; }
  }
  return true;
}
