// nb2_model_desc (C ABI, doubles) -> Nb2ModelDev<R> (kernel parameter block).  Host only.
#pragma once
#include <string>

#include "../../include/nb2.h"
#include "nb2_model.h"

template <class R>
static inline bool nb2_fill_model(const nb2_model_desc& d, Nb2ModelDev<R>& M, std::string& err) {
  if (d.nb <= 0 || d.nb > NB2_MAX_BODIES) { err = "model has " + std::to_string(d.nb) + " moving bodies; compiled limit is " + std::to_string(NB2_MAX_BODIES); return false; }
  if (d.ndof <= 0 || d.ndof > NB2_MAX_DOFS) { err = "model has " + std::to_string(d.ndof) + " dofs; compiled limit is " + std::to_string(NB2_MAX_DOFS); return false; }
  if (d.na < 0 || d.na > d.ndof) { err = "bad action map size"; return false; }
  M.nb = d.nb; M.ndof = d.ndof; M.na = d.na; M.nslots = d.nslots;
  M.pad0 = M.pad1 = M.pad2 = 0;
  M.dt = (R)d.dt;
  for (int k = 0; k < 3; k++) M.gravity[k] = (R)d.gravity[k];
  int nfree = 0, ndof = 0;
  for (int i = 0; i < NB2_MAX_BODIES; i++) {
    M.parent[i] = -1; M.jtype[i] = 0; M.dof_off[i] = 0; M.flags[i] = 0; M.slot_self[i] = -1; M.slot_parent[i] = -1; M.free_idx[i] = -1;
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
    if ((d.flags[i] & NB2_F_HANDOFF) && d.parent[i] != i - 1) { err = "handoff flag on a body whose parent is not i-1"; return false; }
    if (jt == NB2_JT_FREE) M.free_idx[i] = (int16_t)nfree++;
    ndof += (jt == NB2_JT_FREE) ? 6 : 1;
    for (int k = 0; k < 12; k++) M.Xtree[i][k] = (R)d.Xtree[12 * i + k];
    for (int k = 0; k < 10; k++) M.inertia[i][k] = (R)d.inertia[10 * i + k];
  }
  if (ndof != d.ndof) { err = "sum of joint dofs does not match ndof"; return false; }
  M.nfree = nfree;
  for (int j = 0; j < NB2_MAX_DOFS; j++) {
    M.damping[j] = M.spring[j] = M.rest[j] = R(0);
    M.pos_lo[j] = M.vel_lo[j] = M.force_lo[j] = -__builtin_inff();
    M.pos_hi[j] = M.vel_hi[j] = M.force_hi[j] = __builtin_inff();
    M.action_map[j] = 0;
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
  }
  return true;
}
