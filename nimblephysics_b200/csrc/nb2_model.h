// Device-side model description (POD, passed to kernels as a __grid_constant__ parameter so that every
// access is a constant-bank load with a warp-uniform index: all worlds of a batch share one model).
// Layout produced by nimblephysics_b200/modelspec.py::compile_model (canonical frames: each body frame sits on
// its joint frame with the joint axis along +z; welds folded into parents; DFS pre-order numbering).
#pragma once
#include <stdint.h>

#define NB2_MAX_BODIES 64
#define NB2_MAX_DOFS 96

// canonical joint types
#define NB2_JT_REV 1   // rotation about +z of the body frame:     S = [0,0,1,0,0,0]
#define NB2_JT_PRIS 2  // translation along +z of the body frame:  S = [0,0,0,0,0,1]
#define NB2_JT_FREE 3  // q = [log R; p], qdot = body twist:        S = I6  (DART_USE_IDENTITY_JACOBIAN build)

// flags
#define NB2_F_HANDOFF 1   // parent == i-1 and swept by the same lane: leaf->root sweeps hand the contribution over in registers
#define NB2_F_HAS_SLOT 4  // body owns slot_count[i] consecutive slots starting at slot_self[i], one per non-handoff child

// cooperative lanes: up to NB2_MAX_LANES threads share one world; each owns up to NB2_MAX_RANGES contiguous body ranges
#define NB2_MAX_LANES 8
#define NB2_MAX_RANGES 8

template <class R>
struct Nb2ModelDev {
  int nb, ndof, na, nslots;
  int nfree;        // number of FREE bodies
  int lanes;        // threads cooperating on one world (1, 2, 4 or 8)
  int trunk_n, pad2;
  unsigned magic_n2, magic_n, magic_na, pad3;  // ceil(2^32 / x): idx / x == umulhi(idx, magic) for idx < 65536 (group I/O index math)
  R dt;
  R gravity[3];
  int16_t parent[NB2_MAX_BODIES];
  int16_t jtype[NB2_MAX_BODIES];
  int16_t dof_off[NB2_MAX_BODIES];
  int16_t flags[NB2_MAX_BODIES];
  int16_t slot_self[NB2_MAX_BODIES];
  int16_t slot_parent[NB2_MAX_BODIES];  // the slot THIS body writes its contribution to (-1: handoff or root)
  int16_t slot_count[NB2_MAX_BODIES];
  // schedule: lane 0 sweeps the trunk (ancestor-closed), every lane sweeps its limb subtrees (half-open body ranges)
  int16_t trunk_lo[NB2_MAX_RANGES], trunk_hi[NB2_MAX_RANGES];
  int16_t limb_n[NB2_MAX_LANES];
  int16_t limb_lo[NB2_MAX_LANES][NB2_MAX_RANGES], limb_hi[NB2_MAX_LANES][NB2_MAX_RANGES];
  int16_t free_idx[NB2_MAX_BODIES];  // index among FREE bodies or -1
  R Xtree[NB2_MAX_BODIES][12];       // R row-major (9), p (3): x_parent = R x_child + p at q = 0
  R inertia[NB2_MAX_BODIES][10];     // m, h(3) = m c, Ibar(6: xx,yy,zz,xy,xz,yz) about the body origin
  R damping[NB2_MAX_DOFS];
  R spring[NB2_MAX_DOFS];
  R rest[NB2_MAX_DOFS];
  float pos_lo[NB2_MAX_DOFS], pos_hi[NB2_MAX_DOFS];  // limits are compared against fp32 I/O values
  float vel_lo[NB2_MAX_DOFS], vel_hi[NB2_MAX_DOFS];
  float force_lo[NB2_MAX_DOFS], force_hi[NB2_MAX_DOFS];
  int16_t action_map[NB2_MAX_DOFS];
  int16_t act_of_dof[NB2_MAX_DOFS];  // inverse of action_map: action index driving a dof, -1 = unactuated
};

// number of fp32 words the forward pass saves per world for the backward pass
//   per body: V(6) A(6) U(6) psi(1) sc(2) = 21 ; per FREE body: inverse articulated inertia (21) + joint R,p (12)
//   per dof : qdd
#ifdef __CUDACC__
__host__ __device__
#endif
static inline int nb2_saved_words(int nb, int ndof, int nfree) { return nb * 21 + nfree * 33 + ndof; }
