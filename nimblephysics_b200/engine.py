"""Device-side handle of a World: builds the canonical model and owns the nb2_model created through the C ABI."""
from __future__ import annotations

import ctypes

import numpy as np

from . import _cabi
from .modelspec import CanonModel, RawModel, compile_model, flatten_world

FP32, FP64 = 0, 1


class DeviceModel:
    """nb2_model wrapper (include/nb2.h).  One per (World version)."""

    def __init__(self, cm: CanonModel, schedules=()):
        """cm: the canonical model (its own schedule, normally lanes=1, is what the contact kernels sweep with);
        schedules: further compilations of the SAME RawModel with other lane counts (modelspec.compile_model(raw, lanes=K));
        the library picks one per launch from the batch size (include/nb2.h nb2_model_add_schedule)."""
        self.cm = cm
        self.ndof = cm.ndof
        self.na = len(cm.action_map)
        L = _cabi.lib()
        self._desc, self._keep = _cabi.make_desc(cm, with_contacts=len(cm.shape_body) > 0 or len(getattr(cm, "limit_bodies", [])) > 0)
        h = ctypes.c_void_p()
        _cabi.check(L.nb2_model_create(ctypes.byref(self._desc), ctypes.byref(h)))
        self.handle = h
        self.saved_words = L.nb2_saved_words_per_world(h)
        self.has_contacts = bool(L.nb2_model_has_contacts(h))
        self.schedules = [cm]
        for c in schedules:
            d, keep = _cabi.make_desc(c, with_contacts=False)
            _cabi.check(L.nb2_model_add_schedule(h, ctypes.byref(d)))
            self.schedules.append(c)

    @classmethod
    def from_raw(cls, raw: RawModel, lanes=(1, 2, 4, 8), contacts=True):
        """Compile `raw` once per lane count that shortens the sequential depth of a sweep (trunk + longest limb set)."""
        cms, best = [], None
        for K in sorted(lanes):
            try:
                c = compile_model(raw, lanes=K)
            except ValueError:
                if K == 1:
                    raise
                continue
            depth = sum(hi - lo for lo, hi in c.trunk_ranges) + max(sum(hi - lo for lo, hi in rs) for rs in c.limb_ranges)
            if best is None or depth < best:
                cms.append(c)
                best = depth
        if not contacts:
            for c in cms:
                c.shape_body = c.shape_body[:0]
        return cls(cms[0], cms[1:])

    def refresh_inertia(self, world):
        """The World's masses / COMs / moments changed (World.setMasses): recompute the canonical inertias (welded bodies
        summed into their owner) and push them to the device model; the tree, schedules and slots are unchanged."""
        from .modelspec import body_inertia_contribution

        raw = world._raw_model
        k = 0
        for sk in world.skeletons:
            for b in sk._ordered_bodies():
                raw.mass[k], raw.com[k] = b.mass, b.com
                I = b.moment
                raw.moment[k] = [I[0, 0], I[1, 1], I[2, 2], I[0, 1], I[0, 2], I[1, 2]]
                k += 1
        cm = self.cm
        inertia = np.zeros((cm.nb, 10))
        for i in range(raw.nb):
            o = int(cm.body_owner[i])
            if o >= 0:
                inertia[o] += body_inertia_contribution(cm.body_T[i], raw.mass[i], raw.com[i], raw.moment[i])
        for c in self.schedules:
            c.inertia = inertia.copy()
        buf = np.ascontiguousarray(inertia, dtype=np.float64)
        _cabi.check(_cabi.lib().nb2_model_set_inertia(self.handle, buf.ctypes.data))

    def inertia_param_jacobian(self, world) -> np.ndarray:
        """[getMassDims(), 10*nb] : d(canonical inertia parameters)/d(mass vector) at the world's current values."""
        from .modelspec import inertia_param_jacobian

        return inertia_param_jacobian(world._raw_model, self.cm, world._mass_entries())

    def set_lanes(self, lanes: int):
        """Pin the lane count (0 = automatic choice per launch)."""
        _cabi.check(_cabi.lib().nb2_model_set_lanes(self.handle, int(lanes)))

    def lanes_for(self, B: int, backward: bool = False, precision: int = FP32) -> int:
        return int(_cabi.lib().nb2_model_lanes_for(self.handle, int(B), int(backward), int(precision)))

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                _cabi.lib().nb2_model_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    # ---- device pointers (torch tensors on the current CUDA device) ----
    def forward_device(self, B, state_ptr, action_ptr, next_ptr, saved_ptr, stream, precision=FP32):
        _cabi.check(_cabi.lib().nb2_step_forward(self.handle, B, state_ptr, action_ptr, next_ptr, saved_ptr, precision, stream))

    def backward_device(self, B, state_ptr, action_ptr, saved_ptr, gnext_ptr, gstate_ptr, gaction_ptr, stream,
                        precision=FP32, ginertia_ptr=None):
        """ginertia_ptr: optional [10*nb, B] float32 buffer receiving dL/d(inertia parameters) per body and world."""
        _cabi.check(_cabi.lib().nb2_step_backward(self.handle, B, state_ptr, action_ptr, saved_ptr, gnext_ptr,
                                                  gstate_ptr, gaction_ptr, ginertia_ptr, precision, stream))

    def rollout_forward_device(self, B, T, states_ptr, actions_ptr, saved_ptr, stream, precision=FP32):
        _cabi.check(_cabi.lib().nb2_rollout_forward(self.handle, B, T, states_ptr, actions_ptr, saved_ptr, precision, stream))

    def rollout_backward_device(self, B, T, states_ptr, actions_ptr, saved_ptr, gstates_ptr, gactions_ptr, stream, precision=FP32):
        _cabi.check(_cabi.lib().nb2_rollout_backward(self.handle, B, T, states_ptr, actions_ptr, saved_ptr, gstates_ptr,
                                                     gactions_ptr, precision, stream))

    def rollout_contact_tape_bytes(self, B, T, checkpoint_every):
        return int(_cabi.lib().nb2_rollout_contact_tape_bytes(self.handle, B, T, checkpoint_every))

    def rollout_forward_contact_device(self, B, T, states_ptr, actions_ptr, x_ptr, m_ptr, tape_ptr, checkpoint_every, ws_ptr, sticky_ptr, stream):
        _cabi.check(_cabi.lib().nb2_rollout_forward_contact(self.handle, B, T, states_ptr, actions_ptr, x_ptr, m_ptr, tape_ptr, checkpoint_every,
                                                            ws_ptr, sticky_ptr, stream))

    def rollout_backward_contact_device(self, B, T, states_ptr, actions_ptr, x_ptr, m_ptr, tape_ptr, checkpoint_every, gstates_ptr, gactions_ptr,
                                        ws_ptr, sticky_ptr, stream):
        _cabi.check(_cabi.lib().nb2_rollout_backward_contact(self.handle, B, T, states_ptr, actions_ptr, x_ptr, m_ptr, tape_ptr, checkpoint_every,
                                                             gstates_ptr, gactions_ptr, ws_ptr, sticky_ptr, stream))

    def forward_dynamics(self, pos, vel, force):
        """q-ddot [B, n] (float64 CUDA tensors in and out): pointer-style ABA, no integration, no contact stage
        (SimpleFeatherstone::forwardDynamics / Skeleton::computeForwardDynamics + getAccelerations)."""
        import torch

        dev = pos.device
        f = lambda t: t.to(device=dev, dtype=torch.float64).contiguous()
        pos, vel, force = f(pos), f(vel), f(force)
        acc = torch.empty_like(pos)
        with torch.cuda.device(dev):
            _cabi.check(_cabi.lib().nb2_forward_dynamics(self.handle, pos.shape[0], pos.data_ptr(), vel.data_ptr(), force.data_ptr(), acc.data_ptr(),
                                                         torch.cuda.current_stream().cuda_stream))
        return acc

    def contact_workspace_bytes(self, B):
        return int(_cabi.lib().nb2_contact_workspace_bytes(self.handle, B))

    def contact_record_bytes(self, B):
        return int(_cabi.lib().nb2_contact_record_bytes(self.handle, B))

    def forward_contact_device(self, B, state_ptr, action_ptr, next_ptr, saved_ptr, ws_ptr, x_ptr, m_ptr, labels_ptr,
                               status_ptr, nc_ptr, cinfo_ptr, crec_ptr, status_accum_ptr, stream):
        """fused fp64 step with the contact / boxed-LCP stage, one warp per world (include/nb2.h nb2_step_forward_contact)."""
        _cabi.check(_cabi.lib().nb2_step_forward_contact(self.handle, B, state_ptr, action_ptr, next_ptr, saved_ptr, ws_ptr, x_ptr,
                                                         m_ptr, labels_ptr, status_ptr, nc_ptr, cinfo_ptr, crec_ptr, status_accum_ptr, stream))

    def backward_contact_device(self, B, state_ptr, action_ptr, saved_ptr, crec_ptr, ws_ptr, gnext_ptr, gstate_ptr, gaction_ptr,
                                stream, ginertia_ptr=None, status_ptr=None):
        """status_ptr: the forward's status array; worlds that cannot be back-propagated get bit 2048 (and NaN gradients)."""
        _cabi.check(_cabi.lib().nb2_step_backward_contact(self.handle, B, state_ptr, action_ptr, saved_ptr, crec_ptr, ws_ptr,
                                                          gnext_ptr, gstate_ptr, gaction_ptr, ginertia_ptr, status_ptr, stream))

    def set_contact_capacity(self, max_contacts: int):
        _cabi.check(_cabi.lib().nb2_model_set_contact_capacity(self.handle, int(max_contacts)))

    def contact_capacity(self) -> int:
        return int(_cabi.lib().nb2_model_contact_capacity(self.handle))

    # ---- host pointers (numpy / CPU tensors): copies included ----
    def forward_host(self, state: np.ndarray, action: np.ndarray, keep_for_backward=True, precision=FP32,
                     out: np.ndarray = None) -> np.ndarray:
        B = state.shape[0]
        if out is None:
            out = np.empty_like(state)
        _cabi.check(_cabi.lib().nb2_step_forward_host(self.handle, B, state.ctypes.data, action.ctypes.data,
                                                      out.ctypes.data, int(keep_for_backward), precision))
        return out

    def backward_host(self, grad_next: np.ndarray, precision=FP32, out_state=None, out_action=None):
        B = grad_next.shape[0]
        gs = np.empty_like(grad_next) if out_state is None else out_state
        ga = np.empty((B, self.na), np.float32) if out_action is None else out_action
        _cabi.check(_cabi.lib().nb2_step_backward_host(self.handle, B, grad_next.ctypes.data, gs.ctypes.data,
                                                       ga.ctypes.data, precision))
        return gs, ga

    def forward_contact_host(self, state: np.ndarray, action: np.ndarray, keep_for_backward=True, reset_cache=False, out=None, status_out=None):
        """Contact step on HOST arrays (float32 [B, 2n] / [B, a]); the solver cache lives in the model between calls."""
        B = state.shape[0]
        if out is None:
            out = np.empty_like(state)
        _cabi.check(_cabi.lib().nb2_step_forward_contact_host(self.handle, B, state.ctypes.data, action.ctypes.data, out.ctypes.data,
                                                              int(keep_for_backward), int(reset_cache),
                                                              status_out.ctypes.data if status_out is not None else None))
        return out

    def backward_contact_host(self, grad_next: np.ndarray, out_state=None, out_action=None, sticky_out=None):
        B = grad_next.shape[0]
        gs = np.empty_like(grad_next) if out_state is None else out_state
        ga = np.empty((B, self.na), np.float32) if out_action is None else out_action
        _cabi.check(_cabi.lib().nb2_step_backward_contact_host(self.handle, B, grad_next.ctypes.data, gs.ctypes.data, ga.ctypes.data,
                                                               sticky_out.ctypes.data if sticky_out is not None else None))
        return gs, ga


def device_model_for(world) -> DeviceModel:
    """Lazily (re)build the device model of a World.  The cache is validated against the object graph: any setter of a BodyNode / Joint /
    Skeleton / ShapeNode bumps a global edit epoch (world.py); a World whose model was built at an older epoch is re-flattened and,
    when its description really changed, rebuilt (with its LCP cache dropped)."""
    import hashlib

    from .world import edit_epoch

    dm = getattr(world, "_device_model", None)
    if dm is not None and getattr(world, "_dm_epoch", -1) == edit_epoch():
        return dm
    raw = flatten_world(world)
    h = hashlib.sha1(raw.to_json().encode()).hexdigest() + ("|nc" if getattr(world, "_contacts_disabled", False) else "")
    if dm is not None and getattr(world, "_dm_hash", None) == h:
        world._dm_epoch = edit_epoch()
        return dm
    world._raw_model = raw
    # contact-free step requested explicitly -> contacts=False
    dm = DeviceModel.from_raw(raw, contacts=not getattr(world, "_contacts_disabled", False))
    world._device_model = dm
    world._dm_hash, world._dm_epoch = h, edit_epoch()
    world._lcp_cache = None
    return dm
