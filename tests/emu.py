"""TEST INFRASTRUCTURE ONLY: ctypes wrapper of tests/lane_emu/liblane_emu.so, the host emulation of the lane-per-environment
step kernel (one thread per environment, no warp intrinsics => the same functions compile for the CPU).  Lets the CPU test
tier check the kernel's arithmetic against the oracle without a GPU.  Never imported by the genesis_b200 package."""

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lane_emu")
_LIB = None

SCALARS = ("n_links", "n_joints", "n_dofs", "n_qs", "n_geoms", "n_entities", "n_equalities", "n_pairs", "n_verts",
           "max_contact_pairs", "max_broad_pairs", "n_contacts_per_pair", "len_constraints", "meaninertia", "dt")


def lib():
    global _LIB
    if _LIB is None:
        subprocess.check_call(["make", "-s", "-C", _HERE])
        L = ctypes.CDLL(os.path.join(_HERE, "liblane_emu.so"))
        L.gsr_model_begin.argtypes = [ctypes.POINTER(ctypes.c_void_p)]
        L.gsr_model_set_f32.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_void_p, ctypes.c_size_t]
        L.gsr_model_set_i32.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_void_p, ctypes.c_size_t]
        L.gsr_model_set_scalar.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_double]
        L.gsr_model_commit.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
        L.gsr_state_words.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_size_t)]
        L.gsr_field_count.argtypes = [ctypes.c_void_p]
        L.gsr_field_info.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(ctypes.c_char_p), ctypes.POINTER(ctypes.c_int)] + [ctypes.POINTER(ctypes.c_size_t)] * 6
        L.gsr_bind.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        L.gsr_init_state.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        L.gsr_step.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
        L.gsr_destroy.argtypes = [ctypes.c_void_p]
        L.gsr_last_error.argtypes = [ctypes.c_void_p]
        L.gsr_last_error.restype = ctypes.c_char_p
        _LIB = L
    return _LIB


class LaneEmu:
    def __init__(self, model, n_envs, options=None):
        L = lib()
        self.m, self.B = model, n_envs
        h = ctypes.c_void_p()
        L.gsr_model_begin(ctypes.byref(h))
        self.h = h
        for k, v in model.__dict__.items():
            if isinstance(v, np.ndarray):
                if v.dtype.kind == "f":
                    a = np.ascontiguousarray(v, dtype=np.float32)
                    L.gsr_model_set_f32(h, k.encode(), a.ctypes.data, a.size)
                elif v.dtype.kind in "iub":
                    a = np.ascontiguousarray(v, dtype=np.int32)
                    L.gsr_model_set_i32(h, k.encode(), a.ctypes.data, a.size)
        for k in SCALARS:
            L.gsr_model_set_scalar(h, k.encode(), float(getattr(model, k)))
        for i, ax in enumerate("xyz"):
            L.gsr_model_set_scalar(h, f"gravity_{ax}".encode(), float(model.gravity[i]))
        for k, v in (options or {}).items():
            L.gsr_model_set_scalar(h, f"cfg_{k}".encode(), float(v))
        L.gsr_model_set_scalar(h, b"cfg_diagnostics", 1.0)  # the emulation tests read the intermediate quantities of every step
        rc = L.gsr_model_commit(h, 0, n_envs)
        assert rc == 0, L.gsr_last_error(h)
        nw = ctypes.c_size_t()
        L.gsr_state_words(h, ctypes.byref(nw))
        self.F = np.zeros(nw.value, dtype=np.float32)
        self.I = self.F.view(np.int32)
        L.gsr_bind(h, self.F.ctypes.data)
        self.fields = {}
        name, is_int = ctypes.c_char_p(), ctypes.c_int()
        v = [ctypes.c_size_t() for _ in range(6)]
        for i in range(L.gsr_field_count(h)):
            L.gsr_field_info(h, i, ctypes.byref(name), ctypes.byref(is_int), *[ctypes.byref(x) for x in v])
            seg, stride, off, items, comps, istride = (x.value for x in v)
            buf = self.I if is_int.value else self.F
            base = buf[seg + off:]
            shape, strides = ((n_envs, items), (stride * 4, istride * 4)) if comps == 1 else ((n_envs, items, comps), (stride * 4, istride * 4, 4))
            self.fields[name.value.decode()] = np.lib.stride_tricks.as_strided(base, shape, strides)
        L.gsr_init_state(h, None)

    def __del__(self):
        if getattr(self, "h", None):
            lib().gsr_destroy(self.h)
            self.h = None

    def field(self, name):
        return self.fields[name]

    def rows(self, name):
        return np.array(self.fields[name]).reshape(self.B, -1)

    def set(self, name, value):
        f = self.fields[name]
        f[...] = np.asarray(value).reshape(f.shape)

    def step(self, n=1):
        for _ in range(n):
            lib().gsr_step(self.h, 1, None)
