"""ctypes binding of libgsrigid.so (include/gsrigid.h).  There is NO fallback: if the CUDA library cannot be loaded the
rigid path raises, it never degrades to a CPU or PyTorch implementation."""

import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GSR_LIB") or os.path.join(_HERE, "csrc", "libgsrigid.so")  # GSR_LIB: developer builds (make trace)
_lib = None
# A second copy of the library (libgsrigid.1.so, made by the Makefile) gives a second scene on the same GPU its own CUDA module and
# with it its own __constant__ model: two different models then neither re-upload their constants at every step nor race when they
# run on two streams.  Scenes beyond N_INSTANCES alive at once share instance 0 (one stream, constants re-uploaded on every switch).
N_INSTANCES = 2
_libs = {}
_slots = {}  # device index -> [owner id or None] * N_INSTANCES

EXPORTS = (
    "gsr_model_begin", "gsr_model_set_f32", "gsr_model_set_i32", "gsr_model_set_scalar", "gsr_model_commit", "gsr_model_refresh",
    "gsr_model_spec_header", "gsr_is_specialised", "gsr_state_words", "gsr_field_count", "gsr_field_info", "gsr_bind", "gsr_launch_info", "gsr_init_state", "gsr_step",
    "gsr_forward_kinematics", "gsr_forward_kinematics_dirty", "gsr_set_kinematics_policy", "gsr_kinematics_valid", "gsr_reset_envs", "gsr_compute_diagnostics", "gsr_set_state", "gsr_gather_fields", "gsr_step_host", "gsr_launch_count", "gsr_kernel_info",
    "gsr_set_profiling", "gsr_kernel_time_ms", "gsr_debug_trace", "gsr_debug_trace_read",
    "gsr_last_error", "gsr_version", "gsr_destroy",
)


class GsrError(RuntimeError):
    pass


def build(force=False):
    """Compile libgsrigid.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
    if force or not os.path.exists(LIB_PATH):
        subprocess.check_call(["make", "-C", os.path.join(_HERE, "csrc")] + (["-B"] if force else []))
    return LIB_PATH


def source_digest():
    """Digest of the CUDA / C++ sources a specialised copy was compiled from: a copy built from other sources is never picked up."""
    import hashlib

    h = hashlib.sha256()
    d = os.path.join(_HERE, "csrc")
    for f in sorted(os.listdir(d)):
        if f.endswith((".cu", ".cuh", ".h")):
            h.update(f.encode()); h.update(open(os.path.join(d, f), "rb").read())
    return h.hexdigest()[:16]


# extra flags of a model-specialised build: with a quarter less code than the generic kernels, inlining the kinematics, the small
# math helpers (rotate, quat_compose, quat_to_R, rotvec_to_quat), fk_link and the Hessian factorisation pays (Franka 87 -> 103 M
# env-steps/s, measured variant by variant: profiles/README.md item 22); in the generic kernels they stay real functions
SPEC_NVCC_FLAGS = ("-DGSR_KIN_INLINE=__forceinline__", "-DGSR_MATH_NOINLINE=", "-DGSR_FK_NOINLINE=", "-DGSR_HESS_NOINLINE=")


def spec_path(header_text):
    """Where the copy specialised for the model described by `header_text` (gsr_model_spec_header) lives, if it was built."""
    import hashlib

    tag = hashlib.sha256((source_digest() + " ".join(SPEC_NVCC_FLAGS) + header_text).encode()).hexdigest()[:12]
    return os.path.join(_HERE, "csrc", f"libgsrigid.spec_{tag}.so")


def lib_at(path):
    """A library copy by path (model-specialised builds); bound like the generic one."""
    if path not in _libs:
        _libs[path] = _bind(ctypes.CDLL(path))
    return _libs[path]


def instance_path(instance):
    return LIB_PATH if instance == 0 else LIB_PATH[:-3] + f".{instance}.so"


def acquire_instance(device_index, owner):
    """Lowest library instance no live solver uses on this device (0 when all are taken or the extra copy was not built)."""
    slots = _slots.setdefault(device_index, [None] * N_INSTANCES)
    for k in range(N_INSTANCES):
        if slots[k] is None and (k == 0 or os.path.exists(instance_path(k))):
            slots[k] = id(owner)
            return k
    return 0


def release_instance(device_index, instance, owner):
    slots = _slots.get(device_index)
    if slots is not None and isinstance(instance, int) and slots[instance] == id(owner):
        slots[instance] = None


def lib(instance=0):
    global _lib
    if instance not in _libs:
        path = instance_path(instance)
        if not os.path.exists(path):
            raise GsrError(
                f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` (there is no CPU fallback)"
            )
        _libs[instance] = _bind(ctypes.CDLL(path))
        if instance == 0:
            _lib = _libs[instance]
    return _libs[instance]


def _bind(L):
    if True:
        vp, cp, i, sz = ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int, ctypes.c_size_t
        L.gsr_model_begin.argtypes = [ctypes.POINTER(vp)]
        L.gsr_model_set_f32.argtypes = [vp, cp, vp, sz]
        L.gsr_model_set_i32.argtypes = [vp, cp, vp, sz]
        L.gsr_model_set_scalar.argtypes = [vp, cp, ctypes.c_double]
        L.gsr_model_commit.argtypes = [vp, i, i]
        L.gsr_model_refresh.argtypes = [vp]
        L.gsr_model_spec_header.argtypes = [vp, i, vp, sz]
        L.gsr_model_spec_header.restype = ctypes.c_longlong
        L.gsr_state_words.argtypes = [vp, ctypes.POINTER(sz)]
        L.gsr_field_count.argtypes = [vp]
        L.gsr_field_info.argtypes = [vp, i, ctypes.POINTER(cp), ctypes.POINTER(i)] + [ctypes.POINTER(sz)] * 6
        L.gsr_bind.argtypes = [vp, vp]
        L.gsr_launch_info.argtypes = [vp] + [ctypes.POINTER(i)] * 4
        L.gsr_init_state.argtypes = [vp, vp]
        L.gsr_step.argtypes = [vp, i, vp]
        L.gsr_forward_kinematics.argtypes = [vp, vp, vp]
        L.gsr_reset_envs.argtypes = [vp, vp, vp]
        L.gsr_forward_kinematics_dirty.argtypes = [vp, vp]
        L.gsr_set_kinematics_policy.argtypes = [vp, i]
        L.gsr_kinematics_valid.argtypes = [vp]
        L.gsr_compute_diagnostics.argtypes = [vp, vp]
        L.gsr_set_state.argtypes = [vp] + [vp] * 9 + [i, vp]
        L.gsr_gather_fields.argtypes = [vp, i, ctypes.POINTER(cp), ctypes.POINTER(vp), vp]
        L.gsr_step_host.argtypes = [vp, i, vp, vp, vp, vp]
        L.gsr_launch_count.argtypes = [vp]
        L.gsr_launch_count.restype = ctypes.c_longlong
        L.gsr_kernel_info.argtypes = [vp, i, ctypes.POINTER(cp), ctypes.POINTER(ctypes.c_longlong)]
        L.gsr_set_profiling.argtypes = [vp, i]
        L.gsr_kernel_time_ms.argtypes = [vp, i, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_longlong)]
        L.gsr_debug_trace.argtypes = [vp, i]
        L.gsr_debug_trace_read.argtypes = [vp, vp]
        L.gsr_last_error.argtypes = [vp]
        L.gsr_last_error.restype = cp
        L.gsr_version.restype = cp
        L.gsr_destroy.argtypes = [vp]
    return L


def check(handle, rc):
    if rc != 0:
        msg = lib().gsr_last_error(handle)  # reads the handle's own message buffer: any instance's code can do it
        raise GsrError(f"gsrigid error {rc}: {msg.decode() if msg else ''}")
