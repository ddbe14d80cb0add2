#!/usr/bin/env python
"""Model-specialised builds of libgsrigid (the reference specialises its kernels per scene through its JIT, rigid_solver.py:187-218,
380-400; here it is an ahead-of-time build per model).

    python tools/specialize.py franka|go2|grasp [--build]

Builds the scene's model on the host (no GPU needed), asks the library for the specialisation header (gsr_model_spec_header: every
record / workspace offset and every build-time count / option as a literal) and, with --build, compiles
genesis_b200/csrc/libgsrigid.spec_<digest>.so with -DGSR_SPEC_HEADER for that model's group size only.  genesis_b200/_lib.py offers
such a copy to a solver first (spec_candidates); a copy built for another model refuses the commit and the generic library runs."""
import ctypes
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
CSRC = os.path.join(ROOT, "genesis_b200", "csrc")


def scene_for(name, n_envs):
    import genesis_b200 as gs
    from genesis_b200.bench_scenes import make_franka, make_franka_grasp, make_go2

    gs.init(backend=gs.gpu, seed=0)
    if name == "franka":
        return make_franka(n_envs, randomized=True, build=False)[0]
    if name == "go2":
        return make_go2(n_envs, build=False)[0]
    if name == "grasp":
        return make_franka_grasp(n_envs, build=False)[0]
    raise SystemExit(f"unknown model {name}")


def spec_header(solver, n_envs):
    """The header text for the solver's model (host-only: begin / set arrays + scalars / gsr_model_spec_header)."""
    from genesis_b200 import _lib

    L = _lib.lib()
    h = ctypes.c_void_p()
    _lib.check(None, L.gsr_model_begin(ctypes.byref(h)))
    solver._handle, solver._L = h, L
    if solver.model is None:
        solver.build_model()
    solver._upload_model_arrays()
    for k, v in solver.model_options().items():
        L.gsr_model_set_scalar(h, f"cfg_{k}".encode(), float(v))
    n = L.gsr_model_spec_header(h, n_envs, None, 0)
    if n <= 0:
        raise SystemExit(f"gsr_model_spec_header failed: {n}")
    buf = ctypes.create_string_buffer(n)
    L.gsr_model_spec_header(h, n_envs, buf, n)
    L.gsr_destroy(h)
    solver._handle = None
    return buf.value.decode()


def main():
    name = sys.argv[1]
    n_envs = {"franka": 30000, "go2": 4096, "grasp": 8192}.get(name, 1024)
    scene = scene_for(name, n_envs)
    text = spec_header(scene.rigid_solver, n_envs)
    from genesis_b200 import _lib

    os.makedirs(os.path.join(CSRC, "spec"), exist_ok=True)
    path = os.path.join(CSRC, "spec", f"spec_{name}.h")
    old = open(path).read() if os.path.exists(path) else None
    if old != text:
        open(path, "w").write(text)
    group = [l for l in text.splitlines() if l.startswith("#define GSR_SPEC_GROUP")][0].split()[-1]
    print(f"{path}: {len(text.splitlines())} lines, group size {group}")
    if "--build" in sys.argv:
        out = _lib.spec_path(text)  # named by the digest of (sources, description): RigidSolver.build finds it by recomputing both
        if os.path.exists(out):
            print(out, "(up to date)")
            return
        cmd = ["/usr/local/cuda/bin/nvcc", "-O3", "-std=c++17", "-lineinfo", "-gencode", "arch=compute_100a,code=sm_100a", "-prec-div=false",
               "-prec-sqrt=false", "-Xcompiler", "-fPIC", "-shared", f"-DGSR_SPEC_HEADER=\"spec/spec_{name}.h\"", f"-DGSR_DEV_ONLY_G={group}",
               *_lib.SPEC_NVCC_FLAGS,
               "-o", out, "gsr_api.cu", "-lcudart"]
        subprocess.check_call(cmd, cwd=CSRC)
        print(out)


if __name__ == "__main__":
    main()
