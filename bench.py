#!/usr/bin/env python
"""bench.py -- env-steps/s of the batched rigid-body step (BASELINE.json metric) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload franka|go2] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

A "step" is one `scene.step()` (one rigid substep, dt = 0.01) of every environment resident on a GPU.  At N = 1 the
workload is BASELINE.json configs[1]: Franka Panda on a plane, position control, 30000 envs, no self-collision
(tests/test_rigid_benchmarks.py make_franka(is_collision_free=True) + random initial velocities, "franka_random").
`--workload go2` times configs[2] (Go2, 4096 envs) instead.  Environments shard across GPUs with no collective on the
step path (weak scaling: 30000 envs per GPU); the only NCCL traffic is the barrier / max-reduce of the timing protocol.

Rank 0 prints ONE JSON line (contract in the task statement): value = env-steps/s with state resident in HBM,
e2e = the same metric through the host-buffer plugin call gsr_step_host (pinned-host control targets in, qpos/qvel out,
copies inside the timed region), roofline = the dominant kernel's algorithmic HBM bytes / its CUDA-event time against
MEASURED_PEAKS.json, cpu_baseline = the CPU oracle (a port/restatement of the reference, see oracle/) on the host cores.
`--impl reference` times the reference's CPU implementation of the path: the reference itself (Quadrants JIT + MuJoCo
model compiler) cannot be installed offline, so this is the oracle port with all host threads, labelled kind="port".
"""

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (factory, default envs per GPU, description)
    "franka": ("franka", 30000, "franka_random: plane + panda.xml, PD position control at q0, vel0=0.2*clip(randn,-1,1), no self-collision"),
    "go2": ("go2", 4096, "go2: plane + go2.urdf, joints uniform within limits, PD targets, self-collision on (reference defaults)"),
    # BASELINE.json configs[3] / configs[4]: parity-test cases, timed on request only (no CPU-baseline leg)
    "grasp": ("grasp", 8192, "franka_grasp: plane + panda.xml over a cluster of four 4 cm boxes (box-mesh MPR contacts), PD position control"),
    "mixed": ("mixed", 0, "mixed: even ranks step Franka scenes, odd ranks Go2 scenes, 65536 envs in total sharded by env over the GPUs"),
}


def algorithmic_bytes_per_env_step(model, mean_contacts, mean_broad):
    """compulsory HBM traffic of one env-step (fp32 / int32 = 4 B)."""
    nQ, nD, nL, nG = model.n_qs, model.n_dofs, model.n_links, model.n_geoms
    words = 2 * nQ + 2 * nD + 4 * nD + 2 * nD + 2 * nD + 16 * nL + 7 * nG
    return 4 * words + 56 * mean_contacts + 24 * mean_broad + 8


class ClockSampler:
    """nvidia-smi sampling DURING the timed region (B200_PROFILING.md 'clocks' line)."""

    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self._stop, self._t = index, [], threading.Event(), None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self._stop.wait(0.1)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(int(r[0]) for r in self.rows if r[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": int(self.rows[0][1]) if self.rows[0][1].isdigit() else None,
                "reasons": reasons, "samples": len(self.rows)}


def build_scene(workload, n_envs, seed):
    from genesis_b200.bench_scenes import make_franka, make_franka_grasp, make_go2

    if workload == "franka":
        return make_franka(n_envs, randomized=True, seed=seed)
    if workload == "grasp":
        scene, franka, _boxes = make_franka_grasp(n_envs)
        return scene, franka
    return make_go2(n_envs, seed=seed)


def oracle_throughput(workload, n_envs, steps, warmup, threads):
    """Times the CPU oracle (port of the reference algorithm) on a bounded sample of the same workload."""
    import numpy as np
    import genesis_b200 as gs
    from genesis_b200.bench_scenes import FRANKA_Q0, GO2_CTRL, GO2_Q0, make_franka, make_go2
    from oracle.pyoracle import Oracle

    gs.init(backend=gs.gpu, seed=0)
    rng = np.random.default_rng(0)
    if workload == "franka":
        scene, _ = make_franka(0, build=False)
        m = scene.rigid_solver.build_model()
        orc = Oracle(m, n_envs, double=False, options=scene.rigid_solver.model_options(), threads=threads)
        q = np.tile(np.array(FRANKA_Q0), (n_envs, 1))
        v = 0.2 * np.clip(rng.normal(size=(n_envs, 9)), -1, 1)
        v[:, 7:] = 0
        orc.set("qpos", q); orc.set("vel", v); orc.set("ctrl_pos", q)
        orc.set("ctrl_mode", np.full((n_envs, 9), int(gs.CTRL_MODE.POSITION)))
    else:
        scene, _ = make_go2(0, build=False)
        m = scene.rigid_solver.build_model()
        orc = Oracle(m, n_envs, double=False, options=scene.rigid_solver.model_options(), threads=threads)
        lim = np.asarray(m.dofs_limit).reshape(-1, 2)[6:]
        q = np.tile(np.array(GO2_Q0), (n_envs, 1))
        q[:, 7:] = lim[:, 0] + (lim[:, 1] - lim[:, 0]) * rng.random((n_envs, 12))
        ctrl = np.zeros((n_envs, 18)); ctrl[:, 6:] = GO2_CTRL
        mode = np.full((n_envs, 18), int(gs.CTRL_MODE.FORCE)); mode[:, 6:] = int(gs.CTRL_MODE.POSITION)
        orc.set("qpos", q); orc.set("ctrl_pos", ctrl); orc.set("ctrl_mode", mode)
    orc.step(warmup)
    if steps <= 0:  # calibrate to about `-steps` seconds of CPU work (bounded sample)
        probe = orc.step_timed(10)
        steps = int(max(10, min(5000, round(-steps * 10 / max(probe, 1e-6)))))
    sec = orc.step_timed(steps)
    return n_envs * steps / sec, sec, steps


def run_reference(args, rank):
    """--impl reference: the reference's CPU path on the host cores (oracle port; the reference cannot be installed offline)."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    if args.workload not in ("franka", "go2"):
        print(json.dumps({"impl": "reference", "unavailable": f"the CPU reference arm is implemented for the franka and go2 workloads only (asked: {args.workload})"}), flush=True)
        return
    name, n_envs_gpu, desc = WORKLOADS[args.workload]
    sample_envs = 64 * cores if args.workload == "franka" else 16 * cores
    t0 = time.time()
    value, sec, steps = oracle_throughput(args.workload, sample_envs, args.steps, args.warmup, cores)
    sample = f"{sample_envs} envs x {steps} steps of the same scene on {cores} host threads ({sec:.2f} s)"
    print(json.dumps({
        "impl": "reference", "metric": "env-steps/sec", "value": value, "unit": "env-steps/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * sec / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.workload}-{n_envs_gpu}", "description": desc, "sampled_envs": sample_envs},
        "cpu_baseline": {"value": value, "unit": "env-steps/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "reference (Quadrants 0.5.2 JIT + MuJoCo) is not installable offline; this is oracle/ (C++ restatement), all host threads",
        "wall_s": time.time() - t0,
    }), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=100)
    ap.add_argument("--workload", choices=sorted(WORKLOADS), default="franka")
    ap.add_argument("--envs", type=int, default=0, help="envs per GPU (default: the workload's BASELINE size)")
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import numpy as np
    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the rigid step has no CPU fallback (use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    import genesis_b200 as gs

    gs.init(backend=gs.gpu, seed=rank)
    name, n_envs_default, desc = WORKLOADS[args.workload]
    local_workload = args.workload
    if args.workload == "mixed":  # heterogeneous models, one per GPU: no cross-GPU traffic either way
        local_workload = "franka" if rank % 2 == 0 else "go2"
        n_envs_default = 65536 // world
    B = args.envs or n_envs_default
    scene, robot = build_scene(local_workload, B, seed=rank)
    solver = scene.rigid_solver
    m = solver.model
    K, W = args.steps, args.warmup

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident throughput (value)
    for _ in range(W):
        scene.step()
    launches0 = solver.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    with ClockSampler(local) as clocks:
        e0.record()
        for _ in range(K):
            scene.step()
        e1.record()
        barrier()
    ms = max_over_ranks(e0.elapsed_time(e1))
    launches = solver.launch_count() - launches0
    errno = int(solver.field("errno_").max())
    value = world * B * K / (ms * 1e-3)

    # ---- per-kernel CUDA-event timing of the same steps (roofline leg), not under any profiler
    solver.set_profiling(True)
    for _ in range(min(K, 200)):
        solver.substep(1)
    ktimes = solver.kernel_times()
    solver.set_profiling(False)
    mean_contacts = float(solver.field("n_contacts").float().mean())
    bytes_step = algorithmic_bytes_per_env_step(m, mean_contacts, mean_contacts)
    step_kernels = {k: v for k, v in ktimes.items() if v[1] > 0}
    total_kernel_ms = sum(v[0] for v in step_kernels.values())  # all launches of the profiled steps (the reorder kernels run 1 step in 16)
    dom = max(step_kernels, key=lambda k: step_kernels[k][0])
    dom_ms = step_kernels[dom][0] / step_kernels[dom][1]
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    # the step IS one kernel (k_step): its algorithmic bytes are the whole step's compulsory traffic (DESIGN.md "Roofline")
    achieved = B * bytes_step / (dom_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                "peak_source": peak_src, "algorithmic_bytes_per_env_step": bytes_step, "kernel_ms": dom_ms,
                "kernel_share_of_step": step_kernels[dom][0] / total_kernel_ms,
                "kernels_ms": {k: v[0] / v[1] for k, v in step_kernels.items()},
                "whole_step_frac": (B * bytes_step / (ms / K * 1e-3) / 1e9) / peak}
    traffic_path = os.path.join(ROOT, "profiles", f"traffic_{args.workload}.json")
    if os.path.exists(traffic_path):
        roofline["traffic"] = json.load(open(traffic_path)).get(dom)

    # ---- end-to-end through the host-buffer plugin call (pinned host control targets in, qpos / qvel out)
    nD, nQ = m.n_dofs, m.n_qs
    ctrl_h = torch.empty((B, nD), dtype=torch.float32).pin_memory()
    ctrl_h.copy_(solver.field("ctrl_pos"))
    qpos_h = torch.empty((B, nQ), dtype=torch.float32).pin_memory()
    vel_h = torch.empty((B, nD), dtype=torch.float32).pin_memory()
    ctrl_np, qpos_np, vel_np = ctrl_h.numpy(), qpos_h.numpy(), vel_h.numpy()
    Ke = max(10, min(K, 200))
    staged0 = solver.kernel_launches().get("k_scatter", 0)
    for _ in range(3):
        solver.step_host(ctrl_np, qpos_np, vel_np)
    barrier()
    t0 = time.perf_counter()
    e0.record()
    for _ in range(Ke):
        solver.step_host(ctrl_np, qpos_np, vel_np)
    e1.record()
    barrier()
    e2e_ms = max_over_ranks(max(e0.elapsed_time(e1), 1e3 * (time.perf_counter() - t0)))
    e2e = {"value": world * B * Ke / (e2e_ms * 1e-3), "unit": "env-steps/s", "h2d_bytes_per_step": 4 * B * nD,
           "d2h_bytes_per_step": 4 * B * (nQ + nD), "steps": Ke, "ms_per_step": e2e_ms / Ke,
           "api": "gsr_step_host (C ABI) via RigidSolver.step_host",
           "transfer": ("staged cudaMemcpyAsync + k_scatter / k_gather" if solver.kernel_launches().get("k_scatter", 0) > staged0 else
                        "pinned host buffers read / written in place over PCIe by k_step (no staging launches)")}
    assert np.isfinite(qpos_np).all()

    # ---- CPU baseline (rank 0, N = 1 only): bounded sample of the same workload on the host cores
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and args.workload in ("franka", "go2"):
        cores = os.cpu_count() or 1
        n_s = 64 * cores if args.workload == "franka" else 16 * cores
        v, sec, k_s = oracle_throughput(args.workload, n_s, -12, 10, cores)  # ~12 s of CPU work
        cpu_baseline = {"value": v, "unit": "env-steps/s", "cores": cores, "kind": "port",
                        "sample": f"{n_s} envs x {k_s} steps of the same scene, {cores} threads, {sec:.2f} s (oracle/: C++ restatement; the reference is not installable offline)"}

    if rank == 0:
        print(json.dumps({
            "metric": "env-steps/sec", "value": value, "unit": "env-steps/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{args.workload}-{B}", "description": desc, "envs_per_gpu": B, "total_envs": world * B, "dt": 0.01,
                       "solver": "Newton", "integrator": "approximate_implicitfast", "n_dofs": nD, "n_links": m.n_links, "n_geoms": m.n_geoms,
                       "mean_contacts_per_env": mean_contacts, "errno_max": errno,
                       "l2": f"per-GPU state {solver._F.numel() * 4 / 2**20:.0f} MiB > 126 MB L2 (no flush needed)",
                       "launch": solver.launch_info, "parallelism": f"env-sharded x{world}, no step-path collective"},
            "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e, "gpu_launches": launches,
            "launches_per_step": launches / K, "clocks": clocks.summary(),
        }), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
