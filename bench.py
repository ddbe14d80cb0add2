#!/usr/bin/env python
"""bench.py -- env-steps/s of the batched rigid-body step (BASELINE.json metric) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload franka|go2|grasp|mixed] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

A "step" is one `scene.step()` (one rigid substep, dt = 0.01) of every environment resident on a GPU.  The headline leg is
BASELINE.json configs[1]: Franka Panda on a plane, position control, 30000 envs per GPU, no self-collision
(tests/test_rigid_benchmarks.py make_franka(is_collision_free=True) + random initial velocities, "franka_random").  The same
JSON line carries configs[2] (Go2, 4096 envs per GPU) under "go2" and, for N > 1, configs[4] (65536 envs, half Franka half
Go2, every GPU takes the same share of both models so the ranks are cost-balanced) under "mixed".  Environments shard across
GPUs with no collective on the step path (weak scaling); the only NCCL traffic is the barrier / max-reduce of the timing
protocol.

Rank 0 prints ONE JSON line (contract in the task statement): value = env-steps/s with state resident in HBM, e2e = the same
metric through the host-buffer plugin call gsr_step_host (pinned-host control targets in, qpos / qvel out, copies inside the
timed region), roofline = the dominant kernel's algorithmic HBM bytes / its CUDA-event time against MEASURED_PEAKS.json,
cpu_baseline = the CPU oracle (a port / restatement of the reference, see oracle/) on the host cores.

Timing protocol: before anything is timed every scene runs >= 2 periods of the cost-ordered launch (32 steps) so that the
launch order and the on-chip row capacity have settled -- nothing on the step path allocates (gsr_model_commit owns every
buffer); then `--warmup` steps, then exactly `--steps` steps between two CUDA events on the launching stream, barrier +
synchronize on both sides, max over ranks.  The reference's errno poll (every 10 substeps, simulator.py:277-283) stays inside
the loop; it is an asynchronous reduction into a pinned word here, not a host sync.  Clocks are sampled through NVML
(pynvml) by a thread started before the first event, never by forking nvidia-smi inside the window.  The state of every
workload is far larger than the 126 MB L2 (config.l2), so consecutive steps cannot hit in cache.

`--impl reference` times the reference's CPU implementation of the path on the SAME config (30000 Franka envs): the
reference itself (Quadrants JIT + MuJoCo model compiler) cannot be installed offline, so this is the oracle port with all
host threads, labelled kind="port"; it always runs for >= 2 s whatever --steps says.
"""

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (default envs per GPU, description)
    "franka": (30000, "franka_random: plane + panda.xml, PD position control at q0, vel0=0.2*clip(randn,-1,1), no self-collision"),
    "go2": (4096, "go2: plane + go2.urdf, joints uniform within limits, PD targets, self-collision on (reference defaults)"),
    "grasp": (8192, "franka_grasp: plane + panda.xml closing its fingers on a box of a cluster of four 4 cm boxes (finger-box MPR + box-box contacts)"),
    "mixed": (0, "mixed: 65536 envs in total, half Franka half Go2; every GPU steps 32768/N envs of each model back to back (cost-balanced)"),
}
INTERNAL_WARMUP = 32  # two periods of the cost-ordered launch (cfg_reorder_period = 16)
# Scenes that start with a drop are measured in steady state: SURVEY.md section 8(d) asks for >= 200 warm-up steps (the robots land,
# the solver leaves its multi-iteration regime, the on-chip row capacity has adapted); these run whatever --warmup says and are reported
# as `internal_warmup_steps`.  (The first 300 steps of the Go2 drop run at about half the steady-state rate.)
SETTLE_STEPS = {"franka": INTERNAL_WARMUP, "go2": 320, "mixed": 320, "grasp": 96}


def algorithmic_bytes_per_env_step(model, mean_contacts, mean_broad):
    """compulsory HBM traffic of one env-step (fp32 / int32 = 4 B)."""
    nQ, nD, nL, nG = model.n_qs, model.n_dofs, model.n_links, model.n_geoms
    words = 2 * nQ + 2 * nD + 4 * nD + 2 * nD + 2 * nD + 16 * nL + 7 * nG
    return 4 * words + 56 * mean_contacts + 24 * mean_broad + 8


class ClockSampler:
    """SM clock + throttle reasons through NVML (pynvml), sampled by a thread for as long as it is entered; samples taken
    while a timed window is open (`mark(True)` .. `mark(False)`) are the ones summarised."""

    def __init__(self, index):
        self.index, self.rows, self._stop, self._t, self._hot = index, [], threading.Event(), None, False
        self.err = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = index
            if vis:
                ids = [v for v in vis.split(",") if v.strip() != ""]
                if index < len(ids) and ids[index].strip().isdigit():
                    phys = int(ids[index])
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_sm = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception as e:  # no NVML: report it, never fork a process
            self.nv, self.err = None, f"{type(e).__name__}: {e}"

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                sm = int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                reasons = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                self.rows.append((self._hot, sm, reasons))
            except Exception as e:
                self.err = f"{type(e).__name__}: {e}"
            self._stop.wait(0.002 if self._hot else 0.02)

    def mark(self, hot):
        self._hot = hot

    def __enter__(self):
        if self.nv is not None:
            self._t = threading.Thread(target=self._run, daemon=True)
            self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._t is not None:
            self._t.join(timeout=2)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [f"NVML unavailable ({self.err})"], "source": "pynvml"}
        hot = [r for r in self.rows if r[0]] or self.rows
        sm = sorted(r[1] for r in hot)
        bits = 0
        for r in hot:
            bits |= r[2]
        names = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": self.max_sm, "reasons": [n for b, n in names.items() if bits & b],
                "samples": len(hot), "samples_total": len(self.rows), "source": "pynvml (thread started before the timed region; no fork)"}


def build_scene(workload, n_envs, seed):
    from genesis_b200.bench_scenes import make_franka, make_franka_grasp, make_go2

    if workload == "franka":
        return make_franka(n_envs, randomized=True, seed=seed)
    if workload == "grasp":
        scene, franka, _boxes = make_franka_grasp(n_envs)
        return scene, franka
    return make_go2(n_envs, seed=seed)


def make_oracle(workload, n_envs, threads):
    """The CPU oracle (port of the reference algorithm) set up on the same scene and the same seeded inputs as the GPU leg."""
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
    return orc


def oracle_throughput(workload, n_envs, threads, warmup, min_seconds, min_steps=1):
    """Times the oracle for at least `min_seconds` (and `min_steps` steps); returns (env-steps/s, seconds, steps)."""
    orc = make_oracle(workload, n_envs, threads)
    orc.step(max(1, warmup))
    probe_steps = max(1, min_steps // 4)
    probe = max(orc.step_timed(probe_steps) / probe_steps, 1e-6)
    steps = int(max(min_steps, min(100000, round(min_seconds / probe + 0.5))))
    sec = orc.step_timed(steps)
    return n_envs * steps / sec, sec, steps


def run_reference(args, rank):
    """--impl reference: the reference's CPU path on the host cores, on the SAME config as the GPU arm (all envs of the
    workload), for >= 2 s regardless of --steps (a 20-step run of the port is 0.3 s and dominated by thread start-up)."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    wl = args.workload if args.workload in ("franka", "go2") else "franka"
    n_envs, desc = WORKLOADS[wl]
    n_envs = args.envs or n_envs
    t0 = time.time()
    value, sec, steps = oracle_throughput(wl, n_envs, cores, args.warmup, 3.0, min_steps=max(args.steps, 4))
    sample = f"{n_envs} envs (the whole config) x {steps} steps on {cores} host threads ({sec:.2f} s of timed CPU work)"
    print(json.dumps({
        "impl": "reference", "metric": "env-steps/sec", "value": value, "unit": "env-steps/s", "n_gpus": args.gpus, "steps": steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * sec / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{wl}-{n_envs}", "description": desc, "envs_per_gpu": n_envs, "total_envs": n_envs, "dt": 0.01,
                   "solver": "Newton", "integrator": "approximate_implicitfast",
                   "note": "CPU arm: one host, all envs of one GPU's share; steps raised to cover >= 3 s"},
        "cpu_baseline": {"value": value, "unit": "env-steps/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "reference (Quadrants 0.5.2 JIT + MuJoCo) is not installable offline; this is oracle/ (C++ restatement), all host threads",
        "wall_s": time.time() - t0,
    }), flush=True)


class Leg:
    """One workload on this rank: scene(s), timing of the device-resident loop, the per-kernel leg and the host-buffer leg."""

    def __init__(self, workload, envs_per_gpu, rank, world, clocks, barrier, max_over_ranks):
        self.workload, self.rank, self.world, self.clocks = workload, rank, world, clocks
        self.barrier, self.max_over_ranks = barrier, max_over_ranks
        if workload == "mixed":
            # two models on one GPU: each scene gets its own copy of the library (its own CUDA module and __constant__ model,
            # genesis_b200/_lib.py), so they step concurrently on two streams and share the SMs
            per = envs_per_gpu // 2 if envs_per_gpu else 32768 // world
            self.parts = [("franka", per), ("go2", per)]
        else:
            self.parts = [(workload, envs_per_gpu)]
        self.scenes = [build_scene(w, b, seed=rank)[0] for w, b in self.parts]
        self.B = sum(b for _, b in self.parts)
        self.settled = False
        self.streams = None
        # ... when the launches are small (a few waves each): measured on 8 / 4 / 2 GPUs (4096 / 8192 / 16384 envs of each model per GPU)
        # concurrent 199.8 / 106.8 / 38.6 M env-steps/s against 149.3 / 104.5 / 56.6 M back to back -- two saturating launches only
        # get in each other's way
        small = max(b for _, b in self.parts) <= 8192
        if len(self.scenes) > 1 and len({s.rigid_solver._lib_instance for s in self.scenes}) == len(self.scenes) \
                and os.environ.get("GSR_BENCH_SERIAL_SCENES", "0" if small else "1") != "1":
            import torch

            self.streams = [torch.cuda.Stream() for _ in self.scenes]

    @property
    def solvers(self):
        return [s.rigid_solver for s in self.scenes]

    def step(self):
        if self.streams is None:
            for s in self.scenes:
                s.step()
            return
        import torch

        for s, st in zip(self.scenes, self.streams):
            with torch.cuda.stream(st):
                s.step()

    def fork_streams(self):
        """The per-scene streams start after everything already enqueued on the current stream (the timing event included)."""
        if self.streams is not None:
            import torch

            cur = torch.cuda.current_stream()
            for st in self.streams:
                st.wait_stream(cur)

    def join_streams(self):
        if self.streams is not None:
            import torch

            cur = torch.cuda.current_stream()
            for st in self.streams:
                cur.wait_stream(st)

    def settle(self):
        if not self.settled:
            for _ in range(SETTLE_STEPS.get(self.workload, INTERNAL_WARMUP)):
                self.step()
            self.settled = True

    def time_resident(self, K, W):
        import torch

        self.fork_streams()
        self.settle()
        for _ in range(W):
            self.step()
        self.join_streams()
        l0 = sum(s.launch_count() for s in self.solvers)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.barrier()
        self.clocks.mark(True)
        e0.record()
        self.fork_streams()
        for _ in range(K):
            self.step()
        self.join_streams()
        e1.record()
        self.barrier()
        self.clocks.mark(False)
        ms = self.max_over_ranks(e0.elapsed_time(e1))
        launches = sum(s.launch_count() for s in self.solvers) - l0
        return ms, launches

    def roofline(self, ms_per_step, n_prof):
        import torch

        out = []
        for (w, B), solver in zip(self.parts, self.solvers):
            m = solver.model
            solver.set_profiling(True)
            self.clocks.mark(True)
            for _ in range(n_prof):
                solver.substep(1)
            torch.cuda.synchronize()
            self.clocks.mark(False)
            ktimes = solver.kernel_times()
            solver.set_profiling(False)
            mean_contacts = float(solver.field("n_contacts").float().mean())
            mean_broad = float(solver.field("n_broad").float().mean())
            bytes_step = algorithmic_bytes_per_env_step(m, mean_contacts, mean_broad)
            ks = {k: v for k, v in ktimes.items() if v[1] > 0}
            total = sum(v[0] for v in ks.values())
            dom = max(ks, key=lambda k: ks[k][0])
            dom_ms = ks[dom][0] / ks[dom][1]
            peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
            if os.path.exists(peaks_path):
                peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
            else:
                peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
            achieved = B * bytes_step / (dom_ms * 1e-3) / 1e9
            r = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                 "peak_source": peak_src, "algorithmic_bytes_per_env_step": bytes_step, "units_per_launch": B, "kernel_ms": dom_ms,
                 "kernel_share_of_step": ks[dom][0] / total, "kernels_ms": {k: v[0] / v[1] for k, v in ks.items()},
                 "profiled_steps": n_prof, "mean_contacts_per_env": mean_contacts, "mean_broad_pairs_per_env": mean_broad,
                 "workload": w}
            tp = os.path.join(ROOT, "profiles", f"traffic_{w}.json")
            if os.path.exists(tp):
                r["traffic"] = json.load(open(tp)).get(dom)
            out.append(r)
        if len(out) == 1:
            out[0]["whole_step_frac"] = (self.B * out[0]["algorithmic_bytes_per_env_step"] / (ms_per_step * 1e-3) / 1e9) / out[0]["peak"]
            return out[0]
        return {"bound": "hbm", "parts": out}

    def time_e2e(self, Ke):
        import numpy as np
        import torch

        bufs = []
        for solver in self.solvers:
            m = solver.model
            B, nD, nQ = solver._B, m.n_dofs, m.n_qs
            ctrl_h = torch.empty((B, nD), dtype=torch.float32).pin_memory()
            ctrl_h.copy_(solver.field("ctrl_pos"))
            qpos_h = torch.empty((B, nQ), dtype=torch.float32).pin_memory()
            vel_h = torch.empty((B, nD), dtype=torch.float32).pin_memory()
            bufs.append((ctrl_h, qpos_h, vel_h, ctrl_h.numpy(), qpos_h.numpy(), vel_h.numpy()))
        staged0 = sum(s.kernel_launches().get("k_scatter", 0) for s in self.solvers)

        def one():
            for solver, b in zip(self.solvers, bufs):
                solver.step_host(b[3], b[4], b[5])

        for _ in range(5):
            one()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.barrier()
        self.clocks.mark(True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(Ke):
            one()
        e1.record()
        self.barrier()
        wall_ms = 1e3 * (time.perf_counter() - t0)
        self.clocks.mark(False)
        e2e_ms = self.max_over_ranks(max(e0.elapsed_time(e1), wall_ms))
        for b in bufs:
            assert np.isfinite(b[4]).all()
        staged = sum(s.kernel_launches().get("k_scatter", 0) for s in self.solvers) > staged0
        h2d = sum(4 * s._B * s.model.n_dofs for s in self.solvers)
        d2h = sum(4 * s._B * (s.model.n_qs + s.model.n_dofs) for s in self.solvers)
        return {"value": self.world * self.B * Ke / (e2e_ms * 1e-3), "unit": "env-steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": Ke, "ms_per_step": e2e_ms / Ke, "api": "gsr_step_host (C ABI) via RigidSolver.step_host",
                "timer": "max(CUDA events, host wall clock) over ranks; one stream synchronisation per call inside the window",
                "transfer": ("staged cudaMemcpyAsync + k_scatter / k_gather" if staged else
                             "one DMA of the targets into a dense device buffer read by k_step; results posted by k_step into the pinned host buffers over PCIe")}

    def run(self, K, W, cpu_baseline_seconds=0.0):
        ms, launches = self.time_resident(K, W)
        value = self.world * self.B * K / (ms * 1e-3)
        errno = max(int(s.field("errno_").max()) for s in self.solvers)
        roof = self.roofline(ms / K, max(K, 48))
        e2e = self.time_e2e(max(K, 20))
        s0 = self.solvers[0]
        m = s0.model
        cfg = {"workload": f"{self.workload}-{self.B if self.workload != 'mixed' else self.world * self.B}", "description": WORKLOADS[self.workload][1],
               "envs_per_gpu": self.B, "total_envs": self.world * self.B, "dt": 0.01, "solver": "Newton", "integrator": "approximate_implicitfast",
               "errno_max": errno, "internal_warmup_steps": SETTLE_STEPS.get(self.workload, INTERNAL_WARMUP),
               "errno_poll": "every 10 substeps inside the timed loop (asynchronous reduction, no host sync)",
               "l2": f"per-GPU state {sum(s._F.numel() for s in self.solvers) * 4 / 2**20:.0f} MiB > 126 MB L2 (no flush needed)",
               "parallelism": f"env-sharded x{self.world}, no step-path collective",
               "kinematics": ("lazy: the forward kinematics of every configuration run at the start of the next step instead of the end of this one "
                              "(same arithmetic per step; accessors that need link poses trigger them on demand)") if not s0._eager_fk else
                             "eager: forward kinematics at the end of every step (reference order)",
               "specialised_library": bool(s0.specialised)}
        if len(self.parts) == 1:
            cfg.update({"n_dofs": m.n_dofs, "n_links": m.n_links, "n_geoms": m.n_geoms, "launch": s0.launch_info,
                        "mean_contacts_per_env": roof["mean_contacts_per_env"]})
        else:
            cfg["parts"] = [{"workload": w, "envs_per_gpu": b, "launch": s.launch_info} for (w, b), s in zip(self.parts, self.solvers)]
        out = {"value": value, "unit": "env-steps/s", "ms_per_step": ms / K, "steps": K, "warmup": W, "config": cfg, "roofline": roof,
               "e2e": e2e, "gpu_launches": launches, "launches_per_step": launches / K}
        if cpu_baseline_seconds > 0 and self.workload in ("franka", "go2"):
            cores = os.cpu_count() or 1
            n_s = min(self.B, 64 * cores if self.workload == "franka" else 16 * cores)
            v, sec, k_s = oracle_throughput(self.workload, n_s, cores, 10, cpu_baseline_seconds, min_steps=20)
            out["cpu_baseline"] = {"value": v, "unit": "env-steps/s", "cores": cores, "kind": "port",
                                   "sample": f"{n_s} envs x {k_s} steps of the same scene, {cores} threads, {sec:.2f} s (oracle/: C++ restatement; the reference is not installable offline)"}
        return out

    def free(self):
        import torch

        for s in self.scenes:
            s.rigid_solver.destroy()
        self.scenes = []
        torch.cuda.empty_cache()


def accessors_leg(B, K, W, rank, world, clocks, barrier, max_over_ranks):
    """The reference's `franka_accessors` benchmark (tests/test_rigid_benchmarks.py:362-394): every step reads eight observation
    getters, resets 2 % of the environments with set_state(..., partial=True) and re-applies targets / gains / masked qpos and
    velocities with skip_forward=True -- the RL-loop boundary (scope row f3)."""
    import torch
    import genesis_b200 as gs
    from genesis_b200.bench_scenes import make_franka

    scene, franka = make_franka(B, randomized=True, seed=rank)
    s = scene.rigid_solver
    qpos0 = torch.tensor([0, 0, 0, -1.0, 0, 1.0, 0, 0.02, 0.02], device=gs.device).repeat(B, 1)
    vel0 = s.field("vel").clone()
    state0 = s.get_state()
    g = torch.Generator(device="cpu").manual_seed(rank)
    idx = torch.randperm(B, generator=g)[: max(int(0.02 * B), 1)].to(gs.device)
    mask = torch.zeros(B, dtype=torch.bool, device=gs.device)
    mask[idx] = True
    stiff, damp = franka.get_dofs_stiffness(), franka.get_dofs_damping()

    def step():
        scene.step()
        franka.get_ang(); franka.get_vel(); franka.get_dofs_position(); franka.get_dofs_velocity()
        franka.get_links_pos(); franka.get_links_quat(); franka.get_links_vel(); franka.get_contacts()
        s.set_state(0, state0, envs_idx=mask, partial=True)
        franka.control_dofs_position(qpos0)
        franka.set_dofs_stiffness(stiff); franka.set_dofs_damping(damp)
        franka.set_dofs_velocity(vel0, envs_idx=mask, skip_forward=True)
        franka.set_qpos(qpos0, envs_idx=mask, zero_velocity=False, skip_forward=True)

    for _ in range(INTERNAL_WARMUP + W):
        step()
    l0 = s.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    clocks.mark(True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(K):
        step()
    e1.record()
    barrier()
    wall_ms = 1e3 * (time.perf_counter() - t0)
    clocks.mark(False)
    ms = max_over_ranks(max(e0.elapsed_time(e1), wall_ms))
    out = {"value": world * B * K / (ms * 1e-3), "unit": "env-steps/s", "ms_per_step": ms / K, "steps": K,
           "config": {"workload": f"franka_accessors-{B}", "reset_fraction": 0.02,
                      "step_body": "tests/test_rigid_benchmarks.py:377-394 (8 getters incl. get_contacts, set_state(partial=True), masked setters, skip_forward)"},
           "gsr_launches_per_step": (s.launch_count() - l0) / K, "timer": "max(CUDA events, host wall clock): get_contacts reads the contact count on the host, as in the reference"}
    s.destroy()
    torch.cuda.empty_cache()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=100)
    ap.add_argument("--workload", choices=sorted(WORKLOADS), default="franka")
    ap.add_argument("--envs", type=int, default=0, help="envs per GPU (default: the workload's BASELINE size)")
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra-legs", action="store_true", help="skip the go2 / mixed legs of the default run")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        run_reference(args, rank)
        return

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

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    K, W = args.steps, args.warmup
    cpu_s = 0.0 if (args.no_cpu_baseline or world > 1 or rank != 0) else 8.0
    with ClockSampler(local) as clocks:
        B = args.envs or WORKLOADS[args.workload][0]
        leg = Leg(args.workload, B, rank, world, clocks, barrier, max_over_ranks)
        main_out = leg.run(K, W, cpu_baseline_seconds=cpu_s)
        leg.free()
        extras = {}
        if args.workload == "franka" and not args.no_extra_legs and not args.envs:
            g = Leg("go2", WORKLOADS["go2"][0], rank, world, clocks, barrier, max_over_ranks)
            extras["go2"] = g.run(K, W, cpu_baseline_seconds=cpu_s)
            g.free()
            extras["franka_accessors"] = accessors_leg(WORKLOADS["franka"][0], max(K, 20), W, rank, world, clocks, barrier, max_over_ranks)
            if world > 1:
                mx = Leg("mixed", 0, rank, world, clocks, barrier, max_over_ranks)
                extras["mixed"] = mx.run(K, W)
                mx.free()
    if rank == 0:
        line = {"metric": "env-steps/sec", "value": main_out["value"], "unit": "env-steps/s", "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": main_out["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "config": main_out["config"], "roofline": main_out["roofline"],
                "cpu_baseline": main_out.get("cpu_baseline"), "e2e": main_out["e2e"], "gpu_launches": main_out["gpu_launches"],
                "launches_per_step": main_out["launches_per_step"], "clocks": clocks.summary()}
        line.update(extras)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
