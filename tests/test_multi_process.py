"""bench.py's multi-rank protocol on the CPU (gloo, world size 2): env sharding has no data-path collective, so the only
distributed logic is rendezvous, the barrier and the max-over-ranks reduction of the timing."""

import os
import socket
import subprocess
import sys
import textwrap

from conftest import ROOT

WORKER = textwrap.dedent("""
    import os, sys, json
    import torch, torch.distributed as dist
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo")
    # contiguous env shards, as bench.py / DESIGN.md section 7: rank r owns [r*B, (r+1)*B)
    B = 6
    lo, hi = rank * B, (rank + 1) * B
    assert hi - lo == B
    dist.barrier()
    t = torch.tensor([1.0 + rank], dtype=torch.float64)        # pretend elapsed ms: slower on the last rank
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total = torch.tensor([float(B)])
    dist.all_reduce(total)                                       # whole-job units
    if rank == 0:
        print(json.dumps({"max_ms": t.item(), "total_envs": total.item(), "world": world}))
    dist.destroy_process_group()
""")


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_two_rank_timing_protocol(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), str(script)]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=240, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    import json

    line = [l for l in out.stdout.splitlines() if l.startswith("{")][-1]
    res = json.loads(line)
    assert res == {"max_ms": 2.0, "total_envs": 12.0, "world": 2}


def test_reference_arm_runs_on_rank0_only():
    """`bench.py --impl reference` under a 2-rank launch: rank 0 prints the line, the other rank exits 0 silently."""
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "3", "--warmup", "3"],
                         capture_output=True, text=True, timeout=240, env=env, cwd=ROOT)
    assert out.returncode == 0 and out.stdout.strip() == ""
    env = dict(os.environ, RANK="0", WORLD_SIZE="2", LOCAL_RANK="0")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "3", "--warmup", "3"],
                         capture_output=True, text=True, timeout=240, env=env, cwd=ROOT)
    assert out.returncode == 0
    import json

    res = json.loads(out.stdout.strip().splitlines()[-1])
    assert res["impl"] == "reference" and res["value"] > 0 and res["cpu_baseline"]["kind"] == "port"
