#!/bin/bash
# A/B of the step kernel at the long protocol: franka (default leg) and go2 (1000 steps after 300 warm-up), grasp.
O=gpurun_out/${1:-ab}
mkdir -p $O
timeout 900 python bench.py --steps 1000 --warmup 100 --no-cpu-baseline --no-extra-legs > $O/bench_franka.json 2> $O/bench_franka.err
timeout 900 python bench.py --steps 1000 --warmup 300 --no-cpu-baseline --workload go2 > $O/bench_go2.json 2> $O/bench_go2.err
python - <<PY
import json
for f in ("franka","go2"):
    try:
        d=json.loads(open("$O/bench_%s.json"%f).read().strip().splitlines()[-1])
        print(f, "value %.2fM"%(d["value"]/1e6), "kernel_ms %.4f"%d["roofline"]["kernel_ms"], "e2e %.2fM"%(d["e2e"]["value"]/1e6))
    except Exception as e: print(f, "failed", e)
PY
