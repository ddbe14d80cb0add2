#!/bin/bash
# GPU suite, Go2 / Franka bench lines and the Go2 straggler report
O=gpurun_out/${1:-r2t}
mkdir -p $O
( timeout 2400 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.log )
grep -E "^FAILED|^ERROR|passed|failed" $O/pytest_gpu.log | head -40
timeout 600 python bench.py --workload go2 --steps 1000 --warmup 5 > $O/bench_go2.json 2> $O/bench_go2.err
timeout 600 python bench.py --steps 1000 --warmup 5 > $O/bench_franka.json 2> $O/bench_franka.err
timeout 600 python bench.py --workload grasp --steps 100 --warmup 5 > $O/bench_grasp.json 2> $O/bench_grasp.err
python - <<PY
import json
for n in ("bench_go2", "bench_franka", "bench_grasp"):
    try:
        d = json.loads(open("$O/%s.json" % n).read().strip().splitlines()[-1])
        print(n, "value %.4gM e2e %.4gM kernel_ms %.4f" % (d["value"] / 1e6, d["e2e"]["value"] / 1e6, d["roofline"]["kernel_ms"]), "go2 leg:", (d.get("go2") or {}).get("value"))
    except Exception as e:
        print(n, "failed", e)
PY
timeout 500 python tools/go2_tail.py 4096 4 > $O/go2_tail.txt 2>&1; grep -E "^launch|env " $O/go2_tail.txt | head -20
