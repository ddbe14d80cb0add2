#!/bin/bash
# last session of round 2: GPU suite, smoke, the two bench arms at the driver's protocol, the host-call breakdown
O=gpurun_out/${1:-r2i}
mkdir -p $O
( timeout 1200 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.log )
grep -E "^FAILED|^ERROR|passed|failed" $O/pytest_gpu.log | head -20
( timeout 300 python __graft_entry__.py smoke 2>&1 | grep -i smoke > $O/smoke.log ); cut -c1-160 $O/smoke.log
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/bench_driver.json 2> $O/bench_driver.err
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $O/bench_reference.json 2> $O/bench_reference.err
timeout 300 python tools/e2e_breakdown.py 2>&1 | grep -v -i warn | tail -5 > $O/e2e_breakdown.txt; cat $O/e2e_breakdown.txt
python - <<PY
import json
for n in ("bench_driver", "bench_reference"):
    d = json.loads(open("$O/%s.json" % n).read().strip().splitlines()[-1])
    print(n, "value %.4gM e2e %.4gM" % (d["value"] / 1e6, d["e2e"]["value"] / 1e6), "go2:", (d.get("go2") or {}).get("value"), ((d.get("go2") or {}).get("e2e") or {}).get("value"))
PY
