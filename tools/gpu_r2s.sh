#!/bin/bash
# full GPU suite, then the grasp leg with the generic library (adaptive rows, capped) and the specialised copy
O=gpurun_out/${1:-r2s}
mkdir -p $O
( timeout 2400 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.log )
grep -E "^FAILED|^ERROR|passed|failed" $O/pytest_gpu.log | head -40
GSR_NO_SPEC=1 timeout 600 python bench.py --workload grasp --steps 100 --warmup 5 > $O/bench_grasp_generic.json 2> $O/bench_grasp_generic.err
timeout 600 python bench.py --workload grasp --steps 100 --warmup 5 > $O/bench_grasp.json 2> $O/bench_grasp.err
python - <<PY
import json
for n in ("bench_grasp_generic", "bench_grasp"):
    try:
        d = json.loads(open("$O/%s.json" % n).read().strip().splitlines()[-1])
        print(n, "value %.3gM e2e %.3gM" % (d["value"] / 1e6, d["e2e"]["value"] / 1e6), d.get("config", {}).get("launch"))
    except Exception as e:
        print(n, "failed", e)
PY
