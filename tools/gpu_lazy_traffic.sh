#!/bin/bash
# DRAM traffic of the step kernel with the kinematic cache carried through the record (default) and recomputed on chip (GSR_LAZY_FK=1)
O=gpurun_out/${1:-lazy}
mkdir -p $O
for w in franka go2; do
  S=45; [ $w = go2 ] && S=345
  for e in 0 1; do
    GSR_LAZY_FK=$e timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k regex:k_step -s $S -c 2 --csv \
      --log-file $O/traffic_${w}_lazy$e.csv python bench.py --workload $w --steps $((S + 15)) --warmup 5 --no-cpu-baseline --no-extra-legs > /dev/null 2>&1
  done
done
python - <<PY
import csv, glob
for f in sorted(glob.glob("$O/traffic_*.csv")):
    rows = [r for r in csv.reader(open(f)) if len(r) > 5 and "k_step" in ",".join(r)]
    tot = {}
    for r in rows:
        tot.setdefault(r[-3], []).append(float(r[-1].replace(",", "")))
    print(f.split("/")[-1], {k: sum(v) / len(v) for k, v in tot.items()}, rows[0][-2] if rows else "")
PY
