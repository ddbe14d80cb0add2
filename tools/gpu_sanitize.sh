#!/bin/bash
# compute-sanitizer record of the lane-group kernels on the final binary (memcheck, racecheck, synccheck)
O=gpurun_out/${1:-sanitize}
mkdir -p $O
for T in memcheck racecheck synccheck; do
  echo "== $T" >> $O/compute_sanitizer.txt
  timeout 1500 compute-sanitizer --tool $T python tools/sanitize.py group 2>&1 | grep -E "sanitize:|ERROR SUMMARY|RACECHECK SUMMARY|hazard|Error|error" | sort | uniq -c | sort -rn | head -40 >> $O/compute_sanitizer.txt
done
cat $O/compute_sanitizer.txt
