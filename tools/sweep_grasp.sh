for R in 32 72 96; do for W in 4 7; do
  GSR_SMEM_ROWS=$R GSR_WARPS=$W timeout 300 python bench.py --workload grasp --steps 300 --warmup 100 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('grasp rows=$R warps=$W', round(d['value']/1e6,2),'M', round(d['ms_per_step'],3),'ms', d['config']['launch'], 'contacts', d['config']['mean_contacts_per_env'])"
done; done
