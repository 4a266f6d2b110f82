#!/bin/bash
# launch list of one eager step with the captured iteration's edge sampler + sampler A/B on the small shapes
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --eager > /dev/null 2>&1; echo "eager rc=$?"
DG_PROFILE_RANGE=1 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_final_b.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --eager > gpurun_out/ncu_launches_b.log 2>&1; echo "ncu rc=$?"
for w in lrssl gdataset; do for m in randperm select; do
  DG_EDGE_SAMPLER=$m python bench.py --workload $w --steps 200 --warmup 10 --no-cpu-baseline 2>/dev/null > gpurun_out/ab_${w}_${m}.json
  python -c "import json; d=json.load(open('gpurun_out/ab_${w}_${m}.json')); print('$w $m', d['ms_per_step'], d['iters_per_sec'])"
done; done
