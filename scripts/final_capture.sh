#!/bin/bash
# Final measurements of the round (run under gpurun, 1 GPU): bench lines + ncu launch list of one step
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_final_b.json 2> gpurun_out/bench_final_b.err; echo "bench rc=$?"
for w in lrssl gdataset cdataset; do
  python bench.py --workload $w --steps 200 --warmup 10 --no-cpu-baseline > gpurun_out/bench_${w}_b.json 2>> gpurun_out/bench_final_b.err; echo "$w rc=$?"
done
DG_PROFILE_RANGE=1 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_final_b.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --eager > gpurun_out/ncu_launches_b.log 2>&1; echo "ncu rc=$?"
for f in gpurun_out/bench_final_b.json gpurun_out/bench_lrssl_b.json gpurun_out/bench_gdataset_b.json gpurun_out/bench_cdataset_b.json; do cut -c1-230 $f; echo; done
