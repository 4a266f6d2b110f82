#!/bin/bash
# Final round-2 captures (run under gpurun; CSV exports and bench lines go to gpurun_out/, summarised into profiles/ by
# scripts/summarise_r02b.py). Every ncu pass runs only after the same command has exited 0 without ncu.
set -u
B="python bench.py --no-cpu-baseline --no-extra"
python bench.py --steps 20 --warmup 5 > gpurun_out/r02b_bench_full.json 2> gpurun_out/r02b_bench_full.err || { echo "bench failed"; tail -5 gpurun_out/r02b_bench_full.err; exit 1; }
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02b_bench_reference.json 2> gpurun_out/r02b_bench_reference.err; echo "reference arm rc=$?"
for w in syn20m lrssl; do
  $B --workload $w --steps 1 --warmup 3 --eager > gpurun_out/r02b_plain_$w.log 2>&1 || { echo "plain $w failed"; continue; }
  DG_PROFILE_RANGE=1 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/r02b_launches_$w.csv $B --workload $w --steps 1 --warmup 3 --eager > gpurun_out/r02b_ncu_$w.log 2>&1; echo "launch list $w rc=$?"
done
for w in lrssl gdataset cdataset; do
  $B --workload $w --steps 200 --warmup 10 > gpurun_out/r02b_bench_$w.json 2> gpurun_out/r02b_bench_$w.err; echo "bench $w rc=$?"
done
ls -la gpurun_out/ | grep r02b
