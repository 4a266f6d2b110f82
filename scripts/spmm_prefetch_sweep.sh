#!/bin/bash
# A/B of the SpMM L2-prefetch switch inside the real training step (per-launch CUDA-event times of bench.py)
run() {
  name=$1; shift
  env "$@" python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/b_$name.json 2> gpurun_out/b_sweep.err
  python - "$name" <<'PY'
import json,sys
n=sys.argv[1]
d=json.load(open("gpurun_out/b_%s.json"%n))
print(n, d["ms_per_step"])
for l in d["spmm_launches"]:
    if l["ms"] > 0.3: print("   ", l)
PY
}
run pf_on
run pf_off DG_SPMM_PREFETCH_MIN_MB=100000000 DG_SPMM_PREFETCH_MIN_MB_WIDE=100000000
