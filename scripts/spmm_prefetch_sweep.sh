run() { # name, env...
  name=$1; shift
  env "$@" python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/b_$name.json 2> gpurun_out/b_sweep.err
  python - "$name" <<'PY'
import json,sys
n=sys.argv[1]
try:
    d=json.load(open("gpurun_out/b_%s.json"%n))
    print(n, d["ms_per_step"], {k:v["ms_per_step"] for k,v in d["roofline"]["all_spmm_classes"].items()})
except Exception as e:
    print(n, "failed", e)
PY
}
run th120 DG_SPMM_PREFETCH_MIN_MB=120
run th48 DG_SPMM_PREFETCH_MIN_MB=48
run th120_v1 DG_SPMM_PREFETCH_MIN_MB=120 DG_SPMM_VARIANT=1
run th120_v1_pf2 DG_SPMM_PREFETCH_MIN_MB=120 DG_SPMM_VARIANT=1 DG_SPMM_PREFETCH=2
run th120_v2 DG_SPMM_PREFETCH_MIN_MB=120 DG_SPMM_VARIANT=2
run th120_v2_pf2 DG_SPMM_PREFETCH_MIN_MB=120 DG_SPMM_VARIANT=2 DG_SPMM_PREFETCH=2
