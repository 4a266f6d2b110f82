#!/bin/bash
# Round-2 ncu captures of ONE eager syn20m training step (run under gpurun AFTER the same bench command has exited 0 without
# ncu; reports stay in /tmp, CSV exports go to gpurun_out/). Summarised into profiles/ by scripts/summarise_r02.py.
set -u
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --eager"
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct,lts__throughput.avg.pct_of_peak_sustained_elapsed,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size
$B > gpurun_out/r02_plain.log 2>&1 || { echo "plain run failed"; exit 1; }
# (1) every launch with its device time
DG_PROFILE_RANGE=1 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_syn20m.csv $B > gpurun_out/r02_ncu_launches.log 2>&1; echo "launch list rc=$?"
# (2) memory / tensor metrics of every own kernel of the step (DRAM traffic per launch -> roofline.traffic)
DG_PROFILE_RANGE=1 ncu --profile-from-start off --metrics $M --clock-control none -k regex:"spmm_csr|decoder_|gemm_nt|colsum_partial|splitk_reduce|center_normalize|compact_write|act_dropout|attention_" --csv --log-file gpurun_out/r02_own_kernels_metrics.csv $B > gpurun_out/r02_ncu_metrics.log 2>&1; echo "metrics rc=$?"
# (3) --set full of the dominant SpMM class (GCMC layer 0 forward, d=344) and of the decoder kernels
DG_PROFILE_RANGE=1 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:spmm_csr -c 2 -o /tmp/full_spmm $B > gpurun_out/r02_ncu_full_spmm.log 2>&1
ncu -i /tmp/full_spmm.ncu-rep --page raw --csv > gpurun_out/r02_spmm_d344_full_raw.csv 2>/dev/null
DG_PROFILE_RANGE=1 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"decoder_.*tc" -c 2 -o /tmp/full_dec $B > gpurun_out/r02_ncu_full_dec.log 2>&1
ncu -i /tmp/full_dec.ncu-rep --page raw --csv > gpurun_out/r02_decoder_full_raw.csv 2>/dev/null
ls -la gpurun_out/ | tail -n 8
