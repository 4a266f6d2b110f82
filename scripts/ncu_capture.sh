#!/bin/bash
# ncu captures of one eager syn20m training step (run under gpurun; reports stay in /tmp, CSV exports go to gpurun_out/)
set -u
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct,lts__throughput.avg.pct_of_peak_sustained_elapsed,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --eager"
# (1) targeted memory metrics for every own kernel of the step
DG_PROFILE_RANGE=1 ncu --profile-from-start off --metrics $M --clock-control none -k regex:"spmm_csr|decoder_|gemm_nt|colsum_partial|splitk_reduce|center_normalize|compact_write" --csv --log-file gpurun_out/r01b_own_kernels_metrics.csv $B > gpurun_out/ncu_metrics.log 2>&1
# (2) --set full of the dominant SpMM class (first two launches: GCMC layer 0 forward, d=344) and the tcgen05 decoder kernels
DG_PROFILE_RANGE=1 ncu --profile-from-start off --set full --clock-control none -k regex:spmm_csr -c 2 -o /tmp/full_spmm $B > gpurun_out/ncu_full_spmm.log 2>&1
ncu -i /tmp/full_spmm.ncu-rep --page raw --csv > gpurun_out/r01b_spmm_d344_full_raw.csv 2>/dev/null
DG_PROFILE_RANGE=1 ncu --profile-from-start off --set full --clock-control none -k regex:"decoder_.*tc" -c 2 -o /tmp/full_dec $B > gpurun_out/ncu_full_dec.log 2>&1
ncu -i /tmp/full_dec.ncu-rep --page raw --csv > gpurun_out/r01b_decoder_full_raw.csv 2>/dev/null
ls -la gpurun_out/ | tail -n 8
