# gather-pattern microbenchmark: timings (CUDA events) and ncu LSU wavefront / L2 counters per mode
set -x
M=scripts/microbench/gather_patterns
for C in 32 64; do
  $M $C 2 5 256
done
$M 32 1 8 256
$M 32 2 5 2048
ncu --metrics gpu__time_duration.sum,l1tex__data_pipe_lsu_wavefronts_mem_lgds.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,lts__t_bytes.sum,l1tex__t_sector_hit_rate.pct,smsp__inst_executed.sum --clock-control none --csv --log-file gpurun_out/r2a_gather_patterns_ncu.csv $M 32 2 5 256 -1 8000 > gpurun_out/r2a_gather_ncu.log 2>&1
