# Evidence capture for the tile scoring kernel (run under gpurun on one B200): bash tools/r6_evidence.sh <tag>; outputs under gpurun_out/<tag>_*
set -x
TAG=${1:-r6f}
compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_tile_kernel.py -m gpu -x -q > gpurun_out/${TAG}_memcheck.log 2>&1; tail -3 gpurun_out/${TAG}_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_tile_kernel.py -m gpu -x -q -k "blocked_order or small_h2" > gpurun_out/${TAG}_racecheck.log 2>&1; tail -3 gpurun_out/${TAG}_racecheck.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"score_tile" -s 3 -c 1 -f -o gpurun_out/${TAG}_full python bench.py --workload cfg5 --ensemble 48 --steps 1 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/${TAG}_full_bench.log 2>&1
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:"score_tile|env_fused|build_cells|env_tile" -s 12 -c 4 --csv --log-file gpurun_out/${TAG}_cfg5_metrics.csv python bench.py --steps 1 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/${TAG}_cfg5_metrics_bench.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/${TAG}_launches_bench.log 2>&1
ls -la gpurun_out/${TAG}_*
