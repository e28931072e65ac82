# Sliced unit order of the tile kernel (LOCOHD_TILE_SLICE): parity, A/B on the 500-structure ensemble, DRAM traffic.
# One gpurun call; every step has its own timeout and writes into gpurun_out/ as it goes.
TAG=r8c
timeout 60 python -m pytest tests/test_gpu_tile_kernel.py -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/${TAG}_tile_tests.log
ab() {  # label, env assignments
  env $2 timeout 60 python bench.py --ensemble 500 --steps 4 --warmup 2 --no-cpu-baseline --no-extras > gpurun_out/${TAG}_tmp.json 2> gpurun_out/${TAG}_tmp.err
  python -c "
import json;d=json.load(open('gpurun_out/${TAG}_tmp.json'));print('$1', round(d['value']/1e6,1), 'M pairs/s, K2t ms', round(d['kernels']['score']['ms_per_step'],2), 'e2e', round(d['e2e']['value']/1e6,1))" | tee -a gpurun_out/${TAG}_ab.txt
}
ab "slice=0(tile-major)" LOCOHD_TILE_SLICE=0
ab "slice=auto" LOCOHD_TILE_SLICE_MB=32
ab "slice=16" LOCOHD_TILE_SLICE=16
ab "slice=0(tile-major)" LOCOHD_TILE_SLICE=0
ab "slice=auto" LOCOHD_TILE_SLICE_MB=32
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct
for v in 0 16; do
  [ $SECONDS -gt 95 ] && break
  LOCOHD_TILE_SLICE=$v timeout 60 ncu --metrics $M --clock-control none -k regex:score_tile -s 1 -c 1 --csv --log-file gpurun_out/${TAG}_dram_slice$v.csv python bench.py --ensemble 500 --steps 1 --warmup 1 --no-cpu-baseline --no-extras > /dev/null 2>&1
  grep -E "dram__bytes|gpu__time|lts__t" gpurun_out/${TAG}_dram_slice$v.csv | awk -F'","' -v v=$v '{print "slice=" v, $(NF-2), $(NF-1), $NF}' | tee -a gpurun_out/${TAG}_ab.txt
done
echo "elapsed $SECONDS s" | tee -a gpurun_out/${TAG}_ab.txt
