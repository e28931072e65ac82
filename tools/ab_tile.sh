# A/B runs of the tile scoring kernel on the 200-structure ensemble (K2 time per step); usage: bash tools/ab_tile.sh "VAR=val" ...
for cfg in "$@"; do
  env $cfg timeout 300 python bench.py --ensemble 200 --steps 3 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ab_tmp.json 2> gpurun_out/ab_tmp.err
  python -c "
import json;d=json.load(open('gpurun_out/ab_tmp.json'));print('$cfg', round(d['value']/1e6,1), 'M pairs/s, K2 ms', round(d['kernels']['score']['ms_per_step'],2))"
done
