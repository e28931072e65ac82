# A/B of two library builds on configs 2 and 4 (K2 = score_fast_kernel): bash tools/ab_cfg.sh  (expects build/old_lib.so, build/new_lib.so)
for r in 1 2; do for v in old new; do cp build/${v}_lib.so loco_hd_b200/liblocohd_b200.so
  for w in cfg2 cfg4; do timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ab_tmp.json 2>/dev/null
    python -c "
import json;d=json.load(open('gpurun_out/ab_tmp.json'));print('$v $w', round(d['value']/1e6,1), {k:round(x['ms_per_step'],3) for k,x in d['kernels'].items()})"; done; done; done
