#!/bin/bash
for rep in 1 2; do
  for so in "$1" vid_dup_finder_lib_b200/libvdf_b200.so; do
    VDF_B200_SO=$PWD/$so python bench.py --workload hash --steps 391 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('$so', 'sustained %.4f' % d['ms_per_step'], 'kernel %.4f' % r['kernel_ms_per_launch'], 'burst %.4f' % d['burst']['ms_per_step'], d['clocks']['sm_mhz'], d['result_digest'])"
  done
done
