#!/bin/bash
# A/B of two builds of the library on the hashing step (same box, alternating): $1 = other .so
for rep in 1 2; do
  for so in "$1" vid_dup_finder_lib_b200/libvdf_b200.so; do
    VDF_B200_SO=$PWD/$so python bench.py --workload hash --steps 20 --warmup 3 --hash-total 5120 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; n=d['steps']
print('$so', 'ms/step %.4f' % d['ms_per_step'], 'kernel %.4f' % r['kernel_ms_per_launch'], d['clocks']['sm_mhz'], d['clocks']['reasons'])"
  done
done
