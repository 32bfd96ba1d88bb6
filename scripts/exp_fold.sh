#!/bin/bash
# one GPU call: which part of the fold kernel costs time?  (bench lines + one ncu capture of the kernel at 262 k hashes)
mkdir -p gpurun_out
for o in "tc_fold=0" "tc_fold=-1"; do
  python bench.py --steps 3 --warmup 1 --no-secondary --no-cpu-baseline --tol-sweep "" --e2e-steps 1 --opt $o > gpurun_out/exp_${o//=/_}.json 2> gpurun_out/exp_${o//=/_}.err
  python - "$o" <<'PY'
import json,sys
d=json.load(open('gpurun_out/exp_%s.json' % sys.argv[1].replace('=','_')))
print(sys.argv[1], 'ms/step', round(d['ms_per_step'],2), 'kernel', round(d['roofline']['kernel_ms_per_launch'],2), 'frac', round(d['roofline']['frac'],3), 'e2e', round(d['e2e']['ms_per_call'],1))
PY
done
ncu --set full --clock-control none --import-source on -k regex:hamming_tc6 -c 1 -o gpurun_out/r02_tc6_fold python bench.py --steps 1 --warmup 0 --n 262144 --no-secondary --no-cpu-baseline --tol-sweep "" --e2e-steps 1 > /dev/null 2> gpurun_out/ncu_fold.err
ncu -i gpurun_out/r02_tc6_fold.ncu-rep --page raw --csv > gpurun_out/r02_tc6_fold_raw.csv 2>/dev/null
python -m pytest tests -m gpu -q -x 2>&1 | tail -15
