#!/bin/bash
mkdir -p gpurun_out
run() { local name=$1; shift
  python bench.py --steps 3 --warmup 1 --no-secondary --no-cpu-baseline --tol-sweep "0.35,0.42,0.45" --e2e-steps 1 --parity-rows 32 --opt max_edges=600000000 "$@" > gpurun_out/exp3_$name.json 2> gpurun_out/exp3_$name.err
  python - "$name" <<'PY'
import json,sys
try:
    d=json.load(open('gpurun_out/exp3_%s.json' % sys.argv[1]))
    print(sys.argv[1], 'ms/step', round(d['ms_per_step'],2), 'kernel', round(d['roofline']['kernel_ms_per_launch'],2), 'frac', round(d['roofline']['frac'],3), 'edges', d['config']['edges'], 'parity', d.get('parity_sample',{}).get('ok'), [(t['tolerance'], round(t['kernel_ms'],1), t['edges']) for t in d.get('tol_sweep',[])])
except Exception as e:
    print(sys.argv[1], 'failed', e)
PY
}
run fold
run nofold --opt tc_fold=0
python -m pytest tests/test_gpu_search.py -q -x 2>&1 | tail -5
