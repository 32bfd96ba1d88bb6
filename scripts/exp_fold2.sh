#!/bin/bash
mkdir -p gpurun_out
run() { # name, so, opts...
  local name=$1 so=$2; shift 2
  VDF_B200_SO=$so python bench.py --steps 3 --warmup 1 --no-secondary --no-cpu-baseline --tol-sweep "" --e2e-steps 1 --parity-rows 0 "$@" > gpurun_out/exp2_$name.json 2> gpurun_out/exp2_$name.err
  python - "$name" <<'PY'
import json,sys
try:
    d=json.load(open('gpurun_out/exp2_%s.json' % sys.argv[1]))
    print(sys.argv[1], 'ms/step', round(d['ms_per_step'],2), 'kernel', round(d['roofline']['kernel_ms_per_launch'],2), 'frac', round(d['roofline']['frac'],3), 'edges', d['config']['edges'])
except Exception as e:
    print(sys.argv[1], 'failed', e)
PY
}
E=$PWD/vid_dup_finder_lib_b200/exp
run base_nofold "" --opt tc_fold=0
run noagg_nofold $E/libvdf_NOAGG.so --opt tc_fold=0
run small_nofold $E/libvdf_SMALLTILE.so --opt tc_fold=0
run both_nofold $E/libvdf_NOAGGSMALLTILE.so --opt tc_fold=0
run noagg_fold $E/libvdf_NOAGG.so
run base_fold ""
