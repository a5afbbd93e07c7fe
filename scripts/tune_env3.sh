#!/bin/bash
run() {
  env "$@" python bench.py --steps 30 --warmup 5 --no-extras 2>gpurun_out/tune_err.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$*', 'value %.2f G/s' % (d['value']/1e9), 'step %.1f us' % (d['ms_per_step']*1e3), 'kernel %.1f us' % (d['roofline']['kernel_ms']*1e3), 'e2e %.2f G/s' % (d['e2e']['value']/1e9))"
}
run VGX_X=default
run VGX_REG_HW_TILE_UNITS=96
run VGX_REG_HW_TILE_UNITS=128
run VGX_NO_PDL=1
