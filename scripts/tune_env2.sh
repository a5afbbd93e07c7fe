#!/bin/bash
# Times the default library's registration evaluation under combinations of runtime switches.
#   VGX_REG_HW_TILE_UNITS  tile size of the one-CTA-per-tile reduce kernel (0 = persistent CTAs)
#   VGX_REG_LPT            0 = tiles run in index order, 1 = longest-first from measured cost
#   VGX_NO_PDL             1 = no programmatic dependent launches
run() {
  env "$@" python bench.py --steps 30 --warmup 5 --no-extras 2>gpurun_out/tune_err.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$*', 'value %.2f G/s' % (d['value']/1e9), 'step %.1f us' % (d['ms_per_step']*1e3), 'kernel %.1f us' % (d['roofline']['kernel_ms']*1e3))"
}
for u in ${TUNE_UNITS:-32 48 64 96}; do
  for lpt in 0 1; do
    run VGX_REG_HW_TILE_UNITS=$u VGX_REG_LPT=$lpt
  done
done
run VGX_REG_HW_TILE_UNITS=32 VGX_REG_LPT=1 VGX_NO_PDL=1
run VGX_REG_HW_TILE_UNITS=0
