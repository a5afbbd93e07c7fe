#!/bin/bash
# Times the registration evaluation of the default library (and of every variant library) under
# several runtime settings (VGX_REG_HW_TILE_UNITS = tile size of the hardware-scheduled reduce
# kernel; 0 = persistent CTAs).
for lib in "" voxgraph_b200/variants/libvgx_*.so; do
for u in ${TUNE_UNITS:-0 32 64}; do
  if [ -n "$lib" ] && [ ! -f "$lib" ]; then continue; fi
  L=""; if [ -n "$lib" ]; then L=$PWD/$lib; fi
  VGX_LIB=$L VGX_REG_HW_TILE_UNITS=$u python bench.py --steps 30 --warmup 3 --no-extras 2>gpurun_out/tune_err.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('${lib:-default}', 'hw_tile_units=$u', 'value %.2f G/s' % (d['value']/1e9), 'step %.1f us' % (d['ms_per_step']*1e3), 'kernel %.1f us' % (d['roofline']['kernel_ms']*1e3))"
done
done
