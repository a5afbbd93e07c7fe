#!/bin/bash
# Times the registration evaluation for each tuning variant built by voxgraph_b200/build.py
# (VGX_LIB selects the shared library). Run on the GPU box.
for f in voxgraph_b200/variants/libvgx_*.so; do
  VGX_LIB=$PWD/$f python bench.py --steps 30 --warmup 3 --no-extras 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$f'.split('libvgx_')[1], 'value %.2f G/s' % (d['value']/1e9), 'step %.1f us' % (d['ms_per_step']*1e3), 'kernel %.1f us' % (d['roofline']['kernel_ms']*1e3))"
done
