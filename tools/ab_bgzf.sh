#!/bin/bash
# A/B of k_bgzf_inflate build variants on the GPU box: tools/ab_bgzf.sh "-DLPS_BGZF_RING=4096" "-DLPS_BGZF_RING=8192" ...
cd "$(dirname "$0")/.."
for cfg in "$@"; do
  touch longphase-s_b200/csrc/k_bgzf.cu
  make -C longphase-s_b200/csrc EXTRA="$cfg" > /dev/null 2>&1 || { echo "build failed: $cfg"; continue; }
  python tools/bgzf_prof.py ${AB_MB:-1024} 2>/tmp/ab_bgzf.err | python -c "
import json,sys
d=json.loads(sys.stdin.readline())
print('$cfg', 'kernel_ms=%.2f plain_ms=%.2f out_GB/s=%.1f e2e_GB/s=%.1f' % (d['kernel_ms'], d['kernel_ms_plain_decoder'], d['out_gb_per_s'], d['e2e_out_gb_per_s']))" || tail -3 /tmp/ab_bgzf.err
done
touch longphase-s_b200/csrc/k_bgzf.cu; make -C longphase-s_b200/csrc > /dev/null 2>&1
