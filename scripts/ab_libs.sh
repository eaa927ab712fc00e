#!/bin/bash
# A/B of an experimental build of libunpp.so (UNPP_LIB override): training bench lines + the GPU tests on the experimental library.
L=/root/repo/unet_nested4tiny_objects_keypoints_b200
for v in "X=1" "UNPP_LIB=$L/libunpp_mw4.so" "X=2" "UNPP_LIB=$L/libunpp_mw4.so"; do
  env $v timeout 300 python bench.py --workload train --no-extra 2>/tmp/err.txt | python -c "
import sys,json
t=sys.stdin.read().strip().splitlines()
if not t: print('$v'[-14:], 'FAILED'); sys.exit(0)
d=json.loads(t[-1]); print('$v'[-14:], 'train', d['value'], d['ms_per_step'], d['loss'])"
  tail -2 /tmp/err.txt | cut -c1-300
done
UNPP_LIB=$L/libunpp_mw4.so python -m pytest tests -m gpu -q -x 2>&1 | tail -4
