#!/bin/bash
# A/B of experimental builds of libunpp.so (UNPP_LIB override): bench lines of inference and training for each library.
L=/root/repo/unet_nested4tiny_objects_keypoints_b200
for v in "X=1" "UNPP_LIB=$L/libunpp_backoff.so" "UNPP_LIB=$L/libunpp_mw4.so"; do
  env $v timeout 300 python bench.py --no-extra 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v'[-14:], d['value'], d['ms_per_step'])
for k,v in sorted(d['per_kernel'].items()):
    if 'head' in k or 'K64 N32' in k or 'K96' in k: print('  ',k,v['ms'])
"
  env $v timeout 300 python bench.py --workload train --no-extra 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('   train', d['value'], d['ms_per_step'], d['loss'])"
done
UNPP_LIB=$L/libunpp_mw4.so python -m pytest tests -m gpu -q -x 2>&1 | tail -2
