#!/bin/bash
# usage: scratch/sweep.sh "<YH_TIME_ONLY>" name...   times scratch/lib_<name>.so with scratch/time_configs.py
only=$1; shift
for v in "$@"; do
  echo "== $v"
  YH_LIB_PATH=$PWD/scratch/lib_$v.so YH_TIME_ONLY="$only" python scratch/time_configs.py 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l.rstrip()[:200]); continue
    print(d['config'][:10], 'train', d['train_us'], 'post', d['post_us'], 'fused', d['fused_us'])
"
done
