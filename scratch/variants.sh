#!/bin/bash
# usage: scratch/variants.sh name "-DFLAG ..." [name "-D..."]...   builds scratch/lib_<name>.so (experimental builds of the product library)
cd "$(dirname "$0")/.."
while [ $# -gt 0 ]; do
  name=$1; flags=$2; shift 2
  python - "$name" $flags <<'PY' &
import sys
sys.path.insert(0, '.')
from odcp_b200 import build
print(build.build(extra=sys.argv[2:], out='scratch/lib_%s.so' % sys.argv[1]))
PY
done
wait
