#!/bin/bash
# usage: tools/variants.sh name1 "flags1" name2 "flags2" ...   builds variants into gpurun_in/
cd "$(dirname "$0")/.." && mkdir -p gpurun_in; rm -f gpurun_in/lib_*.so
while [ $# -gt 1 ]; do
  name=$1; flags=$2; shift 2
  python -c "
import sys; sys.path.insert(0,'.')
from gudni_b200 import _build
_build.build_cuda(force=True, extra='$flags'.split())" 2>&1 | grep -E "error" ; cp gudni_b200/libgudni_b200.so gpurun_in/lib_$name.so
done
