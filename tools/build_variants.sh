#!/bin/bash
# Builds attention-kernel variants of the library in the dev container (free CPU time) into tools/variants/, from where
# they travel to the GPU box with the snapshot.  tools/build_variants.sh name "flags" [name "flags" ...]
mkdir -p tools/variants
while [ $# -ge 2 ]; do
  AACLIP_LIB_PATH=$PWD/tools/variants/$1.so AACLIP_NVCC_EXTRA="$2" python -c "
import sys; sys.path.insert(0, '.')
import aaclip_b200
from aaclip_b200 import _lib
_lib.build(force=True)" && echo "built $1 [$2]"
  shift 2
done
