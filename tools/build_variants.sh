#!/bin/bash
# Build flag variants of the library side by side: tools/build_variants.sh tag1="-DQMB_X=1 ..." tag2="..."  ->  build_variants/<tag>/libqmb200.so (git-ignored, ships to the GPU box)
set -e
cd "$(dirname "$0")/.."
for spec in "$@"; do
  tag=${spec%%=*}; flags=${spec#*=}
  mkdir -p build_variants/$tag
  make -s -j8 -C qm_control_b200/csrc OBJDIR=$PWD/build_variants/$tag/_build OUT=$PWD/build_variants/$tag/libqmb200.so EXTRA="$flags" 2>&1 | grep -E "error|undefined" && exit 1
  echo "built build_variants/$tag/libqmb200.so ($flags)"
done
