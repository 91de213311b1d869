#!/bin/bash
# Build everything, then run a command on the B200 box: tools/gpu.sh <timeout_s> '<command>'
set -e
cd "$(dirname "$0")/.."
make -s -C oracle
make -s -j8 -C qm_control_b200/csrc 2>&1 | grep -E "error|undefined" && exit 1
T=$1; shift
exec timeout $((T + 1900)) /usr/local/graft/bin/gpurun --timeout "$T" -- "$@"
