#!/bin/bash
# A/B of library builds on one box: tools/ab_bench.sh <tag> [lib ...]   (default: the in-tree library); writes gpurun_out/ab_<tag>.jsonl (one bench line per build)
tag=$1; shift
out=gpurun_out/ab_$tag.jsonl; : > $out
libs="$@"; [ -z "$libs" ] && libs=qm_control_b200/libqmb200.so
for lib in $libs; do
  echo "{\"lib\": \"$lib\"}" >> $out
  QMB200_LIB=$PWD/$lib timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-extras >> $out 2>> gpurun_out/ab_$tag.err
done
python - "$out" <<'PY'
import json, sys
lib = None
for line in open(sys.argv[1]):
    d = json.loads(line)
    if "lib" in d and len(d) == 1: lib = d["lib"]; continue
    k = d["roofline"]["kernel_ms"]; print("%-45s tick %.2f ms  %s" % (lib, d["ms_per_step"], {a: round(b, 2) for a, b in k.items()}))
PY
