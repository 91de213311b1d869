#!/bin/bash
# One B200: profiles evidence, then the bench lines that quote it.  tools/final_r03.sh <tag>
tag=${1:-r03}; out=gpurun_out
tools/collect_r03.sh $tag
python tools/ncu_kernels.py $out/prof_$tag.ncu-rep --batch 1024 --json $out/traffic_r03.json > $out/${tag}_kernel_summary.txt 2>&1
cp $out/traffic_r03.json profiles/traffic_r03.json
python tools/dmma_json.py $out/dmma_$tag.csv 1024 100 $out/r03_riccati_dmma.json && cp $out/r03_riccati_dmma.json profiles/r03_riccati_dmma.json && sed -i "s#profiles/dmma_$tag.csv#profiles/${tag}_riccati_dmma.csv#" profiles/r03_riccati_dmma.json $out/r03_riccati_dmma.json
for k in mpc_flow_kernel mpc_lq_kernel mpc_riccati_kernel mpc_linesearch_kernel wbc_update_kernel; do NCU_KERNEL=$k python tools/ncu_hot.py $out/prof_$tag.ncu-rep 40 > $out/${tag}_${k}_hotlines.txt 2>&1; done
timeout 600 python bench.py > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $out/${tag}_bench_ref.json 2> $out/${tag}_bench_ref.err
for w in mpc wbc mixed; do timeout 300 python bench.py --workload $w --no-cpu-baseline > $out/${tag}_bench_$w.json 2> $out/${tag}_bench_$w.err; done
timeout 300 python bench.py --solver ddp --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > $out/${tag}_bench_ddp.json 2> $out/${tag}_bench_ddp.err
cuobjdump -sass qm_control_b200/libqmb200.so 2>/dev/null | grep -E "UBLKCP|SYNCS|DMMA|UTMA" | awk '{print $2}' | sort | uniq -c | sort -rn | head -20 > $out/${tag}_sass_mnemonics.txt
tail -c 600 $out/${tag}_bench_n1.json; echo; cat $out/${tag}_sass_mnemonics.txt | head -8
