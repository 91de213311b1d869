#!/bin/bash
# Evidence for profiles/ (one B200): tools/collect_r03.sh <tag>
tag=${1:-r03}; out=gpurun_out
BENCH="python bench.py --no-cpu-baseline --no-e2e --no-extras"
KREG='mpc_setup_kernel|mpc_flow_kernel|mpc_lq_kernel|mpc_riccati_kernel|mpc_linesearch_kernel|wbc_update_kernel'
# (1) full capture of every kernel of one tick at 1024 robots (second tick: warm start), source-level
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"$KREG" --launch-skip 6 -c 6 -o $out/prof_$tag -f $BENCH --batch 1024 --steps 1 --warmup 1 > $out/ncu_$tag.log 2>&1
# (2) tensor-pipe (DMMA) / fp64-pipe activity and counted DMMA instructions of the Riccati kernel
timeout 600 ncu --clock-control none -k regex:mpc_riccati_kernel --launch-skip 1 -c 1 --csv --metrics sm__inst_executed_pipe_tensor_subpipe_dmma.sum,smsp__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fp64.sum,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,gpu__time_duration.sum,sm__cycles_active.avg $BENCH --batch 1024 --steps 1 --warmup 1 > $out/dmma_$tag.csv 2>$out/dmma_$tag.err
# (3) launch list of the bench command itself (shares of the tick)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_$tag.csv $BENCH --steps 2 --warmup 1 > $out/launches_$tag.log 2>&1
# (4) compute-sanitizer: memcheck, racecheck, synccheck on every kernel of the library (tiny batches)
for tool in memcheck racecheck synccheck; do timeout 900 compute-sanitizer --tool $tool python tools/sanitize_small.py > $out/san_${tool}_$tag.log 2>&1; tail -3 $out/san_${tool}_$tag.log; done
ls -la $out/prof_$tag.ncu-rep
