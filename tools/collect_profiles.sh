#!/bin/bash
# on the GPU box (one gpurun call): bench lines of both arms, the ncu launch list of the bench command, one full capture of
# a steady-state iteration, the committed loop per iteration, the end-to-end step per stage and per kernel, the fused k_sx
# variant; everything lands in gpurun_out/ for tools/make_profiles.py
cd "$(dirname "$0")/.."
python bench.py --steps 20 --warmup 5 > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err || exit 1
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/final_bench_ref.json 2> gpurun_out/final_bench_ref.err
python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-loop > gpurun_out/plain_bench.log 2>&1 &&
GTF_GRAPH=0 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/final_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-loop > gpurun_out/ncu_launches.log 2>&1
GTF_GRAPH=0 python tools/prof_iter.py 128 > gpurun_out/plain_prof.log 2>&1 &&
GTF_GRAPH=0 ncu --set full --clock-control none --import-source on -k regex:"k_begin|k_send|k_exec|k_node2|k_hv|k_big" -s 16 -c 9 \
    -o gpurun_out/final_full python tools/prof_iter.py 128 > gpurun_out/ncu_full.log 2>&1
python tools/loop_profile.py 128 10 > gpurun_out/final_loop_profile.jsonl 2> gpurun_out/final_loop_profile.err
python tools/e2e_breakdown.py > gpurun_out/final_e2e_breakdown.log 2>&1 &&
GTF_GRAPH=0 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/final_e2e_launches.csv \
    python tools/e2e_breakdown.py > gpurun_out/ncu_e2e.log 2>&1
GTF_FUSED_SX=1 python tools/prof_iter.py 128 > gpurun_out/final_sx_prof.log 2>&1 &&
GTF_FUSED_SX=1 GTF_GRAPH=0 ncu --set full --clock-control none --import-source on -k regex:"k_sx" -s 4 -c 1 \
    -o gpurun_out/final_sx python tools/prof_iter.py 128 > gpurun_out/ncu_sx.log 2>&1
python tools/sweep.py > gpurun_out/final_sweep.jsonl 2> gpurun_out/final_sweep.err
tail -2 gpurun_out/ncu_full.log; wc -l gpurun_out/final_launches.csv gpurun_out/final_sweep.jsonl gpurun_out/final_e2e_launches.csv
