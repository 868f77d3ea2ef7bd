#!/bin/bash
# Round-end evidence: GPU tests, the default bench line, the reference arm, the ncu launch list of the bench command and
# one --set full capture of every kernel of one step.  Run on the GPU box:  bash tools/refresh_profiles.sh <tag>
tag=${1:-r02}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests_$tag.log 2>&1; tail -2 gpurun_out/gpu_tests_$tag.log
timeout 600 python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; tail -c 600 gpurun_out/bench_$tag.json
timeout 600 python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline --no-e2e > gpurun_out/plain_$tag.log 2>&1 || exit 1
timeout 900 ncu -k 'regex:pcb::|pcb_|resample|refocus|contrast|minmax|edge_cumsum' --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_$tag.csv \
    python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline --no-e2e > gpurun_out/ncu_l_$tag.log 2>&1
timeout 300 python tools/profile_stage.py --steps 3 > gpurun_out/plain_prof_$tag.log 2>&1 || exit 1
timeout 900 ncu --set full --clock-control none --import-source on \
    -k 'regex:resample_pipe|refocus_pxp|finalize|contrast_|edge_cumsum' -s 16 -c 8 -f -o gpurun_out/prof_step_$tag \
    python tools/profile_stage.py --steps 3 > gpurun_out/ncu_s_$tag.log 2>&1; tail -2 gpurun_out/ncu_s_$tag.log
ls -la gpurun_out/prof_step_$tag.ncu-rep
