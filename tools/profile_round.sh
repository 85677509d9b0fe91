#!/bin/bash
# One GPU call that refreshes everything under profiles/: the launch list of the bench command and one
# `ncu --set full` capture per kernel (second launch of each).  usage: bash tools/profile_round.sh <tag>   (on the GPU box)
TAG=${1:-r2}
OUT=gpurun_out
NCU="ncu --set full --clock-control none --import-source on -c 1 -s 1 -f"
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --sections 12 > $OUT/bench_under_ncu.json 2> $OUT/bench_under_ncu.err
$NCU -k regex:sepconv_fwd_k51_v3 -o $OUT/fwd_$TAG python tools/run_one.py fwd 16 3 512 512 3 > /dev/null 2>&1
$NCU -k regex:sepconv_bwd_taps_k51_v3 -o $OUT/bwd_$TAG python tools/run_one.py bwd 16 3 512 512 3 > /dev/null 2>&1
$NCU -k regex:sepconv_bwd_input_k51 -o $OUT/gi_$TAG python tools/run_one.py gi 16 3 512 512 3 > /dev/null 2>&1
FLOW=fold $NCU -k regex:warp_torch_tma -o $OUT/warp_fold_$TAG python tools/run_one.py warp 1 3 4096 4096 3 > /dev/null 2>&1
FLOW=fold $NCU -k regex:warp_torch_tma -o $OUT/warp_fold2048_$TAG python tools/run_one.py warp 1 3 2048 2048 3 > /dev/null 2>&1
FLOW=noise $NCU -k regex:warp_torch_tma -o $OUT/warp_noise_$TAG python tools/run_one.py warp 1 3 2048 2048 3 > /dev/null 2>&1
GRAY=1 $NCU -k regex:interp_tail_fwd -o $OUT/tail_$TAG python tools/run_one.py tail 16 3 512 512 3 > /dev/null 2>&1
GRAY=1 $NCU -k regex:sepconv_bwd_taps_k51_kernel -o $OUT/tailbwd_$TAG python tools/run_one.py tailbwd 16 3 512 512 3 > /dev/null 2>&1
$NCU -k regex:sff_degrade -o $OUT/sff_$TAG python tools/run_one.py sff 1 1 4096 4096 3 > /dev/null 2>&1
ls -la $OUT/*_$TAG.ncu-rep $OUT/launches_$TAG.csv
