#!/bin/bash
# ncu captures behind profiles/ (run on the GPU box through gpurun; reports land in gpurun_out/).
#   bash tools/capture_profiles.sh r2
# Numbers printed by a run under ncu are never bench values.
set -u
tag="${1:-r2}"
out=gpurun_out
full="ncu --set full --clock-control none --import-source on -f"
# launch list of the bench command (kernel shares of the timed loops)
# (--batches 2: two batches instead of 78, so that the list reaches the timed loops)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv \
    --log-file $out/${tag}_launches_default.csv \
    python bench.py --steps 20 --warmup 3 --batches 2 --no-sweep --no-cpu-baseline --no-configs > $out/${tag}_launches.log 2>&1
# the step kernels, one launch each (after warm-up launches)
timeout 300 $full -k regex:k_env_step_warp -s 6 -c 1 -o $out/${tag}_step_warp_4096 python tools/prof_step.py > /dev/null 2>&1
timeout 300 $full -k regex:k_env_step_warp -s 6 -c 1 -o $out/${tag}_step_obs_warp_4096 python tools/prof_step.py --obs > /dev/null 2>&1
timeout 300 $full -k regex:k_env_step_pipe -s 4 -c 1 -o $out/${tag}_step_pipe_65536 python tools/prof_step.py --envs 65536 --steps 8 > /dev/null 2>&1
timeout 300 $full -k regex:k_env_step_warp -s 4 -c 1 -o $out/${tag}_step_obs_warp_65536 python tools/prof_step.py --envs 65536 --steps 8 --obs > /dev/null 2>&1
timeout 300 $full -k regex:k_env_step_warp -s 6 -c 1 -o $out/${tag}_step_warp_spawn_8192 python tools/prof_step.py --pool append-spawn-25 --envs 8192 > /dev/null 2>&1
timeout 300 $full -k regex:k_env_step_tile -s 6 -c 1 -o $out/${tag}_step_tile_nav50_4096 python tools/prof_step.py --pool navigation-50 > /dev/null 2>&1
timeout 300 $full -k regex:k_env_step_tile -s 6 -c 1 -o $out/${tag}_step_obs_tile_nav50_4096 python tools/prof_step.py --pool navigation-50 --obs > /dev/null 2>&1
timeout 300 $full -k regex:k_get_obs_warp -s 0 -c 1 -o $out/${tag}_get_obs_4096 python tools/prof_step.py > /dev/null 2>&1
timeout 300 $full -k regex:k_multistep_tile -s 1 -c 1 -o $out/${tag}_multistep_tile_nav50_256 python tools/bench_side_effects.py navigation-50 256 300 > /dev/null 2>&1
# summaries and traffic on the box (the reports themselves are too large to travel back together)
sum() { python tools/ncu_report.py $out/${tag}_$1.ncu-rep $2 > $out/${tag}_$1_ncu.txt 2>&1; }
sum step_warp_4096 4096; sum step_obs_warp_4096 4096; sum step_pipe_65536 65536; sum step_obs_warp_65536 65536
sum step_warp_spawn_8192 8192
sum step_tile_nav50_4096 4096; sum step_obs_tile_nav50_4096 4096; sum get_obs_4096 4096; sum multistep_tile_nav50_256 1
tr() { python tools/ncu_traffic.py $out/${tag}_$1.ncu-rep "$2" $3 > /dev/null 2>&1; }
tr step_warp_4096 "k_env_step_warp<Narrow, goals in place>" 4096x25x25
tr step_obs_warp_4096 "k_env_step_warp<Narrow, goals resident, obs fused>" 4096x25x25
tr step_pipe_65536 "k_env_step_pipe<5 stages, 3 CTAs/SM, goals in place>" 65536x25x25
tr step_obs_warp_65536 "k_env_step_warp<Narrow, goals resident, obs fused>" 65536x25x25
tr step_warp_spawn_8192 "k_env_step_warp<Narrow, goals in place>" 8192x25x25
tr step_tile_nav50_4096 "k_env_step_tile<2 x 2 tiles, 2 envs/CTA>" 4096x50x50
tr step_obs_tile_nav50_4096 "k_env_step_tile<2 x 2 tiles, 2 envs/CTA, obs fused>" 4096x50x50
cp profiles/r2_traffic.json $out/r2_traffic.json
python tools/sass_histogram.py > $out/${tag}_sass_histogram.txt 2>&1
rm -f $out/${tag}_*.ncu-rep
ls -la $out/${tag}_*
