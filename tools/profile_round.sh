#!/bin/bash
# ncu evidence for one round: launch lists of the bench command + one --set full capture of every hot kernel, summarised ON the box
# (gpurun copies back at most 64 MiB: the .ncu-rep files are dropped if they would not fit).
#   gpurun --timeout 1500 -- 'bash tools/profile_round.sh r2h'        (outputs: gpurun_out/<tag>_*)
tag=${1:-r2}
mkdir -p gpurun_out
FE='regex:k_fast_cells|k_orient_describe|k_octree|k_blur7|k_search_candidates|k_search_resolve|k_pose_optimization|k_resize_level|k_project_last|k_grid_build'
BA='regex:k_rs_solve|k_ba_schur_items|k_ba_schur_finish|k_ba_build_points|k_ba_build_poses|k_ba_point_prep|k_ba_backsub|k_ba_errors|k_ba_update'
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/${tag}_launches_fe.csv \
    python bench.py --workload frontend --steps 2 --warmup 3 > gpurun_out/${tag}_launches_fe.log 2>&1
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/${tag}_launches_ba.csv \
    python bench.py --workload ba --steps 2 --warmup 3 > gpurun_out/${tag}_launches_ba.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k "$FE" --launch-skip 64 --launch-count 18 -f -o gpurun_out/${tag}_fe \
    python bench.py --workload frontend --steps 1 --warmup 3 > gpurun_out/${tag}_full_fe.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k "$BA" --launch-skip 200 --launch-count 10 -f -o gpurun_out/${tag}_ba \
    python bench.py --workload ba --steps 1 --warmup 3 > gpurun_out/${tag}_full_ba.log 2>&1
python tools/ncu_summary.py kernel gpurun_out/${tag}_fe.ncu-rep gpurun_out/${tag}_frontend_kernels_full.txt > /dev/null
python tools/ncu_summary.py kernel gpurun_out/${tag}_ba.ncu-rep gpurun_out/${tag}_ba_kernels_full.txt > /dev/null
python tools/ncu_summary.py traffic gpurun_out/${tag}_fe.ncu-rep gpurun_out/${tag}_ba.ncu-rep gpurun_out/${tag}_traffic.json > /dev/null
python tools/ncu_summary.py launches gpurun_out/${tag}_launches_fe.csv gpurun_out/${tag}_launches_frontend.txt > /dev/null
python tools/ncu_summary.py launches gpurun_out/${tag}_launches_ba.csv gpurun_out/${tag}_launches_ba_500kf.txt > /dev/null
sz=$(du -sm gpurun_out | cut -f1)
if [ "$sz" -gt 55 ]; then rm -f gpurun_out/${tag}_fe.ncu-rep; fi
sz=$(du -sm gpurun_out | cut -f1)
if [ "$sz" -gt 55 ]; then rm -f gpurun_out/${tag}_ba.ncu-rep; fi
ls -la gpurun_out/
