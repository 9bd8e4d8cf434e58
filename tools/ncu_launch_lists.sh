#!/bin/bash
# ncu launch lists (gpu__time_duration.sum, cold-cache, serialised: shares, not absolutes) of the bench command and of the
# three fit profiles; summaries with tools/launch_summary.py. Run under gpurun; outputs in gpurun_out/.
set -u
tag=${1:-r02c}
M="--metrics gpu__time_duration.sum --clock-control none --csv"
ncu $M --log-file gpurun_out/${tag}_launch_list.csv python bench.py --quick --steps 2 --warmup 3 > gpurun_out/${tag}_bench_under_ncu.log 2>&1
ncu $M --log-file gpurun_out/${tag}_fit_launches.csv python tools/profile_fit_defaults.py 0 > /dev/null 2>&1
ncu $M --log-file gpurun_out/${tag}_fitlam_launches.csv python tools/profile_fit_defaults.py 0.05 > /dev/null 2>&1
ncu $M --log-file gpurun_out/${tag}_pose_launches.csv python tools/profile_pose.py > /dev/null 2>&1
for f in fit fitlam pose; do python tools/launch_summary.py gpurun_out/${tag}_${f}_launches.csv > gpurun_out/${tag}_${f}_launch_summary.txt; done
head -12 gpurun_out/${tag}_fit_launch_summary.txt; head -6 gpurun_out/${tag}_fitlam_launch_summary.txt; head -8 gpurun_out/${tag}_pose_launch_summary.txt
