#!/bin/bash
# DEVELOPMENT TOOL (GPU box, one GPU): the two ncu captures that profiles/ keeps per round.
#   tools/ncu_capture.sh <tag>          e.g. r04  ->  gpurun_out/<tag>_launches_ncu.csv, gpurun_out/<tag>_full.ncu-rep, gpurun_out/<tag>_full_raw.csv
# 1. launch list: every launch of 6 eager xi-slices of the bench workload's transverse shape (512 x 512, 4 ppc, M = 5) with its device time, DRAM bytes
#    and warp instructions (cold caches, serialised: compare SHARES with the in-graph timeline, not absolutes);
# 2. `--set full` of the slice's top kernels (one slice's worth of the predictor, the two tile kernels around it, the column transform, push, slot copies).
set -e
TAG=${1:-r04}
HERE=$(cd "$(dirname "$0")" && pwd)
ROOT=$(cd "$HERE/.." && pwd)
cd "$ROOT"
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum --clock-control none \
    --csv --log-file gpurun_out/${TAG}_launches_ncu.csv python tools/profile_slices.py --eager --slices 6 > gpurun_out/${TAG}_launches.log 2>&1
ncu --set full --clock-control none --import-source on \
    -k 'regex:k_particles_predict_dep|k_tile_cell_deposit|k_tile_assemble|k_p2_cols_solve|k_particles_advance|k_rank_to_slots|k_p2_rows_c2r' \
    -s 58 -c 29 -f -o gpurun_out/${TAG}_full python tools/profile_slices.py --eager --slices 4 > gpurun_out/${TAG}_full.log 2>&1
ncu -i gpurun_out/${TAG}_full.ncu-rep --page raw --csv > gpurun_out/${TAG}_full_raw.csv
python tools/launch_summary.py gpurun_out/${TAG}_launches_ncu.csv 6 > gpurun_out/${TAG}_launches_summary.txt
python tools/ncu_summary.py gpurun_out/${TAG}_full_raw.csv > gpurun_out/${TAG}_ncu_full_summary.txt
tail -30 gpurun_out/${TAG}_launches_summary.txt
