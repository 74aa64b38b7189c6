#!/bin/bash
# Collects the round's evidence on the GPU box into gpurun_out/ (summaries are copied into profiles/ afterwards).
#   usage (under gpurun): bash tools/collect_profiles.sh <tag>
set -u
T=${1:-r1b}
O=gpurun_out
mkdir -p $O
python bench.py > $O/${T}_bench_n1.json 2> $O/${T}_bench_n1.err
python bench.py --impl reference --steps 2 --warmup 1 > $O/${T}_bench_reference.json 2>> $O/${T}_bench_n1.err
# launch list of the bench command (per-launch times are cold-cache and serialised: compare shares)
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/${T}_launches_bench.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > /dev/null 2>&1
python tools/launch_summary.py $O/${T}_launches_bench.csv > $O/${T}_launches_bench_lm_qr_c2.txt 2>&1
# DRAM traffic + duration of every trailing-update and panel-tree launch of ONE solve
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv \
    -k regex:"qr_apply_pp|qr_tree" --log-file $O/${T}_traffic_qr.csv python tools/profile_qr.py 100000 1000 1 > /dev/null 2>&1
# full captures: panel-0 tree kernel + the four trailing-update launches of panel 0
ncu --set full --import-source on --clock-control none -k regex:"qr_tree|qr_apply_pp" -c 5 -o $O/${T}_tree_pp -f \
    python tools/profile_qr.py 100000 1000 1 > /dev/null 2>&1
python tools/qr_timeline.py 2 > $O/${T}_timeline.txt 2>&1
python tools/apply_timing.py 2 > $O/${T}_apply_phase_timing.txt 2>&1
python tools/leaf_timing.py 200 32 > $O/${T}_tree_lone_block_timing.txt 2>&1
python tools/probe_qr.py > $O/${T}_probe_dense_c2.jsonl 2>&1
python tools/profile_chol.py 250000 4000 2 > $O/${T}_chol_c4shard.txt 2>&1
ls -la $O | tail -20
python tools/probe_configs.py c4 c5 > $O/${T}_probe_configs_c4shard_c5.jsonl 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > $O/${T}_smoke.txt 2>&1
