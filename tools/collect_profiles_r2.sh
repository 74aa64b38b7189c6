#!/bin/bash
# Round-2 evidence, collected on the GPU box into gpurun_out/ (summaries are copied into profiles/ afterwards).
#   usage (under gpurun): bash tools/collect_profiles_r2.sh <tag>
set -u
T=${1:-r2}
O=gpurun_out
mkdir -p $O
python bench.py > $O/${T}_bench_n1.json 2> $O/${T}_bench_n1.err
python bench.py --impl reference --steps 2 --warmup 1 > $O/${T}_bench_reference.json 2>> $O/${T}_bench_n1.err
# the reference arm as torch.distributed.run launches it (OMP_NUM_THREADS=1 in the environment): must still use all cores
OMP_NUM_THREADS=1 python bench.py --impl reference --steps 2 --warmup 1 > $O/${T}_bench_reference_omp1env.json 2>> $O/${T}_bench_n1.err
# launch list of the headline bench command (per-launch times are cold-cache and serialised: compare shares)
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/${T}_launches_bench.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-other-configs --no-resolve > /dev/null 2>&1
python tools/launch_summary.py $O/${T}_launches_bench.csv > $O/${T}_launches_bench_lm_qr_c2.txt 2>&1
# full captures: panel-0 tree kernel + the trailing-update launches of panel 0
ncu --set full --import-source on --clock-control none -k regex:"qr_tree|qr_apply_pp" -c 5 -o $O/${T}_tree_pp -f \
    python tools/profile_qr.py 100000 1000 1 > /dev/null 2>&1
python tools/qr_timeline.py 2 > $O/${T}_timeline.txt 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > $O/${T}_smoke.txt 2>&1
ls -la $O | tail -12
