#!/bin/bash
# compute-sanitizer passes over the hot-path kernels at small sizes (memcheck: out-of-bounds / misaligned accesses, leaks of
# device errors; racecheck: shared-memory hazards; initcheck: reads of uninitialised global memory).  Writes the summaries
# to gpurun_out/sanitize_*.txt.   usage: tools/sanitize.sh [memcheck|racecheck|initcheck|synccheck]...
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TOOLS="${@:-memcheck}"
SEL='emulated or pipelined or redamp or kept or rank_deficient or underdetermined'
for tool in $TOOLS; do
  out=gpurun_out/sanitize_$tool.txt
  : > $out
  run() { echo "### $*" >> $out; timeout 600 compute-sanitizer --tool $tool --target-processes all --error-exitcode 9 "$@" 2>&1 | grep -v "^\s*$" | tail -25 >> $out; echo "exit ${PIPESTATUS[0]}" >> $out; }
  run python -c "import __graft_entry__ as g; g.smoke()"
  run python -m pytest tests/test_gpu_solvers.py -x -q -k "$SEL"
  run python -m pytest tests/test_gpu_sparse.py -x -q -k "fused_lsmr_equals or ragged"
  run python -m pytest tests/test_gpu_ozaki.py -x -q -k "not many_rows"
  grep -c "ERROR SUMMARY: 0 errors" $out | sed "s/^/$tool clean summaries: /"
  grep "ERROR SUMMARY" $out | sort | uniq -c
done
