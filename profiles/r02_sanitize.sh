#!/bin/bash
tag=${1:-r02s}
out=gpurun_out/$tag
mkdir -p $out
for tool in memcheck racecheck synccheck; do
  AXCD_NO_GRAPH=1 timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python profiles/sanitize_small.py > $out/san_$tool.log 2>&1
  echo "$tool exit $?" | tee -a $out/san_$tool.log
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY" $out/san_$tool.log | tail -2
done
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python profiles/sanitize_small.py > $out/san_memcheck_graph.log 2>&1
echo "memcheck (graph launches) exit $?" | tee -a $out/san_memcheck_graph.log
