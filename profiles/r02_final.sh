#!/bin/bash
# Final visit of the round (1 GPU): full GPU tests, sanitizer passes over the small all-kernel exercise, then the
# profiling visit (bench line, reference arm, ncu launch list, --set full captures, sort bandwidth).
tag=${1:-r03f}
out=gpurun_out/$tag
mkdir -p $out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > $out/tests.log 2>&1
echo "pytest exit $?" >> $out/tests.log
tail -4 $out/tests.log
for tool in memcheck racecheck synccheck; do
  AXCD_NO_GRAPH=1 timeout 600 compute-sanitizer --tool $tool --error-exitcode 9 python profiles/sanitize_small.py > $out/san_$tool.log 2>&1
  echo "$tool exit $?" | tee -a $out/san_$tool.log
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY" $out/san_$tool.log | tail -2
done
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python profiles/sanitize_small.py > $out/san_memcheck_graph.log 2>&1
echo "memcheck (graph launches) exit $?" | tee -a $out/san_memcheck_graph.log
bash profiles/r02_profile.sh $tag
