#!/bin/bash
out=gpurun_out/${1:-sanitize}; mkdir -p $out
for tool in memcheck racecheck initcheck synccheck; do
  timeout 600 compute-sanitizer --tool $tool --print-limit 5 python profiles/sanitize_small.py > $out/$tool.log 2>&1
  echo "== $tool: exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard" $out/$tool.log | tail -3; grep -E "^(C0|C2|C3|coherent)" $out/$tool.log | head -4
done
