#!/bin/bash
# compute-sanitizer passes over small invocations of every kernel family; summary -> gpurun_out/r02_sanitizer.txt
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
out=gpurun_out/r02_sanitizer.txt
: > $out
for tool in memcheck racecheck synccheck initcheck; do
  echo "=== compute-sanitizer --tool $tool" >> $out
  timeout 1500 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_driver.py > gpurun_out/sanitizer_$tool.log 2>&1
  echo "exit $?" >> $out
  grep -E "ok$|MISMATCH|ERROR SUMMARY|RACECHECK SUMMARY|Error:|Hazard|hazard" gpurun_out/sanitizer_$tool.log | sort | uniq -c | head -40 >> $out
done
cat $out
