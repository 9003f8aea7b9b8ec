#!/bin/bash
# Run on the GPU box (under gpurun): compute-sanitizer memcheck and racecheck over small encodes that reach every kernel
# family; the logs are the evidence kept under profiles/ (the reference relies on ASan/valgrind jobs).
mkdir -p gpurun_out
for TOOL in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $TOOL --print-limit 20 python tools/sanitize_cases.py > gpurun_out/sanitizer_$TOOL.log 2>&1
  echo "== $TOOL: exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize cases done|Error:|hazard" gpurun_out/sanitizer_$TOOL.log | head -8
done
