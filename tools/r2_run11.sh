#!/bin/bash
mkdir -p gpurun_out
timeout 1700 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_target.py > gpurun_out/r02_sanitize.log 2>&1; echo "sanitizer rc=$?"
grep -E "ok|ERROR SUMMARY|Invalid|Error" gpurun_out/r02_sanitize.log | tail -40
