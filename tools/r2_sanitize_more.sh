#!/bin/bash
mkdir -p gpurun_out
( timeout 900 compute-sanitizer --tool initcheck --error-exitcode 1 python tools/sanitize.py; echo "initcheck rc=$?"; timeout 900 compute-sanitizer --tool synccheck --error-exitcode 1 python tools/sanitize.py; echo "synccheck rc=$?" ) > gpurun_out/sanitizer_more.txt 2>&1
grep -E "rc=|ERROR SUMMARY|Uninitialized|hazard|error" gpurun_out/sanitizer_more.txt | head -20
