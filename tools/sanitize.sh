#!/bin/bash
# compute-sanitizer runs under gpurun: memcheck on smoke() and on the PRM
# connection / multi-GPU exchange tests, racecheck on smoke(); summaries land in gpurun_out/.
#   gpurun --gpus 2 --timeout 2400 -- 'bash tools/sanitize.sh'
set -u
OUT=gpurun_out
run() {  # name tool command...
    local name=$1 tool=$2; shift 2
    timeout 1500 compute-sanitizer --tool "$tool" --print-limit 20 --error-exitcode 7 "$@" > "$OUT/sanitizer_${name}.log" 2>&1
    echo "rc=$?" >> "$OUT/sanitizer_${name}.log"
    grep -E "ERROR SUMMARY|RACECHECK SUMMARY|rc=|passed|failed" "$OUT/sanitizer_${name}.log" | tail -4
}
run smoke_memcheck memcheck python -c "import __graft_entry__ as g; g.smoke()"
run smoke_racecheck racecheck python -c "import __graft_entry__ as g; g.smoke()"
run prm_memcheck memcheck python -m pytest tests/test_gpu_prm_connect.py -x -q -k "one_call or wraparound or sorted or leaves"
run prm_racecheck racecheck python -m pytest tests/test_gpu_prm_connect.py -x -q -k "maze1-6000 or sorted"
run multi_memcheck memcheck --target-processes all python -m pytest tests/test_gpu_multi.py -x -q
