#!/bin/bash
# development aid: the -m gpu suite file by file, every file under its own wall-clock limit and every test under
# pytest-timeout (thread method: a hung kernel kills that pytest process and names the test), logs in gpurun_out/
out=${1:-gpurun_out/gpu_tests.log}
shift
files=${@:-$(ls tests/test_gpu_*.py)}
: > $out
for f in $files; do
    echo "=== $f" >> $out
    timeout 600 python -u -m pytest $f -x -q -m gpu --timeout 120 --timeout-method=thread -p no:cacheprovider 2>&1 | tail -12 >> $out
    echo "rc=$?" >> $out
done
grep -E "^===|passed|failed|error|Timeout|rc=" $out
