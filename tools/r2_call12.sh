#!/usr/bin/env bash
# Round-2, GPU call 12 (N = 1): last validation of the shipped tree — full GPU suite, smoke, one short bench.
set -u
out=gpurun_out/r2c12
mkdir -p "$out"
python -c "import __graft_entry__ as g; g.build()" > "$out/build.log" 2>&1 || { echo "build failed"; tail -5 "$out/build.log"; exit 1; }
echo "=== pytest"; timeout 200 python -m pytest tests -x -q -m gpu > "$out/pytest.log" 2>&1; echo "exit $?"; tail -n 3 "$out/pytest.log"
echo "=== smoke"; timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > "$out/smoke.log" 2>&1; echo "exit $?"; tail -n 1 "$out/smoke.log"
echo "=== bench"; timeout 150 python bench.py --no-extras --no-cpu-baseline > "$out/bench.log" 2> "$out/bench.err"; echo "exit $?"; cut -c1-300 "$out/bench.log" | tail -n 1
