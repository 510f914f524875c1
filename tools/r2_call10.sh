#!/usr/bin/env bash
# Round-2, GPU call 10 (N = 1): fp32-grade backward validation + the full GPU suite.
set -u
out=gpurun_out/r2c10
mkdir -p "$out"
python -c "import __graft_entry__ as g; g.build()" > "$out/build.log" 2>&1 || { echo "build failed"; tail -5 "$out/build.log"; exit 1; }
run() {
  local name=$1; shift
  echo "=== $name"
  local t0=$SECONDS
  timeout "${TMO:-300}" "$@" > "$out/$name.log" 2> "$out/$name.err"
  echo "    exit $? ($((SECONDS - t0)) s)"
  tail -n 25 "$out/$name.log" | cut -c1-600
}
run pytest_fp32_bwd python -m pytest tests/test_fp32_backward_gpu.py -q -m gpu
TMO=200 run sanitizer_fp32_bwd compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_fp32_backward_gpu.py -m gpu -x -q -k "kernel or perceiver_block or two_images"
TMO=600 run pytest_all python -m pytest tests -x -q -m gpu
run smoke python -c "import __graft_entry__ as g; g.smoke()"
