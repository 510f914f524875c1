#!/usr/bin/env bash
# Round-2, GPU call 9 (N = 1): LLaMA decoder layer validation + final full run of the shipped defaults.
set -u
out=gpurun_out/r2c9
mkdir -p "$out"
python -c "import __graft_entry__ as g; g.build()" > "$out/build.log" 2>&1 || { echo "build failed"; tail -5 "$out/build.log"; exit 1; }
run() {
  local name=$1; shift
  local envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done
  shift
  echo "=== $name (${envs[*]:-default})"
  local t0=$SECONDS
  env "${envs[@]}" timeout "${TMO:-300}" "$@" > "$out/$name.log" 2> "$out/$name.err"
  echo "    exit $? ($((SECONDS - t0)) s)"
  tail -n 3 "$out/$name.log" | cut -c1-400
}
run pytest_llama -- python -m pytest tests/test_llama_gpu.py -q -m gpu
TMO=200 run sanitizer_llama -- compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_llama_gpu.py -m gpu -x -q -k "kernels or 100"
TMO=600 run pytest_all -- python -m pytest tests -x -q -m gpu
run smoke -- python -c "import __graft_entry__ as g; g.smoke()"
TMO=420 run bench -- python bench.py
