#!/usr/bin/env bash
# Round-2 first GPU call: validate and A/B the opt-in candidates of branch r2-prep in ONE gpurun.
#   git merge r2-prep   (on main: .worktrees/ is not shipped to the GPU box), then
#   gpurun --timeout 1800 -- 'bash tools/r2_sweep.sh'
# Everything lands in gpurun_out/r2_sweep/.  Each step runs under its own timeout so a hang in a candidate
# (mbarrier waits trap after 2 s) cannot eat the call.
set -u
out=gpurun_out/r2_sweep
mkdir -p "$out"
python -c "import __graft_entry__ as g; g.build()" > "$out/build.log" 2>&1 || { echo "build failed"; tail -5 "$out/build.log"; exit 1; }

run() {  # name, env assignments..., -- command
  local name=$1; shift
  local envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done
  shift
  echo "=== $name (${envs[*]:-default})"
  env "${envs[@]}" timeout "${TMO:-300}" "$@" > "$out/$name.log" 2> "$out/$name.err"
  echo "    exit $?"
  tail -n 3 "$out/$name.log"
}

# 1. GEMM self-test + micro-bench (CLIP shapes, epilogue-only K=64 problems) under each epilogue / dispatch setting
run selftest_default -- build/selftest_gemm --bench
run selftest_epi_tma OTB_GEMM_EPI_TMA=1 -- build/selftest_gemm --bench
run selftest_pairs32 OTB_GEMM2_MIN_PAIRS=32 -- build/selftest_gemm --bench
run selftest_pairs32_tma OTB_GEMM2_MIN_PAIRS=32 OTB_GEMM_EPI_TMA=1 -- build/selftest_gemm --bench

# 2. parity suites with the candidates switched on (kernel tests first: they localise a failure)
run pytest_kernels_tma OTB_GEMM_EPI_TMA=1 -- python -m pytest tests/test_kernels_gpu.py -m gpu -x -q
run pytest_kernels_lnfused OTB_LN_FUSED=1 -- python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "layernorm or ln"
run pytest_multicast -- python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k multi_tensor_cast
run pytest_modules_all OTB_GEMM_EPI_TMA=1 OTB_LN_FUSED=1 -- python -m pytest tests/test_modules_gpu.py -m gpu -x -q

# §8f rank-1 candidate (new kernels, nothing else depends on them): run last among the parity suites
run pytest_lm -- python -m pytest tests/test_lm_gpu.py -m gpu -x -q

# the default path again after the merge (the kernels were re-templated): full GPU suite, no knobs
TMO=600 run pytest_default_full -- python -m pytest tests -m gpu -x -q

# 3. step-level A/B (same box, back to back)
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline"
run bench_default -- $B
run bench_epi_tma OTB_GEMM_EPI_TMA=1 -- $B
run bench_lnfused OTB_LN_FUSED=1 -- $B
run bench_pairs32 OTB_GEMM2_MIN_PAIRS=32 -- $B
run bench_multicast OTB_MULTI_CAST=1 -- $B
run bench_e2e_prefetch -- $B --e2e-prefetch
run bench_all OTB_GEMM_EPI_TMA=1 OTB_LN_FUSED=1 OTB_GEMM2_MIN_PAIRS=32 OTB_MULTI_CAST=1 -- $B
run bench_default_again -- $B
grep -h '"metric"' "$out"/bench_*.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['value'], d['ms_per_step'], d['roofline']['frac'], 'e2e', d['e2e']['value'])
"
