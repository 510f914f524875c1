#!/usr/bin/env bash
# Round-2, GPU call 4: validate attention v3 (batched TMEM loads), the compile-time epilogue classes and the fused
# attention + to_out kernel; A/B them at step level; profiles of the attention kernels.
set -u
out=gpurun_out/r2c4
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
TMO=900 run pytest_all -- python -m pytest tests -m gpu -q
TMO=600 run pytest_all_fused OTB_XATTN_FUSED=1 -- python -m pytest tests/test_modules_gpu.py tests/test_c2_parity_gpu.py tests/test_fullsize_gpu.py tests/test_callers_gpu.py -m gpu -q
TMO=400 run selftest_cases -- build/selftest_gemm
run selftest_bench -- build/selftest_gemm --no-cases --bench
run selftest_bench_nocls OTB_GEMM_EPI_CLS=0 -- build/selftest_gemm --no-cases --bench
run attn_times -- python tools/prof_attn2.py --time
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras --no-self-check --no-kernel-rooflines"
run bench_a -- $B
run bench_fused OTB_XATTN_FUSED=1 -- $B
run bench_nocls OTB_GEMM_EPI_CLS=0 -- $B
run bench_b -- $B
run bench_fused_b OTB_XATTN_FUSED=1 -- $B
TMO=300 run ncu_attn -- ncu --set full --clock-control none --import-source on -k regex:attn -s 2 -c 10 -o "$out/r02_attn_v3" python tools/prof_attn2.py
TMO=900 run bench_full -- python bench.py --steps 20 --warmup 5
TMO=400 run ncu_launches -- ncu --metrics gpu__time_duration.sum --clock-control none -s 1200 -c 620 --csv --log-file "$out/r02_launches.csv" python bench.py --steps 2 --warmup 1 --no-graph --no-cpu-baseline --no-extras --no-self-check --no-kernel-rooflines
echo done
