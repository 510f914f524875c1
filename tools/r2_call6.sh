#!/usr/bin/env bash
# Round-2, GPU call 6 (N = 1): final validation of the shipped defaults + evidence for profiles/.
set -u
out=gpurun_out/r2c6
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
  tail -n 3 "$out/$name.log" | cut -c1-300
}
TMO=900 run pytest_all -- python -m pytest tests -x -q -m gpu
run smoke -- python -c "import __graft_entry__ as g; g.smoke()"
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras --no-self-check --no-kernel-rooflines"
run bench_a -- $B
run bench_mc128 OTB_GEMM_MC_MIN_TILES=128 -- $B
run bench_b -- $B
run bench_mc128_b OTB_GEMM_MC_MIN_TILES=128 -- $B
run selftest_mc128 OTB_GEMM_MC_MIN_TILES=128 -- build/selftest_gemm --bench
run fused_times -- python tools/prof_fused.py --time
TMO=300 run ncu_fused -- ncu --set full --clock-control none --import-source on -k regex:xattn_out_fused -s 1 -c 1 -o "$out/r02_xattn_fused" python tools/prof_fused.py
TMO=300 run ncu_gemm_dom -- ncu --set full --clock-control none -k regex:gemm2 -s 5 -c 1 -o "$out/r02_gemm2_ffn_up" build/selftest_gemm --no-cases --shape 2048 16384 4096 0 0 1
TMO=600 run ncu_launches -- ncu --metrics gpu__time_duration.sum --clock-control none -s 1700 -c 1150 --csv --log-file "$out/r02_launches_full.csv" python bench.py --steps 2 --warmup 1 --no-graph --no-cpu-baseline --no-extras --no-self-check --no-kernel-rooflines
TMO=900 run sanitizer -- compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_xattn_fused_gpu.py tests/test_persimmon_gpu.py tests/test_data_gpu.py tests/test_kernels_gpu.py -m gpu -x -q -k "not c2_shape"
TMO=900 run bench_full -- python bench.py --steps 20 --warmup 5
echo done
