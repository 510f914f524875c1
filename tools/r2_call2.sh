#!/usr/bin/env bash
# Round-2, GPU call 2: validate the second pile (callers / MPT LM / Flamingo / data / Persimmon / mask_ge / attention v2 with
# the new defaults), full bench line, and source-level profiles of the GEMM epilogue and the attention kernels.
set -u
out=gpurun_out/r2c2
mkdir -p "$out"
python -c "import __graft_entry__ as g; g.build()" > "$out/build.log" 2>&1 || { echo "build failed"; tail -5 "$out/build.log"; exit 1; }
export OTB_ATTN_WS=1      # validate the warp-specialised attention forward v2 (library default: off until validated)
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
run pytest_attn -- python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "attn or attention"
run pytest_persimmon -- python -m pytest tests/test_persimmon_gpu.py -m gpu -q
run pytest_callers -- python -m pytest tests/test_callers_gpu.py -m gpu -q
run pytest_data -- python -m pytest tests/test_data_gpu.py -m gpu -q
run pytest_multidev -- python -m pytest tests/test_multi_device_gpu.py -m gpu -q
run pytest_xfused -- python -m pytest tests/test_xattn_fused_gpu.py -m gpu -q
TMO=900 run pytest_rest -- python -m pytest tests -m gpu -q --ignore tests/test_persimmon_gpu.py --ignore tests/test_callers_gpu.py --ignore tests/test_data_gpu.py --ignore tests/test_multi_device_gpu.py --ignore tests/test_xattn_fused_gpu.py
run attn_times -- python tools/prof_attn2.py --time
OTB_ATTN_WS=0 run attn_times_old -- python tools/prof_attn2.py --time
TMO=900 run bench_full -- python bench.py --steps 20 --warmup 5
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras --no-self-check --no-kernel-rooflines"
run bench_a -- $B
run bench_noepitma OTB_GEMM_EPI_TMA=0 -- $B
run bench_noepicls OTB_GEMM_EPI_CLS=0 -- $B
run bench_nolnfused OTB_LN_FUSED=0 -- $B
run bench_nomulticast -- $B --no-multi-cast
run bench_b -- $B
G="build/selftest_gemm --no-cases --shape 2056 4096 1024 0 0 5 --shape 2056 4096 1024 0 0 0 --shape 2048 16384 64 0 0 0 --shape 2048 16384 64 0 0 1"
run selftest_shapes -- $G
run selftest_bench -- build/selftest_gemm --no-cases --bench
run selftest_bench_nocls OTB_GEMM_EPI_CLS=0 -- build/selftest_gemm --no-cases --bench
TMO=400 run selftest_cases -- build/selftest_gemm
for i in 0 1 2 3; do
  TMO=200 run ncu_gemm_epi$i -- ncu --set full --clock-control none --import-source on -k regex:gemm2 -s $((5 + 23 * i)) -c 1 -o "$out/r02_gemm_epi$i" $G
done
TMO=300 run ncu_attn -- ncu --set full --clock-control none --import-source on -k regex:attn -s 10 -c 10 -o "$out/r02_attn_v2" python tools/prof_attn2.py
TMO=400 run ncu_launches -- ncu --metrics gpu__time_duration.sum --clock-control none -s 1200 -c 620 --csv --log-file "$out/r02_launches.csv" python bench.py --steps 2 --warmup 1 --no-graph --no-cpu-baseline --no-extras --no-self-check --no-kernel-rooflines
echo done
