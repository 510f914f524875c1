#!/usr/bin/env bash
# Round-2, GPU call 1: validate the merged candidates, the new parity tests and the new bench; baseline profiles.
#   gpurun --timeout 2700 -- 'bash tools/r2_call1.sh'
set -u
out=gpurun_out/r2c1
mkdir -p "$out"
python -c "import __graft_entry__ as g; g.build()" > "$out/build.log" 2>&1 || { echo "build failed"; tail -5 "$out/build.log"; exit 1; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > "$out/gpu.txt"

run() {  # name, env assignments..., -- command
  local name=$1; shift
  local envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done
  shift
  echo "=== $name (${envs[*]:-default})"
  local t0=$SECONDS
  env "${envs[@]}" timeout "${TMO:-300}" "$@" > "$out/$name.log" 2> "$out/$name.err"
  echo "    exit $? ($((SECONDS - t0)) s)"
  tail -n 3 "$out/$name.log"
}

# the new warp-specialised attention forward (default on) is validated separately at the end: sections 1-6 run the
# round-1 resident kernel so that a bug in the new kernel cannot take the other validations down
export OTB_ATTN_WS=0

# 1. new parity tests first (the round-1 verdict's hole), then the default full suite
TMO=600 run pytest_c2_parity -- python -m pytest tests/test_c2_parity_gpu.py -m gpu -x -q
TMO=900 run pytest_default_full -- python -m pytest tests -m gpu -q --ignore tests/test_lm_gpu.py --ignore tests/test_callers_gpu.py --ignore tests/test_data_gpu.py --ignore tests/test_persimmon_gpu.py --ignore tests/test_multi_device_gpu.py
run pytest_lm -- python -m pytest tests/test_lm_gpu.py -m gpu -x -q
run pytest_callers -- python -m pytest tests/test_callers_gpu.py -m gpu -q
run pytest_data -- python -m pytest tests/test_data_gpu.py -m gpu -q
run pytest_persimmon -- python -m pytest tests/test_persimmon_gpu.py -m gpu -q
run pytest_multidev -- python -m pytest tests/test_multi_device_gpu.py -m gpu -q

# 2. GEMM self-test + micro-bench under each epilogue / dispatch setting
run selftest_default -- build/selftest_gemm --bench
run selftest_epi_tma OTB_GEMM_EPI_TMA=1 -- build/selftest_gemm --bench
run selftest_pairs32_tma OTB_GEMM2_MIN_PAIRS=32 OTB_GEMM_EPI_TMA=1 -- build/selftest_gemm --bench

# 3. parity with the candidates on
run pytest_kernels_tma OTB_GEMM_EPI_TMA=1 -- python -m pytest tests/test_kernels_gpu.py tests/test_c2_parity_gpu.py -m gpu -x -q
run pytest_lnfused OTB_LN_FUSED=1 -- python -m pytest tests/test_kernels_gpu.py tests/test_c2_parity_gpu.py -m gpu -x -q -k "layernorm or ln or gated"
run pytest_modules_all OTB_GEMM_EPI_TMA=1 OTB_LN_FUSED=1 OTB_GEMM2_MIN_PAIRS=32 -- python -m pytest tests/test_modules_gpu.py -m gpu -x -q

# 4. step-level A/B (same box, back to back)
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras --no-self-check --no-kernel-rooflines"
run bench_default -- $B
run bench_epi_tma OTB_GEMM_EPI_TMA=1 -- $B
run bench_lnfused OTB_LN_FUSED=1 -- $B
run bench_pairs32 OTB_GEMM2_MIN_PAIRS=32 -- $B
run bench_pairs32_tma OTB_GEMM2_MIN_PAIRS=32 OTB_GEMM_EPI_TMA=1 -- $B
run bench_multicast -- $B --multi-cast
run bench_all OTB_GEMM_EPI_TMA=1 OTB_LN_FUSED=1 OTB_GEMM2_MIN_PAIRS=32 -- $B --multi-cast
run bench_default_again -- $B
grep -h '"metric"' "$out"/bench_*.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['value'], d['ms_per_step'], d['roofline']['frac'], 'e2e', d['e2e']['value'], d['gpu_launches'])
"
# 5. the full new bench line (extras, self-check, HBM-kernel rooflines, CPU reference arm at batch 8)
TMO=900 run bench_full -- python bench.py --steps 20 --warmup 5
TMO=400 run bench_reference_arm -- python bench.py --impl reference --steps 3 --warmup 1

# 6. baseline profiles of the attention kernels at the step's shapes (ncu --set full) and the step's launch list
TMO=400 run ncu_attn -- ncu --set full --clock-control none --import-source on -k regex:attn -s 9 -c 3 -o "$out/r02_attn_base" python tools/prof_attn.py
TMO=600 run ncu_launches -- ncu --metrics gpu__time_duration.sum --clock-control none -s 1400 -c 700 --csv --log-file "$out/r02_launches_base.csv" python bench.py --steps 2 --warmup 1 --no-graph --no-cpu-baseline --no-extras --no-self-check --no-kernel-rooflines

# 7. the warp-specialised attention forward: kernel + module parity, micro-timings, step-level A/B
export OTB_ATTN_WS=1
run ws_pytest_kernels -- python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "attn or attention"
run ws_pytest_modules -- python -m pytest tests/test_modules_gpu.py tests/test_fullsize_gpu.py tests/test_c2_parity_gpu.py -m gpu -q
run ws_prof_attn -- python tools/prof_attn.py
OTB_ATTN_WS=0 run old_prof_attn -- python tools/prof_attn.py
run ws_bench -- $B
OTB_ATTN_WS=0 run old_bench -- $B
run ws_bench_all OTB_GEMM_EPI_TMA=1 OTB_LN_FUSED=1 OTB_GEMM2_MIN_PAIRS=32 -- $B --multi-cast
TMO=400 run ws_ncu_attn -- ncu --set full --clock-control none --import-source on -k regex:attn_fwd_ws -s 9 -c 3 -o "$out/r02_attn_ws" python tools/prof_attn.py
echo done
