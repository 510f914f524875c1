#!/usr/bin/env bash
# N-GPU check of the shipped multi-GPU default (wire format auto = bf16-direct) against the fp32 format:
#   gpurun --gpus N --timeout 900 -- 'bash tools/r2_scale_check.sh N'
set -u
N=${1:-4}
out=gpurun_out/scale_check_n$N
mkdir -p "$out"
python -c "import __graft_entry__ as g; g.build()" > "$out/build.log" 2>&1 || { echo "build failed"; exit 1; }
if [ "${2:-}" = "tests" ]; then
  timeout 400 python -m pytest tests/test_modules_gpu.py tests/test_fullsize_gpu.py tests/test_c2_parity_gpu.py tests/test_kernels_gpu.py -m gpu -q > "$out/pytest.log" 2>&1
  echo "pytest exit $?"; tail -n 2 "$out/pytest.log"
fi
port=29600
run() {
  local name=$1; shift
  port=$((port + 1))
  echo "=== $name"
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 \
    --master-port "$port" bench.py --gpus "$N" --steps 20 --warmup 5 --no-cpu-baseline --no-self-check --no-kernel-rooflines "$@" \
    > "$out/$name.log" 2> "$out/$name.err"
  echo "    exit $?"; grep -h '"metric"' "$out/$name.log" | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('   ', d['value'], 'samples/s', d['ms_per_step'], 'ms/step', d['config']['grad_allreduce_dtype'], d.get('extras'))"
}
run auto
run fp32 --grad-comm-dtype fp32 --no-extras
echo done
