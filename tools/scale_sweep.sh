#!/usr/bin/env bash
# A/B of the gradient all-reduce options at N GPUs (default 2), back to back on one box:
#   gpurun --gpus 2 --timeout 900 -- 'bash tools/scale_sweep.sh 2'
# Results: gpurun_out/scale_sweep/*.log (one bench JSON line each).  NCCL_DEBUG=INFO of the first run shows whether
# NVLS / registered buffers were picked.
set -u
N=${1:-2}
out=gpurun_out/scale_sweep
mkdir -p "$out"
port=29500
run() {
  local name=$1; shift
  port=$((port + 1))
  echo "=== $name"
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 \
    --master-port "$port" bench.py --gpus "$N" --steps 20 --warmup 5 --no-cpu-baseline --no-extras --no-self-check --no-kernel-rooflines "$@" \
    > "$out/$name.log" 2> "$out/$name.err"
  echo "    exit $?"; grep -h '"metric"' "$out/$name.log" | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('   ', d['value'], 'samples/s', d['ms_per_step'], 'ms/step')"
}
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,COLL run fp32_debug --steps 3
run fp32
run bf16 --grad-comm-dtype bf16
run bf16_direct --grad-comm-dtype bf16-direct
run bf16_direct_registered --grad-comm-dtype bf16-direct --nccl-registered
run fp32_registered --nccl-registered
run bf16_registered --grad-comm-dtype bf16 --nccl-registered
run fp32_again
