#!/usr/bin/env bash
# 8-GPU check: shipped default (bf16-direct), the same with an NCCL-registered wire buffer, NCCL_DEBUG lines of the first.
set -u
N=8
out=gpurun_out/scale_check_n8
mkdir -p "$out"
python -c "import __graft_entry__ as g; g.build()" > "$out/build.log" 2>&1 || { echo "build failed"; exit 1; }
port=29700
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
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,TUNING run auto
run auto_registered --nccl-registered --no-extras
run fp32 --grad-comm-dtype fp32 --no-extras
grep -E "NVLS|Connected all|AllReduce.*(Algo|algo)|TUNING|nChannels" "$out/auto.log" | grep -v "Channel [0-9][0-9]/" | head -40 > "$out/nccl_lines.txt"
echo done
