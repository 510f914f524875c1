#!/usr/bin/env bash
# N-GPU A/B of the wire-format all-reduce's reduction: SUM (+ 1/N in the up-cast) vs ncclAvg, with NCCL's algorithm choice logged:
#   gpurun --gpus 4 --timeout 600 -- 'bash tools/r2_reduce_op_check.sh 4'
set -u
N=${1:-4}
out=gpurun_out/reduce_op_n$N
mkdir -p "$out"
python -c "import __graft_entry__ as g; g.build()" > "$out/build.log" 2>&1 || { echo "build failed"; exit 1; }
port=29700
run() {
  local name=$1; shift
  local envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done
  shift
  port=$((port + 1))
  echo "=== $name (${envs[*]:-})"
  env NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,TUNING "${envs[@]}" timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" \
    --master-addr 127.0.0.1 --master-port "$port" bench.py --gpus "$N" --steps 20 --warmup 5 --no-cpu-baseline --no-self-check \
    --no-kernel-rooflines --no-extras "$@" > "$out/$name.log" 2> "$out/$name.err"
  echo "    exit $?"
  grep -h "AllReduce: 2" "$out/$name.log" "$out/$name.err" | sort | uniq -c | sort -rn | head -3
  grep -h '"metric"' "$out/$name.log" | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('   ', d['value'], 'samples/s', d['ms_per_step'], 'ms/step', d['config'].get('grad_allreduce_op'), 'e2e', d['e2e']['value'])"
}
run sum -- --reduce-op sum
run avg -- --reduce-op avg
if ! grep -h "AllReduce: 2" "$out/sum.log" "$out/sum.err" | grep -q NVLS; then
  run sum_forced_nvls NCCL_ALGO=NVLS -- --reduce-op sum
fi
run sum_b -- --reduce-op sum
grep -h "NCCL INFO" "$out/sum.log" "$out/sum.err" | grep -E "Algorithm|Protocol|AllReduce \||NVLS" | head -20 > "$out/nccl_lines.txt"
echo done
