mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo8.txt 2>&1
N=${1:-8}
TR="timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
i=0
run() {
  i=$((i+1))
  $TR --master-port $((29600+i*10)) bench.py --gpus $N --steps 10 --warmup 3 "$@" > gpurun_out/n${N}_$i.log 2>&1
  echo "== $@"; grep '"metric"' gpurun_out/n${N}_$i.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); nv=d.get('nvlink') or {}
    print(round(d['value'],1), 'GB/s agg;', round(d['per_gpu_value'],1), 'per GPU;', round(d['ms_per_step'],3), 'ms;', d['path'], {k:round(v,3) for k,v in d['roofline']['per_op_ms'].items()}, 'wire GB/s', nv.get('achieved'), 'frac', nv.get('frac'), 'e2e', (d.get('e2e') or {}).get('value'))" || tail -5 gpurun_out/n${N}_$i.log
}
run
run --no-e2e --inplace
run --no-e2e --grid 512 --dtype float_complex
run --no-e2e --grid 512 --dtype float_complex --inplace
run --no-e2e --pdims 1x$N
run --no-e2e --pdims ${N}x1
run --no-e2e --pdims 4x2
run --no-e2e --axis-contiguous
run --no-e2e --grid 2048
run --no-e2e --ctas 296
$TR --master-port 29900 scripts/autotune_bench.py --grid 768 --backend > gpurun_out/autotune_768.log 2>&1; grep -E "SELECTED|autotune|grid:|Total time" gpurun_out/autotune_768.log | head -60
