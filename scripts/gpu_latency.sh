mkdir -p gpurun_out
N=${1:-2}
TR="timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
i=0
run() {
  i=$((i+1))
  $TR --master-port $((29600+i*10)) bench.py --gpus $N --steps 50 --warmup 5 --no-e2e "$@" > gpurun_out/lat_$i.log 2>&1
  echo "== $@"; grep '"metric"' gpurun_out/lat_$i.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); nv=d.get('nvlink') or {}
    print(round(d['ms_per_step']*1e3,1), 'us/step;', d['path'], {k:round(v*1e3,1) for k,v in d['roofline']['per_op_ms'].items()}, 'host us/op', round(d['host_enqueue_us_per_op'],1), 'wire frac', nv.get('frac'))" || tail -5 gpurun_out/lat_$i.log
}
for g in 32 64 128 256 512; do run --grid $g --pdims ${N}x1 --dtype float_complex; done
run --grid 64 --pdims ${N}x1 --dtype float_complex --inplace
run --grid 256 --pdims ${N}x1 --dtype float_complex --inplace
