mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo2.txt 2>&1
(python -m pytest tests/test_gpu_parity.py -q -m gpu -k "two or three") > gpurun_out/t2.log 2>&1; tail -3 gpurun_out/t2.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
i=0
for args in "" "--pdims 2x1" "--inplace" "--inplace --pdims 2x1" "--staged" "--ctas 148" "--ctas 296" "--ctas 592" "--grid 512 --dtype float_complex" "--axis-contiguous" ; do
  i=$((i+1))
  $TR --master-port $((29600+i*10)) bench.py --gpus 2 --steps 10 --warmup 3 --no-e2e $args > gpurun_out/n2_$i.log 2>&1
  echo "== $args"; grep '"metric"' gpurun_out/n2_$i.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(round(d['value'],1), round(d['ms_per_step'],3), d['path'], {k:round(v,3) for k,v in d['roofline']['per_op_ms'].items()}, 'nvl', d.get('nvlink',{}).get('achieved'))" || tail -5 gpurun_out/n2_$i.log
done
