# usage: gpurun --gpus 2 --timeout 300 -- 'bash scripts/r2_n2_colchunks.sh'
# Column chunks (Y<->Z chunked along x) against plane chunks, 640^3 complex128 on 1x2 (2.1 GB pencils, P = 2 on the wire).
mkdir -p gpurun_out
N=2
OUT=gpurun_out
export CUDECOMP_B200_DEVICE_TIMEOUT=20
i=0
bench() { # label, extra args...
  label=$1; shift
  i=$((i+1))
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
    --master-port $((29500+i*10)) bench.py --gpus $N --grid 640 --steps 20 --warmup 3 --no-cpu-baseline --no-e2e "$@" > $OUT/r2f_n${N}_$label.log 2>&1
  grep '"metric"' $OUT/r2f_n${N}_$label.log | tee $OUT/r2f_n${N}_$label.json | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); r = d['roofline']
    print('$label:', round(d['ms_per_step'], 3), 'ms/step;', d['path'], {k: round(v, 3) for k, v in r['per_op_ms'].items()},
          r['bound'], round(r['achieved'], 1), 'GB/s frac', round(r['frac'], 4), 'parity', d.get('parity', {}).get('ok'))" || tail -5 $OUT/r2f_n${N}_$label.log
}
bench inplace_auto --inplace
bench inplace_auto_planes --inplace --no-column-chunks
bench inplace_k4 --inplace --chunks 4
bench inplace_k5 --inplace --chunks 5
timeout 200 python -m pytest tests/test_zz_pipeline_gpu.py -q -m gpu -x -k "column or Stress" -p no:cacheprovider > $OUT/r2f_tests.log 2>&1; tail -2 $OUT/r2f_tests.log
