# usage: gpurun --gpus 2 --timeout 400 -- 'bash scripts/r2_n2_grid.sh'
# Does the CTA count matter through its common factors with the slot interleave? With 296 or 370 CTAs and 2 or 4 boxes a
# CTA sees the same box(es) for the whole launch; a prime count makes every CTA walk all boxes / segments.
# 640^3 complex128 on 1x2 gives the 2.1 GB pencils of the 8-GPU headline configuration (P = 2 on the wire).
mkdir -p gpurun_out
N=2
OUT=gpurun_out
export CUDECOMP_B200_DEVICE_TIMEOUT=20
i=0
bench() { # label, extra args...
  label=$1; shift
  i=$((i+1))
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
    --master-port $((29500+i*10)) bench.py --gpus $N --grid 640 --steps 20 --warmup 3 --no-cpu-baseline --no-e2e "$@" > $OUT/r2e_n${N}_$label.log 2>&1
  grep '"metric"' $OUT/r2e_n${N}_$label.log | tee $OUT/r2e_n${N}_$label.json | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); r = d['roofline']
    print('$label:', round(d['ms_per_step'], 3), 'ms/step;', d['path'], {k: round(v, 3) for k, v in r['per_op_ms'].items()},
          r['bound'], round(r['achieved'], 1), 'GB/s frac', round(r['frac'], 4), 'parity', d.get('parity', {}).get('ok'))" || tail -5 $OUT/r2e_n${N}_$label.log
}
bench oop
bench oop_293 --ctas 293
bench inplace_k8 --inplace --chunks 8
bench inplace_k8_293 --inplace --chunks 8 --ctas 293
bench inplace_k4_293 --inplace --chunks 4 --ctas 293
bench inplace_k16_293 --inplace --chunks 16 --ctas 293
bench inplace_k8_293_tile64k --inplace --chunks 8 --ctas 293 --tile-bytes 65536
bench inplace_k8_293_head50 --inplace --chunks 8 --ctas 293 --phase-head 50
bench inplace_k8_128bit_367 --inplace --chunks 8 --ctas 367 --no-wire-wide
