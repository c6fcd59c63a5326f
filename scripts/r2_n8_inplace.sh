# usage: gpurun --gpus 8 --timeout 240 -- 'bash scripts/r2_n8_inplace.sh'
# Round 2, last 8-GPU call: the in-place round trip with column chunks (Y<->Z chunked along x into 2 KiB rows, X<->Y
# along z), chunk counts 8 (what the library picks) / 4 / 6, then the reference's own FFT benchmark binary in place.
mkdir -p gpurun_out
N=8
OUT=gpurun_out
export CUDECOMP_B200_DEVICE_TIMEOUT=30
i=0
bench() { # label, extra args...
  label=$1; shift
  i=$((i+1))
  timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
    --master-port $((29500+i*10)) bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --no-e2e "$@" > $OUT/r2g_n${N}_$label.log 2>&1
  grep '"metric"' $OUT/r2g_n${N}_$label.log | tee $OUT/r2g_n${N}_$label.json | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); r = d['roofline']
    print('$label:', round(d['ms_per_step'], 3), 'ms/step;', d['path'], {k: round(v, 3) for k, v in r['per_op_ms'].items()},
          r['bound'], round(r['achieved'], 1), 'GB/s frac', round(r['frac'], 4), 'parity', d.get('parity', {}).get('ok'))" || tail -5 $OUT/r2g_n${N}_$label.log
}
bench inplace --inplace
bench inplace_k4 --inplace --chunks 4
bench inplace_k6 --inplace --chunks 6
for r in $(seq 0 $((N-1))); do
  RANK=$r WORLD_SIZE=$N LOCAL_RANK=$r MASTER_ADDR=127.0.0.1 MASTER_PORT=29941 timeout 100 oracle/_ref/benchmark_c2c \
    --gx 1024 --gy 1024 --gz 1024 -r 2 -c 4 -b 4 > $OUT/r2g_n8_refbench_inplace.rank$r.log 2>&1 &
done
wait
grep -E "Process grid|backend|GFLOPS|Max error" $OUT/r2g_n8_refbench_inplace.rank0.log | head -6
