# usage: gpurun --gpus 2 --timeout 900 -- 'bash scripts/r2_n2_fused.sh'
# Round 2, second GPU call: the fused (single phased launch) in-place schedule -- correctness on 4 ranks over 2 GPUs,
# then timing against the round-1 staged schedules, plus the winners of the first call combined.
mkdir -p gpurun_out
N=2
OUT=gpurun_out
export CUDECOMP_B200_DEVICE_TIMEOUT=20
( nproc; free -g; lscpu | grep -E "Model name|Socket|NUMA node\(s\)" ) > $OUT/r2_host.txt 2>&1
i=0
bench() { # label, extra args...
  label=$1; shift
  i=$((i+1))
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
    --master-port $((29500+i*10)) bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --no-e2e "$@" > $OUT/r2b_n${N}_$label.log 2>&1
  grep '"metric"' $OUT/r2b_n${N}_$label.log | tee $OUT/r2b_n${N}_$label.json | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); r = d['roofline']
    print('$label:', round(d['ms_per_step'], 3), 'ms/step;', d['path'], {k: round(v, 3) for k, v in r['per_op_ms'].items()},
          r['bound'], round(r['achieved'], 1), 'GB/s frac', round(r['frac'], 4), 'parity', d.get('parity', {}).get('ok'))" || tail -5 $OUT/r2b_n${N}_$label.log
}
echo "== fused / chunked staged schedules on 4 ranks (2 per GPU): parity + back-to-back stress"
timeout 600 python -m pytest tests/test_zz_pipeline_gpu.py -q -m gpu -x -p no:cacheprovider > $OUT/r2b_pipeline_tests.log 2>&1
tail -15 $OUT/r2b_pipeline_tests.log
echo "== in place, 1x2: fused (default) vs round-1 schedules"
bench inplace_fused --inplace
bench inplace_fused_lag1 --inplace --lag 1
bench inplace_fused_lag3 --inplace --lag 3
bench inplace_fused_k8 --inplace --chunks 8
bench inplace_fused_k32 --inplace --chunks 32
bench inplace_fused_k1 --inplace --chunks 1
bench inplace_launches --inplace --staged-mode 1
bench inplace_fused_ctas444 --inplace --ctas 444
echo "== in place, 2x1"
bench inplace_fused_2x1 --inplace --pdims 2x1
echo "== out of place: combinations of what helped in the first call"
bench default
bench wide_ctas296 --wide --ctas 296
bench wide_tile64k --wide --tile-bytes 65536
bench ctas296_tile64k --ctas 296 --tile-bytes 65536
echo "== 512^3 complex64"
bench c64_512 --grid 512 --dtype float_complex
bench c64_512_inplace_fused --grid 512 --dtype float_complex --inplace
bench c64_512_inplace_fused_k4 --grid 512 --dtype float_complex --inplace --chunks 4
bench c64_512_wide --grid 512 --dtype float_complex --wide
