# usage: gpurun --gpus 2 --timeout 700 -- 'bash scripts/r2_n2_fused2.sh'
# Round 2, third GPU call: the fused in-place schedule with segment-interleaved slots (the padded box interleave of the
# second call spent most of its time decoding empty slots), and the vectorised transpose kernel on the wire.
mkdir -p gpurun_out
N=2
OUT=gpurun_out
export CUDECOMP_B200_DEVICE_TIMEOUT=20
i=0
bench() { # label, extra args...
  label=$1; shift
  i=$((i+1))
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
    --master-port $((29500+i*10)) bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --no-e2e "$@" > $OUT/r2c_n${N}_$label.log 2>&1
  grep '"metric"' $OUT/r2c_n${N}_$label.log | tee $OUT/r2c_n${N}_$label.json | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); r = d['roofline']
    print('$label:', round(d['ms_per_step'], 3), 'ms/step;', d['path'], {k: round(v, 3) for k, v in r['per_op_ms'].items()},
          r['bound'], round(r['achieved'], 1), 'GB/s frac', round(r['frac'], 4), 'parity', d.get('parity', {}).get('ok'))" || tail -5 $OUT/r2c_n${N}_$label.log
}
echo "== fused staged schedules on 4 ranks (2 per GPU): parity + back-to-back stress"
timeout 400 python -m pytest tests/test_zz_pipeline_gpu.py -q -m gpu -x -p no:cacheprovider > $OUT/r2c_pipeline_tests.log 2>&1
tail -4 $OUT/r2c_pipeline_tests.log
echo "== in place, 1x2"
bench inplace_fused --inplace
bench inplace_fused_lag1 --inplace --lag 1
bench inplace_fused_k8 --inplace --chunks 8
bench inplace_fused_k4 --inplace --chunks 4
bench inplace_fused_k32 --inplace --chunks 32
bench inplace_fused_ctas444 --inplace --ctas 444
bench inplace_fused_ctas296 --inplace --ctas 296
bench inplace_fused_tile64k --inplace --tile-bytes 65536
echo "== in place, 2x1"
bench inplace_fused_2x1 --inplace --pdims 2x1
bench inplace_fused_2x1_k8 --inplace --pdims 2x1 --chunks 8
echo "== 512^3 complex64 in place"
bench c64_512_inplace_fused --grid 512 --dtype float_complex --inplace
bench c64_512_inplace_fused_k4 --grid 512 --dtype float_complex --inplace --chunks 4
bench c64_512_inplace_fused_k1 --grid 512 --dtype float_complex --inplace --chunks 1
echo "== axis-contiguous layout (real permutations): vectorised vs element-wise transpose kernel"
bench ac_c128 --axis-contiguous
bench ac_c128_elementwise --axis-contiguous --kernel-variant 3
bench ac_f32 --axis-contiguous --dtype float
bench ac_f32_elementwise --axis-contiguous --dtype float --kernel-variant 3
bench ac_c128_inplace --axis-contiguous --inplace
