# usage: gpurun --gpus 8 --timeout 900 -- 'bash scripts/r2_n8.sh'
# Round 2, the 8-GPU call: headline configuration out of place and in place (fused phased launch), parity inside every
# bench line, 8-rank parity tests with one GPU per rank, the reference's own FFT benchmark binary, the NCCL-restated arm
# as GPU-side baseline, BASELINE configs 2 and 4.
mkdir -p gpurun_out
N=8
OUT=gpurun_out
export CUDECOMP_B200_DEVICE_TIMEOUT=30
nvidia-smi topo -m > $OUT/r2_n8_topo.txt 2>&1
i=0
bench() { # label, extra args...
  label=$1; shift
  i=$((i+1))
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
    --master-port $((29500+i*10)) bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --no-e2e "$@" > $OUT/r2_n${N}_$label.log 2>&1
  grep '"metric"' $OUT/r2_n${N}_$label.log | tee $OUT/r2_n${N}_$label.json | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); r = d['roofline']
    print('$label:', round(d['ms_per_step'], 3), 'ms/step;', d['path'], {k: round(v, 3) for k, v in r['per_op_ms'].items()},
          r['bound'], round(r['achieved'], 1), 'GB/s frac', round(r['frac'], 4), 'parity', d.get('parity', {}).get('ok'))" || tail -5 $OUT/r2_n${N}_$label.log
}
echo "== headline: 1024^3 c128, 2x4"
bench default
bench inplace --inplace
bench default_128bit --no-wire-wide
bench inplace_k8 --inplace --chunks 8
bench inplace_k32 --inplace --chunks 32
bench inplace_launches --inplace --staged-mode 1
bench c64_512 --grid 512 --dtype float_complex
bench c64_512_inplace --grid 512 --dtype float_complex --inplace
bench ac_c128 --axis-contiguous
bench ac_c128_inplace --axis-contiguous --inplace
echo "== 8-rank parity, one GPU per rank; NCCL cross-check; fused schedules"
timeout 500 python -m pytest tests/test_gpu_parity.py tests/test_zz_nccl_crosscheck_gpu.py tests/test_zz_pipeline_gpu.py -q -m gpu -k "eight or nccl or Fused or Stress or fft" -p no:cacheprovider -rs > $OUT/r2_n8_tests.log 2>&1
tail -6 $OUT/r2_n8_tests.log
echo "== the reference's own FFT benchmark binary (benchmark/benchmark.cu, unmodified), 1024^3 c2c, 2x4"
refbench() { # label, args...
  label=$1; shift
  for r in $(seq 0 $((N-1))); do
    RANK=$r WORLD_SIZE=$N LOCAL_RANK=$r MASTER_ADDR=127.0.0.1 MASTER_PORT=$((29940+i)) timeout 300 oracle/_ref/benchmark_c2c "$@" \
      > $OUT/r2_n${N}_refbench_${label}.rank$r.log 2>&1 &
  done
  wait
  i=$((i+1))
  grep -E "Result Summary|FFTSize|GFLOPS|TIME|Max error|SELECTED|time|grid|backend" $OUT/r2_n${N}_refbench_${label}.rank0.log | head -14
}
refbench inplace --gx 1024 --gy 1024 --gz 1024 -r 2 -c 4 -b 4
refbench oop --gx 1024 --gy 1024 --gz 1024 -r 2 -c 4 -b 4 -o
refbench inplace_2048 --gx 2048 --gy 2048 --gz 2048 -r 2 -c 4 -b 4
echo "== FFT caller (python), in place and out of place"
for mode in "" "--inplace"; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29870 \
    bench/fft_benchmark.py --grid 1024 $mode > $OUT/r2_n8_fft${mode#--}.log 2>&1
  grep '^{' $OUT/r2_n8_fft${mode#--}.log | cut -c1-400
done
echo "== GPU-side baseline: the reference's NCCL arm restated (pack + NCCL all-to-all + unpack)"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29970 \
  bench/nccl_restated.py --grid 1024 --out $OUT/r2_n8_nccl_restated.json > $OUT/r2_n8_nccl_restated.log 2>&1
grep '^{' $OUT/r2_n8_nccl_restated.log | cut -c1-500 || tail -5 $OUT/r2_n8_nccl_restated.log
echo "== BASELINE config 4: halo-2 on 2048 x 2048 x 1024 float, 1x8, with the known-answer check"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29890 \
  bench/halo_benchmark.py > $OUT/r2_n8_halo.log 2>&1
grep '^{' $OUT/r2_n8_halo.log > $OUT/r2_n8_halo.json; python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2_n8_halo.json").readline())
    print("parity", d.get("parity", {}).get("ok"))
    for c in d["calls"]:
        print(c["pencil"], c["dim"], c["path"], round(c["us"], 1), "us", round(c["gbs"] or 0, 1), "GB/s")
except Exception as e:
    print("halo benchmark failed:", e)
PY
