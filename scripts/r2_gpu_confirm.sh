# usage: gpurun [--gpus N] -- 'bash scripts/r2_gpu_confirm.sh N'
# First GPU call of the next round: confirms on hardware everything that was written after the round-1 GPU budget was
# spent (the tests/test_zz_* files report XPASS when the code works), then measures the opt-in schedules so that their
# defaults can be decided from numbers. Everything lands in gpurun_out/r2_*.{log,json}; copy what matters to profiles/.
mkdir -p gpurun_out
N=${1:-1}
OUT=gpurun_out
echo "== unconfirmed code paths (XPASS = confirmed)"
timeout 1500 python -m pytest tests/test_zz_api_contract_gpu.py tests/test_zz_perf_report_gpu.py tests/test_zz_pipeline_gpu.py \
  tests/test_zz_ref_benchmark_gpu.py tests/test_zz_ref_ctest_gpu.py tests/test_zz_schedule_gpu.py tests/test_zz_nccl_crosscheck_gpu.py -q -m gpu -rxX -p no:cacheprovider > $OUT/r2_zz_tests.log 2>&1
tail -40 $OUT/r2_zz_tests.log

i=0
bench() { # label, extra args...
  label=$1; shift
  i=$((i+1))
  if [ "$N" = 1 ]; then
    timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline "$@" > $OUT/r2_n${N}_$label.log 2>&1
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
      --master-port $((29500+i*10)) bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline "$@" > $OUT/r2_n${N}_$label.log 2>&1
  fi
  grep '"metric"' $OUT/r2_n${N}_$label.log | tee $OUT/r2_n${N}_$label.json | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); r = d['roofline']
    print('$label:', round(d['ms_per_step'], 3), 'ms/step;', d['path'], {k: round(v, 3) for k, v in r['per_op_ms'].items()},
          r['bound'], round(r['achieved'], 1), 'GB/s frac', round(r['frac'], 4))" || tail -5 $OUT/r2_n${N}_$label.log
}

echo "== default schedules, N=$N"
bench default --no-e2e
bench inplace --no-e2e --inplace
if [ "$N" != 1 ]; then
  # these only change something when a communicator has more than one rank
  echo "== chunked in-place schedule (cudecompB200SetPipelineChunks)"
  for k in 2 4 8 16; do bench inplace_chunks$k --no-e2e --inplace --chunks $k; done
  bench inplace_chunks8_1cta --no-e2e --inplace --chunks 8 --ctas 148
  echo "== TMA bulk row copy (cudecompB200SetKernelVariant)"
  bench bulk --no-e2e --bulk
  bench bulk_inplace_chunks8 --no-e2e --bulk --inplace --chunks 8
  echo "== pairwise slot order"
  bench pairwise --no-e2e --peer-order 1
  echo "== receiver-driven direct transposes (cudecompB200SetTransferMode)"
  bench pull --no-e2e --pull
  bench pull_wide --no-e2e --pull --wide
else
  bench bulk --no-e2e --bulk
fi
echo "== 256-bit LDG/STG row copy (kernel variant 2)"
bench wide --no-e2e --wide
echo "== tile size / balanced grid (cudecompB200SetSchedule)"
for t in 16384 65536; do bench tile$t --no-e2e --tile-bytes $t; done
bench balanced --no-e2e --balance-grid 1
echo "== 512^3 complex64 (BASELINE config 2): handshake- and tail-sensitive"
bench c64_512 --no-e2e --grid 512 --dtype float_complex
for t in 8192 16384; do bench c64_512_tile$t --no-e2e --grid 512 --dtype float_complex --tile-bytes $t; done
bench c64_512_balanced --no-e2e --grid 512 --dtype float_complex --balance-grid 1
bench c64_512_balanced_tile16k --no-e2e --grid 512 --dtype float_complex --balance-grid 1 --tile-bytes 16384
bench c64_512_wide --no-e2e --grid 512 --dtype float_complex --wide
bench c64_512_inplace --no-e2e --grid 512 --dtype float_complex --inplace
if [ "$N" != 1 ]; then
  bench c64_512_inplace_chunks4 --no-e2e --grid 512 --dtype float_complex --inplace --chunks 4
  bench c64_512_pairwise --no-e2e --grid 512 --dtype float_complex --peer-order 1
  bench c64_512_pull --no-e2e --grid 512 --dtype float_complex --pull
fi
echo "== the default line with the end-to-end leg (host-link ceiling, NUMA binding)"
bench default_e2e

if [ "$N" = 1 ]; then
  echo "== ncu: launch list and one full capture of the transpose (permuting) kernel, axis-contiguous layout"
  ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/r2_n1_launches_ac.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --axis-contiguous > $OUT/r2_n1_ncu_ac.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:transposeKernel -c 1 -o $OUT/r2_n1_transpose_full \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --axis-contiguous > $OUT/r2_n1_ncu_ac_full.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:rowCopyBulkKernel -c 1 -o $OUT/r2_n1_bulk_full \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --bulk > $OUT/r2_n1_ncu_bulk_full.log 2>&1
fi

if [ "$N" != 1 ]; then
  echo "== autotune 768^3 with the schedule dimensions in the sweep (CUDECOMP_B200_AUTOTUNE_SCHEDULES=all)"
  CUDECOMP_B200_AUTOTUNE_SCHEDULES=all timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N \
    --master-addr 127.0.0.1 --master-port 29990 scripts/autotune_bench.py --grid 768 --backend > $OUT/r2_n${N}_autotune_all.log 2>&1
  grep -E "SELECTED|\"autotune\"" $OUT/r2_n${N}_autotune_all.log
  CUDECOMP_B200_AUTOTUNE_SCHEDULES=all timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N \
    --master-addr 127.0.0.1 --master-port 29980 scripts/autotune_bench.py --grid 768 --backend --inplace > $OUT/r2_n${N}_autotune_all_inplace.log 2>&1
  grep -E "SELECTED|\"autotune\"" $OUT/r2_n${N}_autotune_all_inplace.log
fi

if [ "$N" != 1 ]; then
  echo "== GPU-side baseline: the reference's NCCL arm restated (pack + NCCL all-to-all + unpack), same metric"
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29970 \
    bench/nccl_restated.py --grid 1024 --out $OUT/r2_n${N}_nccl_restated.json > $OUT/r2_n${N}_nccl_restated.log 2>&1
  grep '^{' $OUT/r2_n${N}_nccl_restated.log || tail -5 $OUT/r2_n${N}_nccl_restated.log
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29960 \
    bench/nccl_restated.py --grid 512 --dtype float_complex > $OUT/r2_n${N}_nccl_restated_512.log 2>&1
  grep '^{' $OUT/r2_n${N}_nccl_restated_512.log || tail -5 $OUT/r2_n${N}_nccl_restated_512.log
  timeout 900 python -m pytest tests/test_zz_nccl_crosscheck_gpu.py -q -m gpu -rxXs -p no:cacheprovider > $OUT/r2_n${N}_nccl_crosscheck.log 2>&1
  tail -5 $OUT/r2_n${N}_nccl_crosscheck.log
fi

if [ "$N" != 1 ] && [ -x oracle/_ref/benchmark_c2c ]; then
  echo "== the reference's own FFT benchmark binary (benchmark/benchmark.cu, unmodified) on this library, 1024^3 c2c"
  refbench() { # label, args...
    label=$1; shift
    for r in $(seq 0 $((N-1))); do
      RANK=$r WORLD_SIZE=$N LOCAL_RANK=$r MASTER_ADDR=127.0.0.1 MASTER_PORT=29940 timeout 600 oracle/_ref/benchmark_c2c "$@" \
        > $OUT/r2_n${N}_refbench_${label}.rank$r.log 2>&1 &
    done
    wait
    grep -E "Result Summary|FFTSize|GFLOPS|TIME|Max error|SELECTED|time" $OUT/r2_n${N}_refbench_${label}.rank0.log | head -20
  }
  PR=$(python -c "print({2:1,4:2,8:2}.get($N,1))"); PC=$((N/PR))
  refbench inplace --gx 1024 --gy 1024 --gz 1024 -r $PR -c $PC -b 4
  refbench oop --gx 1024 --gy 1024 --gz 1024 -r $PR -c $PC -b 4 -o
  refbench oop_ac --gx 1024 --gy 1024 --gz 1024 -r $PR -c $PC -b 4 -o --acx 1 --acy 1 --acz 1
  CUDECOMP_B200_PIPELINE_CHUNKS=8 refbench inplace_chunks8 --gx 1024 --gy 1024 --gz 1024 -r $PR -c $PC -b 4
  refbench autotune --gx 1024 --gy 1024 --gz 1024 -r 0 -c 0 -b 0 -o
fi
echo "== summary table"; python scripts/r2_summarize.py $OUT > $OUT/r2_n${N}_summary.md 2>&1; cat $OUT/r2_n${N}_summary.md | head -60
if [ "$N" = 2 ]; then echo "== copy microbenchmark: push vs pull, 128- vs 256-bit accesses"; nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo bench/microbench_copy.cu -o /tmp/mb && timeout 600 /tmp/mb 2048 > $OUT/r2_n2_microbench_copy.txt 2>&1; grep -E "copy engine|256-bit|PULL|simt U=4 cs grid=370|verify" $OUT/r2_n2_microbench_copy.txt | head -40; fi
