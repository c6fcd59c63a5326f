# usage: gpurun [--gpus N] --timeout T -- 'bash scripts/r2_gpu_confirm.sh N'
#
# First GPU calls of the next round: confirm on hardware everything that was written after the round-1 GPU budget was
# spent (tests/test_zz_* report XPASS when the code works) and measure the opt-in schedules so that their defaults can
# be decided from numbers. Everything lands in gpurun_out/r2_*; copy what matters to profiles/.
#
# A call on N GPUs is charged N x its box time against 180 GPU-minutes per round, so the work is split by where it is
# cheapest:             N=1  (~25 min = 25 GPU-min)  all test_zz_* files (ranks share the GPU), 1-GPU knobs, ncu captures
#                       N=2  (~20 min = 40 GPU-min)  every multi-rank knob (chunks, pull, wide, pairwise, bulk, tile, balance),
#                                                    copy microbenchmark (push vs pull, 128 vs 256 bit), NCCL-restated baseline
#                       N=8  (~9 min  = 72 GPU-min)  only the headline configuration: default, in place, the best two or three
#                                                    candidates from N=2 (edit CANDIDATES_8 below first), reference FFT benchmark
mkdir -p gpurun_out
N=${1:-1}
OUT=gpurun_out
# what the N=8 call tries besides default / in place: "label|bench.py args" -- trim to what N=2 showed to be worth it
CANDIDATES_8=${CANDIDATES_8:-"inplace_chunks8|--inplace --chunks 8
inplace_chunks16|--inplace --chunks 16
pull|--pull
wide|--wide
c64_512|--grid 512 --dtype float_complex
c64_512_balanced|--grid 512 --dtype float_complex --balance-grid 1
c64_512_inplace_chunks4|--grid 512 --dtype float_complex --inplace --chunks 4"}

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

if [ "$N" = 1 ]; then
  echo "== unconfirmed code paths (XPASS = confirmed)"
  timeout 1500 python -m pytest tests/test_zz_api_contract_gpu.py tests/test_zz_perf_report_gpu.py tests/test_zz_pipeline_gpu.py \
    tests/test_zz_ref_benchmark_gpu.py tests/test_zz_ref_ctest_gpu.py tests/test_zz_schedule_gpu.py -q -m gpu -rxX \
    -p no:cacheprovider > $OUT/r2_zz_tests.log 2>&1
  tail -40 $OUT/r2_zz_tests.log
  echo "== 1 GPU: default, kernel variants, tile size, balanced grid"
  bench default --no-e2e
  bench bulk --no-e2e --bulk
  bench wide --no-e2e --wide
  bench balanced --no-e2e --balance-grid 1
  bench c64_512 --no-e2e --grid 512 --dtype float_complex
  bench c64_512_balanced --no-e2e --grid 512 --dtype float_complex --balance-grid 1
  bench c64_512_tile16k --no-e2e --grid 512 --dtype float_complex --tile-bytes 16384
  bench c64_512_wide --no-e2e --grid 512 --dtype float_complex --wide
  echo "== the default line with the end-to-end leg (host-link ceiling, NUMA binding)"
  bench default_e2e
  echo "== ncu: launch list and full captures of the permuting kernel (axis-contiguous layout) and the bulk kernel"
  ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/r2_n1_launches_ac.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --axis-contiguous > $OUT/r2_n1_ncu_ac.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:transposeKernel -c 1 -o $OUT/r2_n1_transpose_full \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --axis-contiguous > $OUT/r2_n1_ncu_ac_full.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:rowCopyBulkKernel -c 1 -o $OUT/r2_n1_bulk_full \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --bulk > $OUT/r2_n1_ncu_bulk_full.log 2>&1
  # text exports travel back in any case; the reports themselves only when they are small (gpurun_out is capped at 64 MiB)
  for rep in $OUT/r2_n1_transpose_full $OUT/r2_n1_bulk_full; do
    [ -f $rep.ncu-rep ] || continue
    ncu -i $rep.ncu-rep --page raw --csv > $rep.raw.csv 2>/dev/null
    ncu -i $rep.ncu-rep --page details > $rep.details.txt 2>/dev/null
    [ $(stat -c %s $rep.ncu-rep) -gt 20000000 ] && rm -f $rep.ncu-rep
  done
fi

if [ "$N" = 2 ] || [ "$N" = 4 ]; then
  echo "== byte-for-byte cross-check against the restated NCCL arm (needs one GPU per rank)"
  timeout 600 python -m pytest tests/test_zz_nccl_crosscheck_gpu.py -q -m gpu -rxXs -p no:cacheprovider > $OUT/r2_n${N}_nccl_crosscheck.log 2>&1
  tail -4 $OUT/r2_n${N}_nccl_crosscheck.log
  echo "== default schedules, N=$N"
  bench default --no-e2e
  bench inplace --no-e2e --inplace
  echo "== chunked in-place schedule (cudecompB200SetPipelineChunks)"
  for k in 4 8 16; do bench inplace_chunks$k --no-e2e --inplace --chunks $k; done
  bench inplace_chunks8_1cta --no-e2e --inplace --chunks 8 --ctas 148
  echo "== kernel variants, who drives, slot order"
  bench bulk --no-e2e --bulk
  bench wide --no-e2e --wide
  bench pull --no-e2e --pull
  bench pull_wide --no-e2e --pull --wide
  bench pull_inplace --no-e2e --pull --inplace
  bench pull_inplace_chunks8 --no-e2e --pull --inplace --chunks 8
  bench pairwise --no-e2e --peer-order 1
  bench balanced --no-e2e --balance-grid 1
  bench tile16384 --no-e2e --tile-bytes 16384
  echo "== 512^3 complex64 (BASELINE config 2): handshake- and tail-sensitive"
  bench c64_512 --no-e2e --grid 512 --dtype float_complex
  bench c64_512_tile16k --no-e2e --grid 512 --dtype float_complex --tile-bytes 16384
  bench c64_512_balanced --no-e2e --grid 512 --dtype float_complex --balance-grid 1
  bench c64_512_wide --no-e2e --grid 512 --dtype float_complex --wide
  bench c64_512_pull --no-e2e --grid 512 --dtype float_complex --pull
  bench c64_512_inplace --no-e2e --grid 512 --dtype float_complex --inplace
  bench c64_512_inplace_chunks4 --no-e2e --grid 512 --dtype float_complex --inplace --chunks 4
  echo "== the default line with the end-to-end leg"
  bench default_e2e
  echo "== GPU-side baseline: the reference's NCCL arm restated (pack + NCCL all-to-all + unpack), same metric"
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29970 \
    bench/nccl_restated.py --grid 1024 --out $OUT/r2_n${N}_nccl_restated.json > $OUT/r2_n${N}_nccl_restated.log 2>&1
  grep '^{' $OUT/r2_n${N}_nccl_restated.log || tail -5 $OUT/r2_n${N}_nccl_restated.log
  echo "== autotune 768^3 with the schedule dimensions in the sweep (CUDECOMP_B200_AUTOTUNE_SCHEDULES=all)"
  CUDECOMP_B200_AUTOTUNE_SCHEDULES=all timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N \
    --master-addr 127.0.0.1 --master-port 29990 scripts/autotune_bench.py --grid 768 --backend --inplace > $OUT/r2_n${N}_autotune_all_inplace.log 2>&1
  grep -E "SELECTED|\"autotune\"" $OUT/r2_n${N}_autotune_all_inplace.log
  if [ "$N" = 2 ]; then
    echo "== copy microbenchmark: push vs pull, 128- vs 256-bit accesses"
    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo bench/microbench_copy.cu -o /tmp/mb && \
      timeout 600 /tmp/mb 2048 > $OUT/r2_n2_microbench_copy.txt 2>&1
    grep -E "copy engine|256-bit|PULL|simt U=4 cs grid=370|verify" $OUT/r2_n2_microbench_copy.txt | head -50
  fi
fi

if [ "$N" = 8 ]; then
  echo "== headline configuration on 8 GPUs"
  bench default
  bench inplace --no-e2e --inplace
  echo "$CANDIDATES_8" | while IFS='|' read -r label args; do
    [ -n "$label" ] && bench $label --no-e2e $args < /dev/null
  done
  echo "== GPU-side baseline: the reference's NCCL arm restated"
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29970 \
    bench/nccl_restated.py --grid 1024 --out $OUT/r2_n${N}_nccl_restated.json > $OUT/r2_n${N}_nccl_restated.log 2>&1
  grep '^{' $OUT/r2_n${N}_nccl_restated.log || tail -5 $OUT/r2_n${N}_nccl_restated.log
  if [ -x oracle/_ref/benchmark_c2c ]; then
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
    refbench inplace --gx 1024 --gy 1024 --gz 1024 -r 2 -c 4 -b 4
    refbench oop --gx 1024 --gy 1024 --gz 1024 -r 2 -c 4 -b 4 -o
    CUDECOMP_B200_PIPELINE_CHUNKS=8 refbench inplace_chunks8 --gx 1024 --gy 1024 --gz 1024 -r 2 -c 4 -b 4
  fi
fi

echo "== summary table"
python scripts/r2_summarize.py $OUT > $OUT/r2_n${N}_summary.md 2>&1
head -70 $OUT/r2_n${N}_summary.md
