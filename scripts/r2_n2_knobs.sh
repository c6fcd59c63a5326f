# usage: gpurun --gpus 2 --timeout 900 -- 'bash scripts/r2_n2_knobs.sh'
#
# Round 2, first GPU call: every multi-rank schedule knob that shipped unmeasured in round 1, on 2 GPUs (the cheapest
# place where the wire path runs), plus the NCCL-restated arm (cross-check + timing) and the push/pull copy
# microbenchmark. Everything lands in gpurun_out/r2_n2_*; scripts/r2_summarize.py makes the table.
mkdir -p gpurun_out
N=2
OUT=gpurun_out
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
          r['bound'], round(r['achieved'], 1), 'GB/s frac', round(r['frac'], 4))" || tail -5 $OUT/r2_n${N}_$label.log
}
nvidia-smi topo -m > $OUT/r2_n2_topo.txt 2>&1
echo "== default / in place"
bench default
bench inplace --inplace
for k in 4 8 16; do bench inplace_chunks$k --inplace --chunks $k; done
echo "== kernel variants, who drives, tiles, grids"
bench bulk --bulk
bench wide --wide
bench pull --pull
bench pull_wide --pull --wide
bench pull_inplace --pull --inplace
bench tile16384 --tile-bytes 16384
bench tile131072 --tile-bytes 131072
bench ctas148 --ctas 148
bench ctas296 --ctas 296
bench ctas592 --ctas 592
bench balanced --balance-grid 1
bench grid2x1 --pdims 2x1
echo "== 512^3 complex64"
bench c64_512 --grid 512 --dtype float_complex
bench c64_512_tile16k --grid 512 --dtype float_complex --tile-bytes 16384
bench c64_512_balanced --grid 512 --dtype float_complex --balance-grid 1
bench c64_512_pull --grid 512 --dtype float_complex --pull
bench c64_512_inplace --grid 512 --dtype float_complex --inplace
echo "== NCCL arm restated: byte-for-byte cross-check (one GPU per rank) and timing"
timeout 400 python -m pytest tests/test_zz_nccl_crosscheck_gpu.py -q -m gpu -rxXs -p no:cacheprovider > $OUT/r2_n2_nccl_crosscheck.log 2>&1
tail -4 $OUT/r2_n2_nccl_crosscheck.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29970 \
  bench/nccl_restated.py --grid 1024 --out $OUT/r2_n2_nccl_restated.json > $OUT/r2_n2_nccl_restated.log 2>&1
grep '^{' $OUT/r2_n2_nccl_restated.log || tail -5 $OUT/r2_n2_nccl_restated.log
echo "== copy microbenchmark: push vs pull, 128- vs 256-bit"
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo bench/microbench_copy.cu -o /tmp/mb && \
  timeout 400 /tmp/mb 2048 > $OUT/r2_n2_microbench_copy.txt 2>&1
grep -E "copy engine|256-bit|PULL|simt U=4 cs grid=370|verify" $OUT/r2_n2_microbench_copy.txt | head -40
python scripts/r2_summarize.py $OUT > $OUT/r2_n2_summary.md 2>&1
