# usage: gpu_final.sh N   -- the per-N measurement set recorded under profiles/
mkdir -p gpurun_out
N=${1:-8}
TR="timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
i=0
run() {
  i=$((i+1))
  $TR --master-port $((29600+i*10)) bench.py --gpus $N --steps 10 --warmup 3 "$@" > gpurun_out/f${N}_$i.log 2>&1
  echo "== bench.py --gpus $N $@"; grep '"metric"' gpurun_out/f${N}_$i.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); r=d['roofline']
    print(round(d['value'],1), 'GB/s agg;', round(d['per_gpu_value'],1), 'per GPU;', round(d['ms_per_step'],3), 'ms;', d['path'], {k:round(v,3) for k,v in r['per_op_ms'].items()}, r['bound'], round(r['achieved'],1), 'frac', round(r['frac'],4), 'e2e', (d.get('e2e') or {}).get('value'))" || tail -5 gpurun_out/f${N}_$i.log
}
run
run --no-e2e --inplace
run --no-e2e --grid 512 --dtype float_complex
echo "== 8-rank parity"; (python -m pytest tests/test_gpu_parity.py -q -m gpu -k "eight or four_ranks and Baseline") > gpurun_out/t_eight_${N}.log 2>&1; tail -3 gpurun_out/t_eight_${N}.log
echo "== fft_benchmark 1024"
$TR --master-port 29800 bench/fft_benchmark.py --grid 1024 > gpurun_out/fft_${N}.log 2>&1; grep '^{' gpurun_out/fft_${N}.log || tail -5 gpurun_out/fft_${N}.log
$TR --master-port 29810 bench/fft_benchmark.py --grid 1024 --axis-contiguous > gpurun_out/fft_ac_${N}.log 2>&1; grep '^{' gpurun_out/fft_ac_${N}.log || tail -5 gpurun_out/fft_ac_${N}.log
echo "== halo_benchmark 2048x2048x1024 float, 1xN, halo 2"
$TR --master-port 29820 bench/halo_benchmark.py > gpurun_out/halo_${N}.log 2>&1; grep '^{' gpurun_out/halo_${N}.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l)
    for c in d['calls']: print(c['pencil'], 'dim', c['dim'], c['path'], round(c['us'],1), 'us', round(c['moved_bytes']/1e6,1), 'MB', round(c['gbs'] or 0,1), 'GB/s')" || tail -5 gpurun_out/halo_${N}.log
echo "== autotune 768^3"
$TR --master-port 29900 scripts/autotune_bench.py --grid 768 --backend > gpurun_out/autotune_768_${N}.log 2>&1; grep -E "SELECTED|\"autotune\"" gpurun_out/autotune_768_${N}.log
