mkdir -p gpurun_out
for layout in "" "--axis-contiguous"; do
for c in 148 222 296 370 444 592 888; do
  python bench.py --grid 512 --steps 20 --warmup 3 --no-e2e --no-cpu-baseline --ctas $c $layout > gpurun_out/sweep_$c.log 2>&1
  grep '"metric"' gpurun_out/sweep_$c.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('$layout ctas $c', round(d['value'],1), round(d['roofline']['frac'],4), {k:round(v,4) for k,v in d['roofline']['per_op_ms'].items()})"
done; done
