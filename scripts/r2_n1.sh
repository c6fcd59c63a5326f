# usage: gpurun --timeout 1500 -- 'bash scripts/r2_n1.sh'
# Round 2, 1-GPU call: the whole GPU suite (4-rank cases share the GPU), the driver's bench lines, the permuting kernels,
# and ncu evidence of the HEAD binary: launch lists + one --set full capture per hot kernel.
mkdir -p gpurun_out
OUT=gpurun_out
git_rev=$(cat .git_rev 2>/dev/null || echo unknown)
echo "== GPU suite"
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider -rfEs > $OUT/r2_n1_gpu_suite.log 2>&1
tail -12 $OUT/r2_n1_gpu_suite.log
echo "== smoke under ncu (what the driver's launch-list step runs)"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1000 --csv --log-file $OUT/r2_n1_launches_smoke.csv \
  python -c "import __graft_entry__ as g; g.smoke()" > $OUT/r2_n1_smoke_ncu.log 2>&1; echo "ncu_rc=$?"; tail -2 $OUT/r2_n1_smoke_ncu.log
bench() { # label, args...
  label=$1; shift
  timeout 600 python bench.py --gpus 1 "$@" > $OUT/r2_n1_$label.log 2>&1
  grep '"metric"' $OUT/r2_n1_$label.log | tee $OUT/r2_n1_$label.json | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); r = d.get('roofline') or {}
    print('$label:', round(d['ms_per_step'], 3), 'ms/step;', d.get('path'), {k: round(v, 3) for k, v in (r.get('per_op_ms') or {}).items()},
          r.get('bound'), round(r.get('achieved') or 0, 1), 'frac', round(r.get('frac') or 0, 4), 'parity', (d.get('parity') or {}).get('ok'),
          'e2e', round((d.get('e2e') or {}).get('value') or 0, 1), 'cpu', (d.get('cpu_baseline') or {}).get('value'))" || tail -5 $OUT/r2_n1_$label.log
}
echo "== the driver's two lines"
( time python bench.py --impl reference --gpus 1 --steps 10 --warmup 3 ) > $OUT/r2_n1_reference_arm.log 2>&1; tail -4 $OUT/r2_n1_reference_arm.log | cut -c1-600
( time python bench.py ) > $OUT/r2_n1_default_full.log 2>&1; grep '"metric"' $OUT/r2_n1_default_full.log > $OUT/r2_n1_default_full.json; tail -4 $OUT/r2_n1_default_full.log | cut -c1-300
Q="--steps 10 --warmup 3 --no-e2e --no-cpu-baseline"
bench default $Q
bench wide $Q --wide
bench bulk $Q --bulk
bench inplace $Q --inplace
bench ac_c128 $Q --axis-contiguous
bench ac_c128_elementwise $Q --axis-contiguous --kernel-variant 3
bench ac_c128_inplace $Q --axis-contiguous --inplace
bench ac_f32 $Q --axis-contiguous --dtype float
bench ac_f32_elementwise $Q --axis-contiguous --dtype float --kernel-variant 3
bench ac_f64 $Q --axis-contiguous --dtype double
bench c64_512 $Q --grid 512 --dtype float_complex
bench c64_512_ac $Q --grid 512 --dtype float_complex --axis-contiguous
echo "== ncu launch lists"
P="--steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-parity"
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/r2_n1_launches.csv python bench.py $P > $OUT/r2_n1_ncu_list.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/r2_n1_launches_ac.csv python bench.py $P --axis-contiguous > $OUT/r2_n1_ncu_list_ac.log 2>&1
echo "== ncu --set full, one launch per hot kernel"
P1="--steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-parity"
full() { # name, kernel regex, bench args...
  name=$1; k=$2; shift; shift
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 12 -c 1 -f -o $OUT/r2_n1_${name}_full python bench.py $P1 "$@" > $OUT/r2_n1_ncu_${name}.log 2>&1
  rep=$OUT/r2_n1_${name}_full.ncu-rep
  if [ -f $rep ]; then
    ncu -i $rep --page raw --csv > $OUT/r2_n1_${name}_full.raw.csv 2>/dev/null
    ncu -i $rep --page details > $OUT/r2_n1_${name}_full.details.txt 2>/dev/null
    [ $(stat -c %s $rep) -gt 12000000 ] && rm -f $rep
    grep -E "dram__bytes_(read|write).sum|gpu__time_duration.sum" $OUT/r2_n1_${name}_full.raw.csv | head -2 >/dev/null
    echo "captured $name"
  else
    tail -3 $OUT/r2_n1_ncu_${name}.log
  fi
}
full rowcopy "rowCopyKernel"
full transposevec_c128 "transposeVecKernel" --axis-contiguous
full transposevec_f32 "transposeVecKernel" --axis-contiguous --dtype float
full transpose_elementwise_f32 "transposeKernel" --axis-contiguous --dtype float --kernel-variant 3
full bulk "rowCopyBulkKernel" --bulk
ls -la $OUT | grep r2_n1 | head -60
