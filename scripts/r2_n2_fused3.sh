# usage: gpurun --gpus 2 --timeout 800 -- 'bash scripts/r2_n2_fused3.sh'
# Round 2, fifth GPU call: the reworked phased launch (lag 1, head of pushes before the unpacks join, one fencing thread
# per CTA), 256-bit stores on the wire by default, tile geometry of the 8-byte transpose, and the peer-store ncu capture
# with NVLink counters.
mkdir -p gpurun_out
N=2
OUT=gpurun_out
export CUDECOMP_B200_DEVICE_TIMEOUT=20
i=0
bench() { # label, extra args...
  label=$1; shift
  i=$((i+1))
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
    --master-port $((29500+i*10)) bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --no-e2e "$@" > $OUT/r2d_n${N}_$label.log 2>&1
  grep '"metric"' $OUT/r2d_n${N}_$label.log | tee $OUT/r2d_n${N}_$label.json | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); r = d['roofline']
    print('$label:', round(d['ms_per_step'], 3), 'ms/step;', d['path'], {k: round(v, 3) for k, v in r['per_op_ms'].items()},
          r['bound'], round(r['achieved'], 1), 'GB/s frac', round(r['frac'], 4), 'parity', d.get('parity', {}).get('ok'))" || tail -5 $OUT/r2d_n${N}_$label.log
}
echo "== fused staged schedules on 4 ranks (2 per GPU): parity + back-to-back stress; NCCL cross-check"
timeout 400 python -m pytest tests/test_zz_pipeline_gpu.py tests/test_zz_nccl_crosscheck_gpu.py -q -m gpu -x -p no:cacheprovider > $OUT/r2d_pipeline_tests.log 2>&1
tail -4 $OUT/r2d_pipeline_tests.log
echo "== in place, 1x2"
bench inplace_fused --inplace
bench inplace_fused_k32 --inplace --chunks 32
bench inplace_fused_k8 --inplace --chunks 8
bench inplace_fused_head0 --inplace --phase-head 0
bench inplace_fused_head50 --inplace --phase-head 50
bench inplace_fused_lag2 --inplace --lag 2
bench inplace_fused_128bit --inplace --no-wire-wide
bench inplace_fused_tile16k --inplace --tile-bytes 16384
echo "== in place, 2x1"
bench inplace_fused_2x1 --inplace --pdims 2x1
echo "== out of place"
bench default
bench default_128bit --no-wire-wide
bench default_2x1 --pdims 2x1
echo "== 512^3 complex64"
bench c64_512 --grid 512 --dtype float_complex
bench c64_512_inplace --grid 512 --dtype float_complex --inplace
bench c64_512_inplace_k4 --grid 512 --dtype float_complex --inplace --chunks 4
bench c64_512_inplace_k8 --grid 512 --dtype float_complex --inplace --chunks 8
echo "== 8-byte vectorised transpose: tile geometry"
bench ac_f64_geom0 --axis-contiguous --dtype double
CUDECOMP_B200_TRANSPOSE_GEOM=1 bench ac_f64_geom1 --axis-contiguous --dtype double
echo "== ncu: the product's row-copy kernel storing into the peer GPU (single process, no handshake), with NVLink counters"
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -I include -I include/mpi_shim -I cudecomp_b200/csrc \
  bench/microbench_handshake.cu cudecomp_b200/csrc/launch_params.cc cudecomp_b200/csrc/plan.cc cudecomp_b200/csrc/geometry.cc -o /tmp/hs || exit 1
/tmp/hs --profile-remote > $OUT/r2d_peer_copy_times.txt 2>&1; cat $OUT/r2d_peer_copy_times.txt
NVL=nvltx__bytes.sum,nvltx__bytes_data_user.sum,nvltx__bytes_data_protocol.sum,nvlrx__bytes.sum,nvlrx__bytes_data_user.sum,nvlrx__bytes_data_protocol.sum,nvltx__bytes.sum.per_second,nvlrx__bytes.sum.per_second
timeout 600 ncu --set full --metrics $NVL --clock-control none --import-source on -k regex:rowCopy -c 12 -f -o $OUT/r2d_rowcopy_peer_full /tmp/hs --profile-remote > $OUT/r2d_ncu_peer.log 2>&1
tail -3 $OUT/r2d_ncu_peer.log
if [ -f $OUT/r2d_rowcopy_peer_full.ncu-rep ]; then
  ncu -i $OUT/r2d_rowcopy_peer_full.ncu-rep --page raw --csv > $OUT/r2d_rowcopy_peer_full.raw.csv 2>/dev/null
  [ $(stat -c %s $OUT/r2d_rowcopy_peer_full.ncu-rep) -gt 25000000 ] && rm -f $OUT/r2d_rowcopy_peer_full.ncu-rep
fi
echo "== handshake cost"
/tmp/hs > $OUT/r2d_handshake_microbench.txt 2>&1; cat $OUT/r2d_handshake_microbench.txt
