mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -I include -I include/mpi_shim -I cudecomp_b200/csrc bench/microbench_handshake.cu -o /tmp/hs || exit 1
/tmp/hs --profile-remote
ncu --set full --clock-control none --import-source on -k regex:rowCopy -o gpurun_out/prof_rowcopy_peer_r1 /tmp/hs --profile-remote > gpurun_out/ncu_peer.log 2>&1
tail -3 gpurun_out/ncu_peer.log
TR="timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
$TR --master-port 29610 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/n2b_1.log 2>&1; grep '"metric"' gpurun_out/n2b_1.log | cut -c1-1500
$TR --master-port 29620 bench.py --gpus 2 --steps 10 --warmup 3 --inplace --no-e2e > gpurun_out/n2b_2.log 2>&1; grep '"metric"' gpurun_out/n2b_2.log | cut -c1-700
(python -m pytest tests/test_gpu_parity.py -q -m gpu -k "Autotune or fft or two") > gpurun_out/t2b.log 2>&1; tail -3 gpurun_out/t2b.log
