#!/bin/bash
# Host-side code of the library under AddressSanitizer + UndefinedBehaviorSanitizer (no GPU needed).
#   scripts/sanitize_host.sh [pytest args]      default: the single- and multi-process host suites
# Builds an instrumented libcudecomp.so (every .cc with g++ -fsanitize=address,undefined; the kernels stay as nvcc built
# them), swaps it in for the run, runs pytest with the sanitizer runtimes preloaded (child ranks inherit them), restores
# the release library, and lists every report found in the pytest log and in the per-rank logs.
set -u
cd "$(dirname "$0")/.."
B=/tmp/cudecomp_sanitize_build
mkdir -p $B
ARGS=("$@")
[ ${#ARGS[@]} -eq 0 ] && ARGS=(tests/test_abi.py tests/test_abi_golden.py tests/test_planner_properties.py tests/test_launch_emulation.py
                               tests/test_host_multirank.py tests/test_api_contract.py tests/test_autotune_candidates.py -q)
python -c "from cudecomp_b200.build import build_library; build_library()" || exit 1
for f in cudecomp_b200/csrc/*.cc; do
  g++ -O1 -g -std=c++17 -fPIC -w -fsanitize=address,undefined -fno-omit-frame-pointer -I/usr/local/cuda/include -Iinclude \
      -Iinclude/mpi_shim -Icudecomp_b200/csrc -c $f -o $B/$(basename $f).o &
done
wait
g++ -shared -fsanitize=address,undefined -o $B/libcudecomp.so $B/*.cc.o cudecomp_b200/build/kernels.cu.o \
    -L/usr/local/cuda/lib64 -lcudart -lrt -lpthread -ldl || exit 1
cp cudecomp_b200/lib/libcudecomp.so $B/libcudecomp.release.so
restore() { cp $B/libcudecomp.release.so cudecomp_b200/lib/libcudecomp.so; touch cudecomp_b200/lib/libcudecomp.so; }
trap restore EXIT
cp $B/libcudecomp.so cudecomp_b200/lib/libcudecomp.so
touch cudecomp_b200/lib/libcudecomp.so cudecomp_b200/lib/libcudecomp_realmpi.so
rm -rf /tmp/cdb200_*/
ASAN_OPTIONS=detect_leaks=0 UBSAN_OPTIONS=print_stacktrace=0:halt_on_error=0 \
  LD_PRELOAD=$(g++ -print-file-name=libasan.so):$(g++ -print-file-name=libubsan.so) \
  python -m pytest "${ARGS[@]}" -s -p no:cacheprovider > $B/run.log 2>&1
echo "pytest exit code: $?"
tail -2 $B/run.log
echo "sanitizer reports:"
grep -h "runtime error\|ERROR: AddressSanitizer" $B/run.log /tmp/cdb200_*/rank*.log 2>/dev/null | sort | uniq -c | sort -rn | head -40
echo "(end of reports)"
