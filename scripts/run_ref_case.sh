# usage: run_ref_case.sh <nranks> <exe> <args...>   (poor man's mpirun over the library's bootstrap)
N=$1; shift
PORT=$((20000 + RANDOM % 20000))
for r in $(seq 1 $((N-1))); do
  RANK=$r WORLD_SIZE=$N MASTER_ADDR=127.0.0.1 MASTER_PORT=$PORT "$@" > /dev/null 2>&1 &
done
RANK=0 WORLD_SIZE=$N MASTER_ADDR=127.0.0.1 MASTER_PORT=$PORT "$@"
rc=$?
wait
exit $rc
