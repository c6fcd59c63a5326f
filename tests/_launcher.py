"""Runs a list of test cases on N ranks (one OS process per rank, like the reference's mpiexec -n 4 runs,
tests/ctest/CMakeLists.txt:30-33). Ranks rendezvous through the library's own bootstrap on 127.0.0.1."""
import json
import os
import signal
import socket
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def free_port():
    for _ in range(50):
        with socket.socket() as s:
            s.bind(("127.0.0.1", 0))
            port = s.getsockname()[1]
        # the bootstrap listens on port+1 by convention (MASTER_PORT + 1); make sure that one is free too
        with socket.socket() as s2:
            try:
                s2.bind(("127.0.0.1", port + 1))
                return port
            except OSError:
                continue
    raise RuntimeError("no free port pair")


def run_ranks(nranks, mode, cases, timeout=600, extra_env=None):
    """mode: 'gpu' (run through the C ABI on the GPU) or 'plan' (host-only planning dump).
    Returns results[rank] = list of per-case dicts."""
    tmp = tempfile.mkdtemp(prefix="cdb200_")
    payload = os.path.join(tmp, "payload.json")
    with open(payload, "w") as f:
        json.dump(dict(mode=mode, cases=cases), f)
    port = free_port()
    procs = []
    for r in range(nranks):
        env = dict(os.environ)
        env.update(RANK=str(r), WORLD_SIZE=str(nranks), LOCAL_RANK=str(r), MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port), PYTHONPATH=ROOT + os.pathsep + env.get("PYTHONPATH", ""))
        env.setdefault("CUDECOMP_B200_DEVICE_TIMEOUT", "60")
        env.setdefault("CUDECOMP_B200_HOST_TIMEOUT", "120")
        if extra_env:
            env.update(extra_env)
        log = open(os.path.join(tmp, "rank%d.log" % r), "w")
        procs.append((subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "_worker.py"), payload, tmp],
                                       env=env, stdout=log, stderr=subprocess.STDOUT, start_new_session=True), log))
    failed = None
    try:
        for r, (p, log) in enumerate(procs):
            try:
                rc = p.wait(timeout=timeout)
            except subprocess.TimeoutExpired:
                failed = "rank %d timed out after %ds" % (r, timeout)
                break
            if rc != 0 and failed is None:
                failed = "rank %d exited with code %d" % (r, rc)
    finally:
        for p, log in procs:
            if p.poll() is None:
                try:
                    os.killpg(p.pid, signal.SIGKILL)  # exactly the process group this launcher started
                except ProcessLookupError:
                    pass
                p.wait()
            log.close()
    logs = {}
    for r in range(nranks):
        with open(os.path.join(tmp, "rank%d.log" % r)) as f:
            logs[r] = f.read()
    if failed:
        raise RuntimeError(failed + "\n" + "\n".join("--- rank %d log ---\n%s" % (r, logs[r][-4000:]) for r in logs))
    results = []
    for r in range(nranks):
        with open(os.path.join(tmp, "rank%d.json" % r)) as f:
            results.append(json.load(f))
    return results, logs
