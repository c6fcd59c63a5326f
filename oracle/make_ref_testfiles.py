#!/usr/bin/env python
"""Generates the case lists the reference's tests/test_runner.py would feed to its legacy test executables
(--testfile), by importing that script from /root/reference and calling its own generator on its own
tests/test_config.yaml. Run in the build container only; outputs are committed under tests/golden/ref_cases/.

    python oracle/make_ref_testfiles.py            # 4 ranks, like the reference's CTest setup
"""
import argparse
import importlib.util
import os
import sys

REF_TESTS = "/root/reference/tests"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "ref_cases")

CONFIGS = ["transpose_test_cc", "transpose_test_halo_cc", "transpose_test_padding_cc", "transpose_test_gdimdist_cc",
           "transpose_test_mix_cc", "transpose_test_ac_cc", "transpose_test_rank_order_cc", "halo_test_cc",
           "halo_test_halomix_cc", "halo_test_padding_cc", "halo_test_gdimdist_cc", "halo_test_mix_cc",
           "halo_test_ac_cc", "halo_test_rank_order_cc"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ngpu", type=int, default=4)
    ap.add_argument("--out", default=OUT)
    ap.add_argument("--max-per-config", type=int, default=300,
                    help="keep at most this many evenly spaced cases per configuration (0 = all 40k)")
    args = ap.parse_args()
    if not os.path.isdir(REF_TESTS):
        sys.exit("reference not mounted; case lists are generated in the build container only")
    spec = importlib.util.spec_from_file_location("ref_test_runner", os.path.join(REF_TESTS, "test_runner.py"))
    runner = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(runner)
    out_dir = args.out
    os.makedirs(out_dir, exist_ok=True)
    total = 0
    index = {}
    for name in CONFIGS:
        config = runner.load_yaml_config(os.path.join(REF_TESTS, "test_config.yaml"), name)
        cmds = runner.generate_command_lines(config, args)
        path = os.path.join(out_dir, "%s.n%d.txt" % (name, args.ngpu))
        full = len(cmds)
        if args.max_per_config and full > args.max_per_config:
            step = full / float(args.max_per_config)
            cmds = [cmds[int(i * step)] for i in range(args.max_per_config)]
        with open(path, "w") as f:
            for c in cmds:
                f.write(c.strip() + "\n")
        total += len(cmds)
        exe = "halo_test" if name.startswith("halo") else "transpose_test"
        index[name] = dict(file=os.path.basename(path), executable=exe, dtypes=config["dtypes"], generated=full,
                           kept=len(cmds), nranks=args.ngpu)
        print("%-34s %5d of %5d cases  dtypes %s" % (name, len(cmds), full, config["dtypes"]))
    import json
    with open(os.path.join(out_dir, "index.n%d.json" % args.ngpu), "w") as f:
        json.dump(index, f, indent=1, sort_keys=True)
    print("total", total, "->", out_dir)


if __name__ == "__main__":
    main()
