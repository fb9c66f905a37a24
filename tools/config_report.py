#!/usr/bin/env python
"""Run the BASELINE.json configurations at full size on one B200 and write the report bench.py embeds in its N = 1 line
(`configs`): iterations, rate, SURVEY 8(d) roofline fraction per mode, bit-exact comparison with the golden record in the
reference-order mode.  Byte model: bench.bytes_per_iteration (B_sgs = 8 nnz + 32 n per apply).

    python tools/config_report.py [--configs 1,2,3,4,5] [--modes fast,tree] [--out profiles/r02_config_report.json]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import bench  # noqa: E402
import sparse_matrix_math_b200 as smm  # noqa: E402
from sparse_matrix_math_b200 import binding as B  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="1,2,3,4,5")
    ap.add_argument("--modes", default="fast,tree")
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r02_config_report.json"))
    args = ap.parse_args()
    todo = [int(c) for c in args.configs.split(",")]
    modes = tuple(args.modes.split(","))
    peak, src = bench.measured_peak_gbs()
    L = smm.lib()
    report = {"device": smm.device_info(), "peak_gbs": peak, "peak_source": src,
              "configs": bench.configs_report(smm, B, L, peak, which=[c for c in todo if c != 5], modes=modes)}
    if 5 in todo:
        A = smm.CSRMatrix.generate(B.GEN_CONVDIFF3D, 512, 512, 512, 0.0)
        report["configs"].append(bench.run_config(smm, B, L, "5: ConjugateGradient, 3D 7-point Poisson 512^3, eps 1e-6, b=A*1", "5", "cg", A, None, "ones",
                                                  1e-6, -1, modes, peak, bench.golden_fullsize()))
        del A
    for c in report["configs"]:
        for r in c["runs"]:
            print(json.dumps({"config": c["config"], **r}), flush=True)
    json.dump(report, open(args.out, "w"), indent=1)
    print("wrote", args.out)


if __name__ == "__main__":
    main()
