"""Record the reference-order (multithreaded build arithmetic) results of the BASELINE configurations at FULL size with
the CPU oracle (oracle/smm_oracle.c, pinned bit-exact against the real reference by tests/test_oracle_pinned.py).
Takes ~20 minutes and ~12 GB for config 5; run in the build container:
    python tests/golden/make_fullsize_golden.py [1 2 2s 3 5]
Writes tests/golden/fullsize_reference.json (iteration counts, the solver's residual as float bits, an order-independent
checksum of x).  tests/test_gpu_fullsize.py compares the GPU's REFERENCE_TREE mode with it, bit for bit."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import matgen  # noqa: E402
import oracle_lib as ol  # noqa: E402

OUT = os.path.join(HERE, "fullsize_reference.json")


def stencil(nx, ny, nz, use_z, lo, diag, hi):
    L = ol.oracle()
    L.smm_oracle_stencil_nnz.restype = C.c_int64
    L.smm_oracle_stencil_nnz.argtypes = [C.c_int] * 4
    L.smm_oracle_gen_stencil.argtypes = [C.c_int] * 4 + [C.c_float] * 3 + [C.c_void_p] * 3
    rows = nx * ny * nz
    nnz = L.smm_oracle_stencil_nnz(nx, ny, nz, use_z)
    st = np.zeros(rows + 1, np.int32); po = np.zeros(nnz, np.int32); va = np.zeros(nnz, np.float32)
    L.smm_oracle_gen_stencil(nx, ny, nz, use_z, lo, diag, hi, st.ctypes.data_as(C.c_void_p), po.ctypes.data_as(C.c_void_p), va.ctypes.data_as(C.c_void_p))
    return ol.CSR(rows, rows, st, po, va, 0)


def checksum(x):
    bits = x.view(np.uint32).astype(np.uint64)
    return {"sum_bits": int(bits.sum() & np.uint64(0xFFFFFFFFFFFFFFFF)), "xor_bits": int(np.bitwise_xor.reduce(bits)),
            "sum_f64": float(x.astype(np.float64).sum())}


CONFIGS = {
    "1": dict(name="CG, 2D 5-point Poisson 1024^2, eps 1e-6, b=A*1", solver="cg", pre=0, rhs="ones", eps=1e-6,
              make=lambda: stencil(1024, 1024, 1, 0, -1.0, 4.0, -1.0)),
    "2": dict(name="BiCGStab, convdiff3d 128^3 c=0.5, eps 1e-6, b=A*x*", solver="bicgstab", pre=0, rhs="xstar", eps=1e-6,
              make=lambda: stencil(128, 128, 128, 1, -1.5, 6.0, -0.5)),
    "2s": dict(name="BiCGStab+SGS, convdiff3d 128^3 c=0.5, eps 1e-6, b=A*x*", solver="bicgstab", pre=1, rhs="xstar", eps=1e-6,
               make=lambda: stencil(128, 128, 128, 1, -1.5, 6.0, -0.5)),
    "3": dict(name="BiCGStab+SGS, convdiff3d 256^3 c=0.5, eps 1e-6, b=A*x*", solver="bicgstab", pre=1, rhs="xstar", eps=1e-6,
              make=lambda: stencil(256, 256, 256, 1, -1.5, 6.0, -0.5)),
    "5": dict(name="CG, 3D 7-point Poisson 512^3, eps 1e-6, b=A*1", solver="cg", pre=0, rhs="ones", eps=1e-6,
              make=lambda: stencil(512, 512, 512, 1, -1.0, 6.0, -1.0)),
}


def main():
    todo = sys.argv[1:] or ["1", "2", "2s"]
    data = json.load(open(OUT)) if os.path.exists(OUT) else {}
    for key in todo:
        cfg = CONFIGS[key]
        m = cfg["make"]()
        xs = np.ones(m.rows, np.float32) if cfg["rhs"] == "ones" else matgen.xstar(m.rows)
        b = ol.spmv(m, 0, None, xs)
        t = time.time()
        o = ol.solve(cfg["solver"], m, b, np.zeros(m.rows, np.float32), -1, cfg["eps"], 1, precond=cfg["pre"])
        rec = {"name": cfg["name"], "rows": m.rows, "nnz": m.nnz, "status": o["status"], "iterations": o["iterations"],
               "residual_bits": int(np.float32(o["residual"]).view(np.uint32)), "residual": o["residual"],
               "max_abs_error": float(np.max(np.abs(o["x"] - xs))), "x": checksum(o["x"]),
               "oracle_seconds": time.time() - t, "oracle_threads": ol.oracle().smm_oracle_threads()}
        data[key] = rec
        print(key, json.dumps(rec), flush=True)
        json.dump(data, open(OUT, "w"), indent=1)


if __name__ == "__main__":
    main()
