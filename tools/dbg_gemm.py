import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "wiki-grx-gym_b200"), os.path.join(ROOT, "tests")]
import torch
from test_gemm_gpu import _run
for (v, e, M, N, K, sp) in [(1, 2, 128, 64, 64, 1), (1, 2, 128, 16, 8, 1), (2, 3, 128, 64, 64, 1), (2, 3, 128, 16, 32, 1), (1, 2, 10485, 512, 256, 1)]:
    try:
        got, ref, extra = _run(v, e, M, N, K, use_tc=1, splits=sp)
        err = (got - ref).abs()
        print(f"variant {v} M{M} N{N} K{K}: max err {float(err.max()):.4g} scale {float(ref.abs().max()):.4g} nz(got) {int((got != 0).sum())}/{got.numel()}")
        if float(err.max()) > 0.01 * float(ref.abs().max()):
            print(" got[0,:8]", got[0, :8].tolist()); print(" ref[0,:8]", ref[0, :8].tolist())
            print(" got[5,:8]", got[5, :8].tolist()); print(" ref[5,:8]", ref[5, :8].tolist())
    except Exception as ex:
        print("variant", v, M, N, K, "EXC", ex)
