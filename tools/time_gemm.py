import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "wiki-grx-gym_b200")]
import torch
from grx_b200 import _lib as L
lib = L.lib()
def run(variant, epi, M, N, K, splits=1, use_tc=1, iters=20):
    r = lambda *s: torch.randn(*s, device="cuda")
    if variant == 0: A, B, bias, aux, bo = r(M, K), r(N, K), r(N), None, None
    elif variant == 1: A, B, bias, aux, bo = r(M, K), r(K, N), None, r(M, N), None
    else: A, B, bias, aux, bo = r(K, M), r(K, N), None, None, torch.zeros(M, device="cuda")
    Cm = torch.zeros(M, N, device="cuda")
    p = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    f = lambda: L.check(lib.grx_gemm_debug(variant, epi, M, N, K, p(A), p(B), p(Cm), p(bias), p(aux), p(bo), splits, use_tc, st))
    for _ in range(3): f()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): f()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / iters * 1e3
    stamps = (C.c_uint64 * 16)()
    L.check(lib.grx_gemm_debug_stamps(stamps))
    st = [int(x) for x in stamps]
    rel = [round((x - st[0]) / 1e3, 2) for x in st[1:7]]
    print("   epilogue detail (us after entry): ld %s  staged %s  fenced %s  stored %s  all_tiles %s" % tuple(round((x - st[0]) / 1e3, 2) for x in st[8:13]))
    print("   CTA0 us after entry: setup %s  tma0 %s  stage0 %s  mma_end %s  acc0 %s  epi_end %s" % tuple(rel))
    print(f"dbg={os.environ.get('GRX_TC_DEBUG','0')} variant {variant} M{M} N{N} K{K} splits{splits} tc{use_tc}: {us:.1f} us  {2*M*N*K/us/1e6:.1f} TFLOP/s")
run(0, 1, 10485, 256, 512)
run(0, 0, 10485, 256, 512)
run(0, 0, 10485, 256, 32)
run(0, 1, 10485, 256, 32)
run(0, 1, 10485, 512, 168)
run(1, 2, 10485, 512, 256)
run(2, 3, 256, 512, 10485, splits=41)
run(0, 1, 4096, 256, 512)
print("--- K sweep")
for K in (32, 64, 128, 256, 512, 1024):
    run(0, 1, 10485, 256, K)
print("--- M sweep (N=128: one N tile)")
for M in (128 * 37, 128 * 74, 128 * 148, 128 * 296, 128 * 592):
    run(0, 1, M, 128, 512)
