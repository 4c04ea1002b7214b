"""How fast is the single-plane (bf16 mode) GEMM with fp32 C vs bf16-only output?  (epilogue / store-bandwidth probe)"""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from drin_b200 import _lib  # noqa: E402

lib = _lib.load()
p = lambda t: C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def run(layout, M, N, K, planes, out):
    a = torch.randn((M, K), device="cuda").to(torch.bfloat16)
    b = torch.randn((N, K) if layout == 0 else (K, N), device="cuda").to(torch.bfloat16)
    a_lo = torch.zeros_like(a) if planes == 2 else None
    b_lo = torch.zeros_like(b) if planes == 2 else None
    c = torch.empty(M, N, device="cuda") if out in ("f32", "both") else None
    oh = torch.empty(M, N, dtype=torch.bfloat16, device="cuda") if out in ("bf16", "both") else None
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def call():
        _lib.check(lib.drin_gemm(C.c_int32(layout), p(a), p(a_lo), C.c_int32(K), p(b), p(b_lo), C.c_int32(b.shape[1]),
                                 C.c_int64(M), C.c_int32(N), C.c_int64(K), p(c), C.c_int32(N), None, p(oh), None,
                                 C.c_int32(N), C.c_int32(1), None, C.c_int32(0), stream), "gemm")
    for _ in range(3):
        call()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        call()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(f"layout {layout} M={M} N={N} K={K} planes={planes} out={out}: {ms * 1e3:.1f} us, "
          f"{2.0 * M * N * K / ms / 1e9:.0f} TFLOP/s algorithmic", flush=True)



for out in ("f32", "bf16", "both"):
    run(0, 90112, 768, 768, 1, out)
run(0, 45056, 768, 2048, 1, "f32")
run(1, 90112, 768, 768, 1, "f32")
for out in ("f32", "both"):
    run(0, 90112, 768, 768, 2, out)
run(0, 45056, 768, 2048, 2, "f32")
run(1, 90112, 768, 768, 2, "f32")
run(0, 8192, 768, 768, 2, "both")
