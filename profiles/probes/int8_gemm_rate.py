"""What the int8 tensor cores of this GPU deliver on a plain library GEMM (cuBLASLt through torch._int_mm), as the
yardstick for a full-mode formulation as one u8 GEMM over tip pairs (DESIGN.md, round-2 list).  Also fp8 via
torch._scaled_mm when available.  usage: python profiles/probes/int8_gemm_rate.py"""
import torch
dev = torch.device("cuda", 0)
for n in (4096, 8192, 16384):
    a = torch.randint(-2, 2, (n, n), dtype=torch.int8, device=dev)
    b = torch.randint(-2, 2, (n, n), dtype=torch.int8, device=dev)
    for _ in range(2):
        c = torch._int_mm(a, b)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); c = torch._int_mm(a, b); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print("int8 GEMM %5d^3: %.3f ms = %.0f TOP/s" % (n, best, 2.0 * n ** 3 / best / 1e9), flush=True)
try:
    n = 8192
    a = torch.randn(n, n, device=dev).to(torch.float8_e4m3fn)
    b = torch.randn(n, n, device=dev).to(torch.float8_e4m3fn).t()
    one = torch.tensor(1.0, device=dev)
    for _ in range(2):
        c = torch._scaled_mm(a, b, scale_a=one, scale_b=one, out_dtype=torch.bfloat16)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); c = torch._scaled_mm(a, b, scale_a=one, scale_b=one, out_dtype=torch.bfloat16); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print("fp8  GEMM %5d^3: %.3f ms = %.0f TFLOP/s" % (n, best, 2.0 * n ** 3 / best / 1e9), flush=True)
except Exception as exc:  # noqa: BLE001
    print("fp8 GEMM not available:", exc)
