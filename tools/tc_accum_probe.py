"""How exact is the tensor core's fp32 accumulation?  D = A B^T (bf16-valued operands, K = 4096) through the library's
tcgen05 mainloop and through torch's fp32 SIMT matmul, both against fp64: signed bias (towards zero = truncation) and rms."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from halva_b200.dpa_loss import debug_gemm  # noqa: E402

torch.backends.cuda.matmul.allow_tf32 = False
g = torch.Generator(device="cuda").manual_seed(1)
for K in (512, 4096, 16384):
    A = (1.5 * torch.randn(1024, K, device="cuda", generator=g)).to(torch.bfloat16)
    B = (0.02 * torch.randn(2048, K, device="cuda", generator=g)).to(torch.bfloat16)
    A[:, : K // 8] = A[:, : K // 8].abs()          # a planted positive component: large same-sign accumulators
    B[:, : K // 8] = B[:, : K // 8].abs() * 4
    ref = A.double() @ B.double().t()
    D = debug_gemm(A, B).double()
    T = (A.float() @ B.float().t()).double()
    for name, X in (("tcgen05", D), ("torch fp32", T)):
        e = X - ref
        ulp = torch.finfo(torch.float32).eps * ref.abs()
        print(f"K={K:6d} {name:10s} mean(e*sign(ref))/ulp {(e * ref.sign() / ulp).mean().item():+.3f}  rms(e)/ulp {(e / ulp).pow(2).mean().sqrt().item():.3f}"
              f"  mean |ref| {ref.abs().mean().item():.2f}", flush=True)
