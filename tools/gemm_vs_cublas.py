"""Kernel-level comparison of the library's tcgen05 mainloop with cuBLAS (torch.matmul) on the path's GEMM shapes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from halva_b200 import debug_gemm

def timeit(f, n=20, warm=3):
    for _ in range(warm): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

R, V, H = 16384, 32000, 4096
g = torch.Generator(device="cuda").manual_seed(0)
dz = (torch.randn(R, V, generator=g, device="cuda") * 0.01).to(torch.bfloat16)
W = (torch.randn(V, H, generator=g, device="cuda") * 0.02).to(torch.bfloat16)
Hm = torch.randn(R, H, generator=g, device="cuda").to(torch.bfloat16)
fl = 2.0 * R * V * H
# logits: [R,H] x [V,H]^T   (K-major, K-major)
t = timeit(lambda: debug_gemm(Hm, W, 0, 0, 256)); print(f"ours  logits  {t:.3f} ms {fl/t/1e9:.0f} TF/s")
t = timeit(lambda: Hm @ W.t());                     print(f"cublas logits {t:.3f} ms {fl/t/1e9:.0f} TF/s")
# dH: [R,V] x [V,H]       (A K-major, B MN-major)
t = timeit(lambda: debug_gemm(dz, W, 0, 1, 256));   print(f"ours  dH      {t:.3f} ms {fl/t/1e9:.0f} TF/s")
t = timeit(lambda: dz @ W);                         print(f"cublas dH     {t:.3f} ms {fl/t/1e9:.0f} TF/s")
# dW: [V,R] x [R,H]       (A MN-major, B MN-major)
t = timeit(lambda: debug_gemm(dz, Hm, 1, 1, 256));  print(f"ours  dW      {t:.3f} ms {fl/t/1e9:.0f} TF/s")
t = timeit(lambda: dz.t() @ Hm);                    print(f"cublas dW     {t:.3f} ms {fl/t/1e9:.0f} TF/s")
