"""tcgen05 mainloop unit tests: D = A B^T through libhalva_dpa's debug hook vs torch fp32 matmul."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _mk(M, N, K, a_major, b_major, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    A = torch.randn(M, K, generator=g, device="cuda").to(torch.bfloat16)
    B = torch.randn(N, K, generator=g, device="cuda").to(torch.bfloat16)
    ref = A.float() @ B.float().t()
    As = A if a_major == 0 else A.t().contiguous()
    Bs = B if b_major == 0 else B.t().contiguous()
    return As, Bs, ref


@pytest.mark.parametrize("cg", [2, 1])
@pytest.mark.parametrize("a_major,b_major,bn", [(0, 0, 256), (0, 0, 128), (0, 1, 256), (1, 1, 256), (1, 0, 256)])
@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (256, 512, 256), (384, 320, 200), (1000, 1024, 4096)])
def test_debug_gemm(a_major, b_major, bn, M, N, K, cg):
    """cg = tcgen05 cta_group of the 256-wide tiles: 2 = CTA pairs sharing one M=256 MMA (default), 1 = single CTA."""
    from halva_b200 import _lib, debug_gemm
    if (a_major == 1 and M % 8) or (b_major == 1 and N % 8):
        pytest.skip("alignment")
    if bn == 128 and cg == 1:
        pytest.skip("the 128-wide (dual-accumulator) tile shape is cta_group 1 only: covered by cg=2's run")
    As, Bs, ref = _mk(M, N, K, a_major, b_major)
    _lib.set_cta_group(cg)
    try:
        D = debug_gemm(As, Bs, a_major, b_major, bn)
    finally:
        _lib.set_cta_group(2)
    torch.cuda.synchronize()
    err = (D - ref).abs().max().item()
    scale = ref.abs().max().item()
    assert err <= 2e-3 * scale + 1e-3, f"max abs err {err} (scale {scale})"


@pytest.mark.parametrize("a_major,b_major", [(0, 0), (0, 1), (1, 1), (1, 0)])
@pytest.mark.parametrize("M,N,K", [(256, 512, 64), (384, 1024, 328), (1000, 1312, 4096), (2304, 4096, 1024)])
def test_debug_gemm_512_wide_tiles(a_major, b_major, M, N, K):
    """The 256 x 512 tile of the long-K gradient GEMMs: one accumulator stage of all 512 TMEM columns, two N = 256 MMAs
    per K step, CTA pairs (incl. ragged M / N / K edges and several tiles per CTA)."""
    from halva_b200 import debug_gemm
    if (a_major == 1 and M % 8) or (b_major == 1 and N % 8):
        pytest.skip("alignment")
    As, Bs, ref = _mk(M, N, K, a_major, b_major, seed=3)
    D = debug_gemm(As, Bs, a_major, b_major, 512)
    torch.cuda.synchronize()
    err = (D - ref).abs().max().item()
    scale = ref.abs().max().item()
    assert err <= 2e-3 * scale + 1e-3, f"max abs err {err} (scale {scale})"
