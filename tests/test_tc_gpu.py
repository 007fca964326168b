"""GPU tier: the tcgen05/TMEM building block (A in TMEM, B in shared memory, 3xTF32) against fp64."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("K,N", [(64, 32), (32, 64), (16, 16), (128, 128)])
def test_tc_block_3xtf32_is_fp32_accurate(K, N):
    from gcm import _cabi

    dev = torch.device("cuda:0")
    gen = torch.Generator().manual_seed(K * 1000 + N)
    A = torch.randn(128, K, generator=gen).to(dev)
    B = (torch.randn(N, K, generator=gen) / K ** 0.5).to(dev)
    ref = (A.double() @ B.double().t())
    ref32 = A @ B.t()
    out = {}
    for passes in (3, 1, 4):
        D = torch.zeros(128, N, device=dev)
        _cabi.check(_cabi.lib().gcm_tc_selftest(A.data_ptr(), B.data_ptr(), D.data_ptr(), K, N, passes,
                                                _cabi.stream_ptr(dev)), "gcm_tc_selftest")
        torch.cuda.synchronize()
        out[passes] = float((D.double() - ref).abs().max() / ref.abs().max())
    err32 = float((ref32.double() - ref).abs().max() / ref.abs().max())
    print(f"K={K} N={N}: 3xTF32 err {out[3]:.2e}, bf16-lo err {out[4]:.2e}, tf32 err {out[1]:.2e}, fp32 (cuBLAS) err {err32:.2e}")
    assert out[1] < 2e-3                       # plain tf32: layouts / descriptors are right
    assert out[3] < 2e-6                       # 3xTF32: fp32-class accuracy
    assert out[4] < 2e-6                       # lo term on the bf16 path (packed A in TMEM): same class
