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


@pytest.mark.parametrize("rows,K,Ho,act,out_bf16", [(1000, 128, 128, 3, 1), (257, 32, 64, 0, 0), (128 * 300 + 5, 64, 128, 0, 1)])
def test_linear_tc_matches_bf16_rounded_reference(rows, K, Ho, act, out_bf16):
    """gcm_linear_tc (tcgen05 bf16 MMA, fp32 accumulate): against a float64 product of the bf16-rounded operands."""
    from gcm import _cabi

    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(rows)
    X = (0.5 * torch.randn(rows, K, generator=g)).to(dev)
    W = (torch.randn(Ho, K, generator=g) / K ** 0.5).to(dev)
    bias = torch.randn(Ho, generator=g).to(dev)
    out = torch.empty(rows, Ho, device=dev, dtype=torch.bfloat16 if out_bf16 else torch.float32)
    _cabi.check(_cabi.lib().gcm_linear_tc(X.data_ptr(), K, K, W.data_ptr(), bias.data_ptr(), act, rows, Ho, out.data_ptr(),
                                          Ho, out_bf16, _cabi.stream_ptr(dev)), "gcm_linear_tc")
    ref = X.bfloat16().double() @ W.bfloat16().double().t() + bias.double()
    if act == 3:
        ref = torch.exp(2 * ref.clamp(-40, 40))
    tol = 1e-2 if out_bf16 else 1e-5
    assert float(((out.double() - ref).abs() / ref.abs().clamp(min=1.0)).max()) < tol


@pytest.mark.parametrize("rows,Ho,Hi,bias", [(5000, 128, 128, True), (64 * 700 + 13, 48, 32, False), (100, 128, 64, True)])
def test_outer_reduce_tc_matches_bf16_rounded_reference(rows, Ho, Hi, bias):
    """gcm_outer_reduce_tc: dW += A^T X with bf16 operands (exact products, fp32 accumulation), db += column sums
    of A in fp32; accumulates into its outputs; deterministic (two runs are bit-identical)."""
    from gcm import _cabi

    dev = torch.device("cuda:0")
    lib = _cabi.lib()
    g = torch.Generator().manual_seed(rows)
    A = torch.randn(rows, Ho, generator=g).to(dev)
    X = torch.randn(rows, Hi, generator=g).to(dev)
    ws = torch.empty(int(lib.gcm_outer_reduce_tc_workspace(rows)), device=dev)
    outs = []
    for _ in range(2):
        dW = torch.ones(Ho, Hi, device=dev)
        db = torch.ones(Ho, device=dev)
        _cabi.check(lib.gcm_outer_reduce_tc(A.data_ptr(), Ho, Ho, X.data_ptr(), Hi, Hi, rows, ws.data_ptr(), dW.data_ptr(),
                                            db.data_ptr() if bias else None, _cabi.stream_ptr(dev)), "gcm_outer_reduce_tc")
        outs.append((dW, db))
    ref = 1 + A.bfloat16().double().t() @ X.bfloat16().double()
    scale = float(ref.abs().max())
    assert float((outs[0][0].double() - ref).abs().max()) < 2e-5 * scale
    if bias:
        assert float((outs[0][1].double() - (1 + A.double().sum(0))).abs().max()) < 1e-4 * max(1.0, rows ** 0.5)
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])


@pytest.mark.parametrize("rows,K1,K2,Ho,act", [(16384, 128, 128, 128, 1), (300, 64, 0, 32, 3), (129, 16, 32, 128, 0)])
def test_linear_tc32_is_fp32_accurate(rows, K1, K2, Ho, act):
    """gcm_linear_tc32: the X1 product in 3xTF32 must be fp32-class (against fp64), the optional X2 product is bf16."""
    from gcm import _cabi

    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(rows + K1)
    X1 = (8.0 * torch.randn(rows, K1, generator=g)).to(dev)           # large magnitudes, like the window sums S and G
    W1 = (torch.randn(Ho, K1, generator=g) / K1 ** 0.5).to(dev)
    X2 = torch.tanh(torch.randn(rows, max(K2, 1), generator=g)).to(dev)
    W2 = (torch.randn(Ho, max(K2, 1), generator=g) / max(K2, 1) ** 0.5).to(dev)
    bias = torch.randn(Ho, generator=g).to(dev)
    out = torch.empty(rows, Ho, device=dev)
    status = torch.zeros(2, dtype=torch.int32, device=dev)
    _cabi.check(_cabi.lib().gcm_linear_tc32(X1.data_ptr(), K1, K1, W1.data_ptr(), X2.data_ptr() if K2 else None, K2, K2,
                                            W2.data_ptr() if K2 else None, bias.data_ptr(), act, rows, Ho, out.data_ptr(), Ho,
                                            status.data_ptr(), _cabi.stream_ptr(dev)), "gcm_linear_tc32")
    z = X1.double() @ W1.double().t() + bias.double()
    if K2:
        z = z + X2.bfloat16().double() @ W2.bfloat16().double().t()
    z32 = (X1 @ W1.t()).double() + bias.double() + ((X2.bfloat16().float() @ W2.bfloat16().float().t()).double() if K2 else 0)
    if act == 1:
        ref, ref32 = torch.tanh(z), torch.tanh(z32)
    elif act == 3:
        ref, ref32 = torch.exp(2 * z.clamp(-40, 40)), torch.exp(2 * z32.clamp(-40, 40))
    else:
        ref, ref32 = z, z32
    rel = lambda x: float(((x - ref).abs() / ref.abs().clamp(min=1.0)).max())
    err, err32 = rel(out.double()), rel(ref32)
    print(f"rows={rows} K1={K1} K2={K2} Ho={Ho} act={act}: 3xTF32 err {err:.2e}, fp32 matmul err {err32:.2e}")
    assert err < (1e-4 if act == 3 else 1e-5) + 2 * err32      # exp(2z) multiplies a relative error of z by 2|z|
    assert int(status[0]) == 0


@pytest.mark.parametrize("K,N", [(64, 64), (128, 64), (32, 128)])
def test_tc_block_ss_form_3xtf32(K, N):
    """Same product with BOTH operands in shared memory (canonical K-major descriptors for A as for B): the form the
    sparse GraphConv tile uses, where warps gather rows straight into the A tile."""
    from gcm import _cabi

    dev = torch.device("cuda:0")
    gen = torch.Generator().manual_seed(K * 7 + N)
    A = torch.randn(128, K, generator=gen).to(dev)
    B = (torch.randn(N, K, generator=gen) / K ** 0.5).to(dev)
    D = torch.zeros(128, N, device=dev)
    _cabi.check(_cabi.lib().gcm_tc_selftest(A.data_ptr(), B.data_ptr(), D.data_ptr(), K, N, 5, _cabi.stream_ptr(dev)),
                "gcm_tc_selftest")
    torch.cuda.synchronize()
    ref = A.double() @ B.double().t()
    assert float((D.double() - ref).abs().max() / ref.abs().max()) < 2e-6


@pytest.mark.parametrize("rows,Ho,Hi", [(64 * 900 + 7, 64, 64), (5000, 128, 128), (200, 48, 32)])
def test_outer_reduce_tc32_is_fp32_accurate(rows, Ho, Hi):
    """gcm_outer_reduce_tc32 (3xTF32): dW += A^T X against float64, with the error of a plain fp32 matmul as the budget."""
    from gcm import _cabi

    dev = torch.device("cuda:0")
    lib = _cabi.lib()
    g = torch.Generator().manual_seed(rows + Ho)
    A = torch.randn(rows, Ho, generator=g).to(dev)
    X = torch.randn(rows, Hi, generator=g).to(dev)
    ws = torch.empty(int(lib.gcm_outer_reduce_tc_workspace(rows)), device=dev)
    dW = torch.zeros(Ho, Hi, device=dev)
    db = torch.zeros(Ho, device=dev)
    _cabi.check(lib.gcm_outer_reduce_tc32(A.data_ptr(), Ho, Ho, X.data_ptr(), Hi, Hi, rows, ws.data_ptr(), dW.data_ptr(),
                                          db.data_ptr(), _cabi.stream_ptr(dev)), "gcm_outer_reduce_tc32")
    ref = A.double().t() @ X.double()
    scale = float(ref.abs().max())
    err = float((dW.double() - ref).abs().max()) / scale
    err32 = float(((A.t() @ X).double() - ref).abs().max()) / scale
    print(f"rows={rows}: 3xTF32 err {err:.2e}, fp32 matmul err {err32:.2e}")
    assert err < 2e-6 + 2 * err32
    assert float((db.double() - A.double().sum(0)).abs().max()) < 1e-4 * max(1.0, rows ** 0.5)

