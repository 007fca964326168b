import sys, torch
sys.path.insert(0, "graph-conv-memory_b200")
from gcm import _cabi
dev = torch.device("cuda:0")
gen = torch.Generator().manual_seed(77)
P = torch.randn(128, 128, generator=gen).to(dev)
Q = torch.randn(128, 64, generator=gen).to(dev)
D = torch.zeros(128, 64, device=dev)
_cabi.check(_cabi.lib().gcm_tc_selftest_mn(P.data_ptr(), Q.data_ptr(), D.data_ptr(), _cabi.stream_ptr(dev)), "x")
torch.cuda.synchronize()
ref = P.double().t() @ Q.double()
print("D abs max", float(D.abs().max()), "ref max", float(ref.abs().max()))
err = (D.double() - ref).abs()
print("err max", float(err.max()))
print("rows ok:", [int(r) for r in range(128) if float(err[r].max()) < 1e-3][:40])
print("cols ok:", [int(c) for c in range(64) if float(err[:, c].max()) < 1e-3][:40])
# try to identify which element D[m][n] equals
Pd, Qd = P.double(), Q.double()
for m in (0, 1, 4, 5, 8, 33):
    for n in (0, 1, 4, 9):
        v = float(D[m, n])
        # search over (m', n') such that ref[m', n'] ~ v
        hit = ((ref - v).abs() < 1e-3).nonzero()
        print(m, n, v, hit[:3].tolist())
