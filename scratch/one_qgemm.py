import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mobilequant_b200 import kernels as Kn
cuda = torch.device("cuda:0")
M, N, K = 8192, 11264, 2048
a = torch.randint(0, 256, (M, K), dtype=torch.uint8, device=cuda)
b = torch.randint(0, 256, (N, K), dtype=torch.uint8, device=cuda)
rowsum = a.to(torch.int32).sum(1).to(torch.int32); sxw = torch.full((N,), 1e-5, device=cuda)
ow = torch.full((N,), 128, dtype=torch.int32, device=cuda); c0 = torch.zeros(N, dtype=torch.int32, device=cuda)
G = (N + 127) // 128; so = torch.full((G,), 0.05, device=cuda); oo = torch.full((G,), 128.0, device=cuda)
lut = torch.randn(256, device=cuda)
rs = torch.zeros(M, dtype=torch.int32, device=cuda)
for _ in range(3):
    Kn.qgemm(a, b, rowsum, sxw, ow, c0, Kn.EPI_ACTMUL, so=so, oo=oo, qmax=255.0, qgroup=128, lut=lut, s2=0.01, o2=128.0, rowsum_out=rs)
torch.cuda.synchronize()
