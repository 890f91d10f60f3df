import os, sys, torch, time
sys.path.insert(0, "/root/repo")
from mobilequant_b200 import kernels as Kn
cuda = torch.device("cuda:0")
def run(M, N, K, mode=Kn.EPI_QUANT, iters=20):
    a = torch.randint(0, 256, (M, K), dtype=torch.uint8, device=cuda)
    b = torch.randint(0, 256, (N, K), dtype=torch.uint8, device=cuda)
    rowsum = a.to(torch.int32).sum(1).to(torch.int32); sxw = torch.full((N,), 1e-5, device=cuda)
    ow = torch.full((N,), 128, dtype=torch.int32, device=cuda); c0 = torch.zeros(N, dtype=torch.int32, device=cuda)
    G = (N + 127) // 128; so = torch.full((G,), 0.05, device=cuda); oo = torch.full((G,), 128.0, device=cuda)
    lut = torch.randn(256, device=cuda)
    kw = dict(so=so, oo=oo, qmax=255.0, qgroup=128)
    if mode == Kn.EPI_ACTMUL: kw.update(lut=lut, s2=0.01, o2=128.0)
    if mode == Kn.EPI_RESID: kw.update(resid=torch.zeros(M, N, device=cuda), qmax=65535.0)
    out = None
    for _ in range(3): out = Kn.qgemm(a, b, rowsum, sxw, ow, c0, mode, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): Kn.qgemm(a, b, rowsum, sxw, ow, c0, mode, **kw)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print(f"M={M} N={N} K={K} mode={mode}: {ms*1e3:.1f} us  {2*M*N*K/ms/1e9:.1f} TOP/s")
for spec in (sys.argv[1:] or ["1/8", "1/16", "2/8", "2/16", "4/16"]):
    cl, ne = spec.split("/")
    os.environ["MQ_QGEMM_CL"] = cl; os.environ["MQ_QGEMM_NE"] = ne
    print("== cluster", cl, "epilogue warps", ne, flush=True)
    for M in (8192,):
        run(M, 2560, 2048); run(M, 2048, 2048, Kn.EPI_RESID); run(M, 11264, 2048, Kn.EPI_ACTMUL); run(M, 2048, 5632, Kn.EPI_RESID)
        run(M, 8192, 8192, Kn.EPI_I32)
# library proxy for the INT8 peak
a = torch.randint(-128, 127, (8192, 8192), dtype=torch.int8, device=cuda); b = torch.randint(-128, 127, (8192, 8192), dtype=torch.int8, device=cuda)
for _ in range(3): torch._int_mm(a, b.t())
torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): torch._int_mm(a, b.t())
e1.record(); torch.cuda.synchronize(); ms = e0.elapsed_time(e1) / 20
print(f"torch._int_mm 8192^3: {ms*1e3:.1f} us {2*8192**3/ms/1e9:.1f} TOP/s")
