"""qgemm micro-bench: the four GEMMs of a TinyLlama block at batch 8 x seq 1024, default dispatch, with / without tail-wave splitting."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
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
    for _ in range(3): Kn.qgemm(a, b, rowsum, sxw, ow, c0, mode, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): Kn.qgemm(a, b, rowsum, sxw, ow, c0, mode, **kw)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print(f"  N={N:6d} K={K:5d} mode={mode}: {ms*1e3:7.1f} us  {2*M*N*K/ms/1e9:7.1f} TOP/s", flush=True)
    return ms, 2.0 * M * N * K


for split in (sys.argv[1:] or ["1", "0"]):
    os.environ["MQ_QGEMM_SPLIT"] = split
    print("== tail-wave splitting", "on" if split != "0" else "off", flush=True)
    tot_ms, tot_ops = 0.0, 0.0
    for N, K, mode in ((2560, 2048, Kn.EPI_QUANT), (2048, 2048, Kn.EPI_RESID), (11264, 2048, Kn.EPI_ACTMUL), (2048, 5632, Kn.EPI_RESID)):
        ms, ops = run(8192, N, K, mode)
        tot_ms += ms; tot_ops += ops
    print(f"  block total {tot_ms*1e3:.1f} us -> {tot_ops/tot_ms/1e9:.1f} TOP/s", flush=True)
