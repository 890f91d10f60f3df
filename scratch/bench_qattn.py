"""qattn micro-bench: tcgen05 kernel vs the two mma.sync kernels, CUDA events.  usage: bench_qattn.py [impl,...] [iters]"""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mobilequant_b200 import kernels as K
from oracle import int_ref as ir
cuda = torch.device("cuda:0")
f32 = np.float32
impls = sys.argv[1].split(",") if len(sys.argv) > 1 else ["tc", "smem"]
iters0 = int(sys.argv[2]) if len(sys.argv) > 2 else 10
shapes = [tuple(int(x) for x in s.split("x")) for s in sys.argv[3].split(",")] if len(sys.argv) > 3 else \
    [(8, 1024, 32, 4, 64), (32, 1024, 32, 4, 64), (8, 1024, 32, 32, 64), (2, 2048, 32, 4, 64), (1, 4096, 32, 4, 64), (4, 1024, 16, 4, 128)]


def run(B, T, nh, nkv, hd, iters):
    q = torch.randint(0, 256, (B, nh, T, hd), dtype=torch.uint8, device=cuda)
    k = torch.randint(0, 256, (B, nkv, T, hd), dtype=torch.uint8, device=cuda)
    vt = torch.randint(0, 256, (B, nkv, hd, T), dtype=torch.uint8, device=cuda)
    bufs = dict(q=q, k=k, vt=vt, rsq=q.to(torch.int32).sum(-1).to(torch.int32), rsk=k.to(torch.int32).sum(-1).to(torch.int32))
    smax = 255 * 255 * hd * 0.02 * 0.018 * 0.12
    qs = (f32(2 * smax / 65535), f32(32768), f32(65535))
    lut = torch.from_numpy(ir.exp_tables(qs[0], hd).view(np.int32)).to(cuda)
    params = [f32(126), f32(131), f32(124), f32(0.02) * f32(0.018), qs[0], qs[1], qs[2], f32(1.0 / 65535), f32(65535),
              f32(1.0 / 65535) * f32(0.015), f32(0.7 / 255), f32(128)]
    out = torch.empty(B * T, nh * hd, dtype=torch.uint8, device=cuda)
    rs = torch.zeros(B * T, dtype=torch.int32, device=cuda)
    res = {}
    for impl in impls:
        os.environ["MQB200_QATTN"] = {"tc": "tc!", "smem": "smem", "3pass": "3pass"}[impl]
        for _ in range(2): K.qattn(bufs, B, T, nh, nkv, hd, params, lut, out=out, rowsum_out=rs)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters): K.qattn(bufs, B, T, nh, nkv, hd, params, lut, out=out, rowsum_out=rs)
        e1.record(); torch.cuda.synchronize()
        res[impl] = (e0.elapsed_time(e1) / iters, out.clone())
    same = all(torch.equal(res[impls[0]][1], res[i][1]) for i in impls[1:])
    el = B * nh * T * (T + 1) / 2
    print(f"B={B} T={T} nh={nh} nkv={nkv} hd={hd}: " + "  ".join(f"{i} {res[i][0]*1e3:.1f} us ({el/res[i][0]/1e6:.1f} G scores/s)" for i in impls) + f"  same={same}", flush=True)


for s in shapes:
    run(*s, iters=iters0)
