"""Sustained qgemm loop with NVML clock/power sampling: is the tensor path power/clock limited?"""
import os, sys, time, threading, torch, pynvml
sys.path.insert(0, "/root/repo")
from mobilequant_b200 import kernels as Kn
cuda = torch.device("cuda:0")
pynvml.nvmlInit(); h = pynvml.nvmlDeviceGetHandleByIndex(0)
samples = []; stop = False
def sampler():
    while not stop:
        samples.append((pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0,
                        pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)))
        time.sleep(0.02)
def loop(name, fn, secs=2.0):
    global samples, stop
    for _ in range(3): fn()
    torch.cuda.synchronize()
    samples = []; stop = False
    t = threading.Thread(target=sampler); t.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 0; t0 = time.time(); e0.record()
    while time.time() - t0 < secs:
        for _ in range(50): fn()
        n += 50
        torch.cuda.synchronize()
    e1.record(); torch.cuda.synchronize()
    stop = True; t.join()
    ms = e0.elapsed_time(e1) / n
    clk = sorted(s[0] for s in samples[len(samples)//3:]); pw = sorted(s[1] for s in samples[len(samples)//3:])
    rs = set(s[2] for s in samples)
    print(f"{name}: {ms*1e3:.1f} us/launch  sm clock median {clk[len(clk)//2]} MHz (min {clk[0]})  power median {pw[len(pw)//2]:.0f} W max {pw[-1]:.0f} W  throttle reasons {[hex(r) for r in rs]}", flush=True)
    return ms
def mk(M, N, K, mode):
    a = torch.randint(0, 256, (M, K), dtype=torch.uint8, device=cuda); b = torch.randint(0, 256, (N, K), dtype=torch.uint8, device=cuda)
    rowsum = a.to(torch.int32).sum(1).to(torch.int32); sxw = torch.full((N,), 1e-5, device=cuda)
    ow = torch.full((N,), 128, dtype=torch.int32, device=cuda); c0 = torch.zeros(N, dtype=torch.int32, device=cuda)
    G = (N + 127) // 128; so = torch.full((G,), 0.05, device=cuda); oo = torch.full((G,), 128.0, device=cuda)
    kw = dict(so=so, oo=oo, qmax=255.0, qgroup=128)
    if mode == Kn.EPI_ACTMUL: kw.update(lut=torch.randn(256, device=cuda), s2=0.01, o2=128.0)
    if mode == Kn.EPI_RESID: kw.update(resid=torch.zeros(M, N, device=cuda), qmax=65535.0)
    out = Kn.qgemm(a, b, rowsum, sxw, ow, c0, mode, **kw)
    return lambda: Kn.qgemm(a, b, rowsum, sxw, ow, c0, mode, out=None if mode == Kn.EPI_RESID else out, **kw)
M = 8192
for ne in ("8", "16"):
    os.environ["MQ_QGEMM_NE"] = ne; os.environ["MQ_QGEMM_CL"] = "1"
    f = mk(M, 11264, 2048, Kn.EPI_ACTMUL)
    os.environ.pop("MQ_QGEMM_DBG", None)
    ms = loop(f"w1w3 ACTMUL NE={ne}", f); print(f"   {2*M*11264*2048/ms/1e9:.0f} TOP/s")
    os.environ["MQ_QGEMM_DBG"] = "1"
    ms = loop(f"w1w3 ACTMUL NE={ne} epilogue skipped", f); print(f"   {2*M*11264*2048/ms/1e9:.0f} TOP/s")
os.environ.pop("MQ_QGEMM_DBG", None); os.environ["MQ_QGEMM_CL"] = "2"; os.environ["MQ_QGEMM_NE"] = "8"
f = mk(8192, 8192, 8192, Kn.EPI_I32)
ms = loop("8192^3 raw pair", f); print(f"   {2*8192**3/ms/1e9:.0f} TOP/s")
a = torch.randint(-128, 127, (8192, 8192), dtype=torch.int8, device=cuda); b = torch.randint(-128, 127, (8192, 8192), dtype=torch.int8, device=cuda).t()
ms = loop("torch._int_mm 8192^3", lambda: torch._int_mm(a, b)); print(f"   {2*8192**3/ms/1e9:.0f} TOP/s")
x = torch.randn(8192, 8192, device=cuda, dtype=torch.bfloat16); y = torch.randn(8192, 8192, device=cuda, dtype=torch.bfloat16)
ms = loop("torch bf16 matmul 8192^3", lambda: torch.matmul(x, y)); print(f"   {2*8192**3/ms/1e9:.0f} TFLOP/s")
