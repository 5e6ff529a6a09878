"""Developer tool (GPU box): time train_generator at a given geometry. Usage: time_generator.py [N H W iters]"""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import srgan_b200 as S

a = [int(x) for x in sys.argv[1:]]
N, H, W = a[:3] if len(a) >= 3 else (16, 96, 96)
iters = a[3] if len(a) > 3 else 10
torch.manual_seed(0)
g = S.SRResNet().cuda()
crit = S.ReconstructionLoss()
opt = S.Adam(g.parameters(), lr=1e-4)
lr = torch.rand(N, 3, H, W, device="cuda")
hr = torch.rand(N, 3, 4 * H, 4 * W, device="cuda")
for _ in range(3):
    vals = S.train_generator(g, None, lr, hr, None, crit, opt)
print("losses", vals)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.time()
e0.record()
for _ in range(iters):
    out = S.train_generator_async(g, None, lr, hr, None, crit, opt)
e1.record()
t_host = time.time() - t0
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / iters
flop = 13277952.0 * H * W * N
print(f"train_generator {N}x3x{H}x{W}: {ms:.3f} ms/step (host enqueue {1e3*t_host/iters:.3f} ms)  {N/ms*1e3:.1f} patches/s  "
      f"{flop/ms/1e9:.1f} TFLOP/s algorithmic = {flop/ms/1e9/1657.8*100:.1f}% of bf16 peak")
# phase split
def timed(fn, n=5):
    torch.cuda.synchronize(); e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / n
g.train()
def fwd():
    with torch.no_grad(): g(lr)
print("forward only (train-mode BN): %.3f ms" % timed(fwd))
sr = g(lr)
c, t = crit(hr, sr)
print("loss fwd: %.3f ms" % timed(lambda: crit(hr, sr.detach())))
(c + t).backward()
g.eval()
print("forward eval: %.3f ms" % timed(fwd))
print("launches", g.launch_count())
