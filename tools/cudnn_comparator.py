"""Developer tool (GPU box): the SECOND comparator SURVEY 8d recommends -- the same 3-generator pixel-loss train step
written with stock torch.nn modules and run through PyTorch's own CUDA path (cuDNN convolutions, bf16 autocast,
channels_last, fused Adam), eager and CUDA-graph captured.  It is NOT part of the product path or of bench.py's
contract; it exists to put the hand-written kernels next to what torch + cuDNN achieve on the same B200.
Usage: python tools/cudnn_comparator.py [steps]"""
import sys
import time

import torch
import torch.nn as nn
import torch.nn.functional as F


class ResidualBlock(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.conv1 = nn.Conv2d(c, c, 3, padding=1); self.bn1 = nn.BatchNorm2d(c)
        self.conv2 = nn.Conv2d(c, c, 3, padding=1); self.bn2 = nn.BatchNorm2d(c)

    def forward(self, x):
        return self.bn2(self.conv2(F.relu(self.bn1(self.conv1(x))))) + x


class Net(nn.Module):
    def __init__(self):
        super().__init__()
        self.conv1 = nn.Conv2d(3, 64, 9, padding=4)
        self.blocks = nn.Sequential(*[ResidualBlock(64) for _ in range(16)])
        self.conv2 = nn.Conv2d(64, 64, 3, padding=1)
        self.up = nn.Sequential(nn.Conv2d(64, 256, 3, padding=1), nn.PixelShuffle(2), nn.ReLU(),
                                nn.Conv2d(64, 256, 3, padding=1), nn.PixelShuffle(2), nn.ReLU())
        self.conv3 = nn.Conv2d(64, 3, 9, padding=4)

    def forward(self, x):
        o1 = F.leaky_relu(self.conv1(x), 0.2)
        return self.conv3(self.up(self.conv2(self.blocks(o1)) + o1))


def dw(x, k):
    return F.conv2d(x, k.expand(3, 1, 3, 3), padding=1, groups=3)


def recon_loss(hr, sr, px, py, lap):
    e0 = torch.maximum(dw(hr, px).abs(), dw(hr, py).abs())
    e = ((e0 - e0.mean()) / e0.std() * 0.2 + 1).clamp(0, 2)
    edge = ((hr - sr).abs() * e).sum() / e.sum()
    tv = F.relu((dw(sr, lap).abs() * (1 - e)).mean())
    return edge + tv


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
    torch.backends.cudnn.benchmark = True
    dev = "cuda"
    K, B, H = 3, 16, 96
    gens = [Net().to(dev).to(memory_format=torch.channels_last) for _ in range(K)]
    opts = [torch.optim.Adam(g.parameters(), lr=1e-4, fused=True, capturable=True) for g in gens]
    lr = torch.rand(B, 3, H, H, device=dev).contiguous(memory_format=torch.channels_last)
    hr = torch.rand(B, 3, 4 * H, 4 * H, device=dev)
    px = torch.tensor([[-5., 0, 5]] * 3, device=dev).view(1, 1, 3, 3)
    py = px.transpose(2, 3).contiguous()
    lap = torch.full((1, 1, 3, 3), -0.125, device=dev); lap[0, 0, 1, 1] = 1

    def step():
        for g, o in zip(gens, opts):
            with torch.autocast("cuda", dtype=torch.bfloat16):
                sr = g(lr)
            loss = recon_loss(hr, sr.float(), px, py, lap)
            o.zero_grad(set_to_none=True)
            loss.backward()
            o.step()

    def timed(fn, n):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    for _ in range(5):
        step()
    ms = timed(step, steps)
    print(f"torch+cuDNN bf16 autocast channels_last, eager : {ms:.2f} ms/step  {B / ms * 1e3:.0f} patches/s")
    try:
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):
                step()
        torch.cuda.current_stream().wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        for o in opts:
            o.zero_grad(set_to_none=True)
        with torch.cuda.graph(graph):
            step()
        ms = timed(graph.replay, steps)
        print(f"torch+cuDNN bf16 autocast channels_last, graphed: {ms:.2f} ms/step  {B / ms * 1e3:.0f} patches/s")
    except Exception as ex:  # pragma: no cover
        print("graph capture of the torch path failed:", repr(ex)[:200])
    print("torch", torch.__version__, "cudnn", torch.backends.cudnn.version())


if __name__ == "__main__":
    main()
