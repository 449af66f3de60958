"""Quick on-GPU check + timing of dsn_detect_head_conv1x1 (tcgen05 path) against the float64 convolution.
    timeout 120 python profiles/experiments/check_detconv.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "..")))
from desenet_b200 import ops  # noqa: E402

torch.manual_seed(0)
for dt in (torch.float16, torch.bfloat16):
    for bs, K, ny, nx, N in ((1, 64, 16, 8, 16), (2, 128, 20, 20, 33), (3, 256, 40, 40, 33), (1, 512, 80, 80, 33), (2, 64, 16, 24, 255)):
        x = torch.randn(bs, K, ny, nx).to(dt)
        w = (torch.randn(N, K, 1, 1) / K ** 0.5).to(dt)
        b = torch.randn(N).to(dt)
        out = ops.detect_head_conv1x1(x.cuda(), w.cuda(), b.cuda())
        torch.cuda.synchronize()
        exact = torch.nn.functional.conv2d(x.double(), w.double(), b.double())
        err = (out.cpu().double() - exact).abs()
        ulp = 2.0 ** (-10 if dt == torch.float16 else -7)
        bad = (err > ulp * exact.abs() + 1e-4).sum().item()
        print(f"{str(dt):16s} bs{bs} K{K} {ny}x{nx} N{N}: max err {err.max():.3e} max|ref| {exact.abs().max():.2f} bad {bad}/{err.numel()}", flush=True)
        if bad:
            e = err.reshape(bs, N, -1)
            print("   bad per output channel (first image):", (e[0] > ulp * exact.reshape(bs, N, -1)[0].abs() + 1e-4).sum(1).tolist()[:40])
            print("   bad per pixel block of 8 (first image, ch 0):", (e[0, 0].reshape(-1, 8) > 1e-2).sum(1).tolist()[:64])
# timing at the DeSeNet-s sizes, bs 64
for dt in (torch.float16, torch.float32):
    tot = 0.0
    for K, s in ((128, 80), (256, 40), (512, 20)):
        x = torch.randn(64, K, s, s, device="cuda").to(dt)
        w = (torch.randn(33, K, 1, 1, device="cuda") / K ** 0.5).to(dt)
        b = torch.randn(33, device="cuda").to(dt)
        conv = torch.nn.Conv2d(K, 33, 1).cuda().to(dt)
        for name, fn in (("dsn", lambda: ops.detect_head_conv1x1(x, w, b)), ("torch", lambda: conv(x))):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(50):
                fn()
            e.record()
            torch.cuda.synchronize()
            us = a.elapsed_time(e) / 50 * 1e3
            gb = x.numel() * x.element_size() + 64 * 33 * s * s * x.element_size()
            print(f"{str(dt):14s} K{K} {s}x{s} bs64 {name:6s}: {us:8.1f} us  {gb / us * 1e-3:7.1f} GB/s", flush=True)
