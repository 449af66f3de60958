"""Runs dsn_detect_head_conv1x1 (tcgen05 path) on the three DeSeNet-s levels at bs 64 for an ncu capture."""
import os
import sys

import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "..")))
from desenet_b200 import ops  # noqa: E402

dt = torch.float16
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for K, s in ((128, 80), (256, 40), (512, 20)):
    x = torch.randn(64, K, s, s, device="cuda").to(dt)
    w = (torch.randn(33, K, 1, 1, device="cuda") / K ** 0.5).to(dt)
    b = torch.randn(33, device="cuda").to(dt)
    for _ in range(2):
        flush.zero_()
        ops.detect_head_conv1x1(x, w, b)
torch.cuda.synchronize()
print("ok")
