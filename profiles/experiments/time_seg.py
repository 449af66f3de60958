"""Times the detect-mode seg kernel alone (cfg 2: bs 64, 640 px, C 2, fp32 mask; rotating inputs and outputs so that the
working set between reuses exceeds the L2) for the library named by DSN_LIB (a DSN_VARIANT build) or the default one,
and checks the mask against the C oracle on two images.  python profiles/experiments/time_seg.py [bs] [S]"""
import hashlib
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "..")))
from desenet_b200 import ops  # noqa: E402
from oracle import cport  # noqa: E402

bs = int(sys.argv[1]) if len(sys.argv) > 1 else 64
S = int(sys.argv[2]) if len(sys.argv) > 2 else 640
dev = torch.device("cuda", 0)
g = torch.Generator().manual_seed(0)
lows = [torch.randn(bs, 2, S // 8, S // 8, generator=g).to(dev) for _ in range(4)]
masks = [torch.empty(bs, S, S, device=dev) for _ in range(3)]


def run(i):
    ops.seg_fused(lows[i % 4], (S, S), mask_out=masks[i % 3])


for i in range(10):
    run(i)
torch.cuda.synchronize()
best = 1e9
for rep in range(3):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(200):
        run(i)
    b.record()
    torch.cuda.synchronize()
    best = min(best, a.elapsed_time(b) / 200 * 1e3)
run(0)
torch.cuda.synchronize()
m = masks[0].cpu().numpy()
full = cport.upsample_bilinear_ac(lows[0][:2].cpu().numpy(), S, S)
ok = np.array_equal(m[:2], cport.argmax_first(full).astype(np.float32))
nbytes = bs * S * S * 4 + bs * 2 * (S // 8) ** 2 * 4
print(f"{os.path.basename(os.environ.get('DSN_LIB', 'default')):34s} seg detect bs{bs} S{S}: {best:7.2f} us  {nbytes / best * 1e-3:7.1f} GB/s  "
      f"oracle_ok={ok} sha={hashlib.sha256(m.tobytes()).hexdigest()[:12]}")
