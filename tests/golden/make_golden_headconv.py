"""Generates tests/golden/real_desenet_s128_headconv.npz from the UNMODIFIED reference (imported in place through
oracle/ref_shim.py): the input, weight and bias of the seg head's last convolution (core/models/yolo.py:182,
`SegMaskPSP.out[3] = nn.Conv2d(c_hid, n_segcls, 1)`) of the same random-init DeSeNet-s and input image as
real_desenet_s128.npz, whose `seg_low` is that convolution's output.  Build container only:
    python tests/golden/make_golden_headconv.py"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))
from oracle import ref_shim  # noqa: E402


def main():
    ref_shim.load()
    torch.manual_seed(0)
    model = ref_shim.build_desenet_s(n_segcls=2)
    conv = model.model[-2].out[3]
    cap = {}
    h = conv.register_forward_hook(lambda m, a, o: cap.update(feat=a[0].detach().clone(), out=o.detach().clone()))
    img = torch.rand(2, 3, 128, 128)
    with torch.no_grad():
        model(img)
    h.remove()
    old = np.load(os.path.join(HERE, "real_desenet_s128.npz"))
    assert np.array_equal(old["seg_low"], cap["out"].numpy()), "not the model / image of real_desenet_s128.npz"
    path = os.path.join(HERE, "real_desenet_s128_headconv.npz")
    np.savez_compressed(path, feat=cap["feat"].numpy(), weight=conv.weight.detach().numpy(), bias=conv.bias.detach().numpy(),
                        seg_low=cap["out"].numpy())
    print(path, os.path.getsize(path) // 1024, "KiB", tuple(cap["feat"].shape), tuple(conv.weight.shape))


if __name__ == "__main__":
    main()
