"""Generates tests/golden/scale_coords.npz by running the UNMODIFIED reference's scale_coords / clip_coords
(core/utils/general.py:598-623, imported in place through oracle/ref_shim.py).  Build container only:
    python tests/golden/make_golden_coords.py
Cases mirror the callers: detect.py:218 (no ratio_pad, .round()), val.py:251 (ratio_pad from the dataloader),
common.py:778 (no ratio_pad), boxes partly outside the image so that clip_coords matters."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))
from oracle import ref_shim  # noqa: E402

CASES = [
    dict(img1=(640, 640), img0=(1080, 1920), ratio_pad=None, round=True),
    dict(img1=(384, 640), img0=(720, 1280), ratio_pad=None, round=False),
    dict(img1=(640, 640), img0=(333, 517), ratio_pad=((640 / 517, 640 / 517), (0.0, 114.3)), round=False),
    dict(img1=(1280, 1280), img0=(2160, 3840), ratio_pad=None, round=True),
    dict(img1=(640, 640), img0=(640, 640), ratio_pad=None, round=False),
]


def main():
    ref = ref_shim.load()
    g = torch.Generator().manual_seed(5)
    arrs = {"n_cases": np.int64(len(CASES))}
    for k, c in enumerate(CASES):
        n = 257
        xy = torch.rand(n, 2, generator=g) * torch.tensor([c["img1"][1], c["img1"][0]]) * 1.2 - 30
        wh = torch.rand(n, 2, generator=g) * 300
        boxes = torch.cat((xy - wh / 2, xy + wh / 2), 1).float()
        out = ref.general.scale_coords(c["img1"], boxes.clone(), c["img0"], c["ratio_pad"])
        if c["round"]:
            out = out.round()
        arrs[f"in{k}"] = boxes.numpy()
        arrs[f"out{k}"] = out.numpy()
    np.savez_compressed(os.path.join(HERE, "scale_coords.npz"), **arrs)
    print("scale_coords.npz written")


if __name__ == "__main__":
    main()
