"""scale_coords / clip_coords: the oracle port reproduces the reference-generated
fixture (CPU), and the CUDA kernel reproduces it bit-for-bit through the C ABI (GPU)."""
import numpy as np
import pytest
import torch

from oracle import port
from tests.golden.make_golden_coords import CASES


def test_oracle_scale_coords_matches_reference_fixture(golden):
    g = golden("scale_coords.npz")
    assert int(g["n_cases"]) == len(CASES)
    for k, c in enumerate(CASES):
        out = port.scale_coords(c["img1"], torch.from_numpy(g[f"in{k}"]), c["img0"], c["ratio_pad"], c["round"])
        assert np.array_equal(out.numpy(), g[f"out{k}"]), k


@pytest.mark.gpu
def test_cuda_scale_coords_bit_exact(golden):
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import desenet_b200 as dsn
    from desenet_b200 import ops
    g = golden("scale_coords.npz")
    for k, c in enumerate(CASES):
        # 1. the reference signature, in place, on a row-strided view of a (n, 6) detection tensor
        det = torch.zeros(g[f"in{k}"].shape[0], 6)
        det[:, :4] = torch.from_numpy(g[f"in{k}"])
        det[:, 4] = 0.5
        det = det.cuda()
        ret = dsn.scale_coords(c["img1"], det[:, :4], c["img0"], c["ratio_pad"])
        assert ret.data_ptr() == det.data_ptr()
        got = det[:, :4].round() if c["round"] else det[:, :4]
        assert np.array_equal(got.cpu().numpy(), g[f"out{k}"]), k
        assert (det[:, 4] == 0.5).all()
        # 2. batched, fused rounding, per-image counts: rows beyond the count stay untouched
        buf = torch.zeros(2, g[f"in{k}"].shape[0], 6)
        buf[:, :, :4] = torch.from_numpy(g[f"in{k}"])
        buf = buf.cuda()
        params = torch.tensor([ops.letterbox_params(c["img1"], c["img0"], c["ratio_pad"])] * 2, dtype=torch.float32).cuda()
        counts = torch.tensor([buf.shape[1], 100], dtype=torch.int32).cuda()
        ops.scale_coords_(buf, params, counts, do_round=c["round"])
        assert np.array_equal(buf[0, :, :4].cpu().numpy(), g[f"out{k}"])
        assert np.array_equal(buf[1, :100, :4].cpu().numpy(), g[f"out{k}"][:100])
        assert np.array_equal(buf[1, 100:, :4].cpu().numpy(), g[f"in{k}"][100:])
    b = torch.tensor([[-5.0, 700.0, 650.0, 3.0]]).cuda()
    dsn.clip_coords(b, (480, 640))
    assert b.cpu().tolist() == [[0.0, 480.0, 640.0, 3.0]]
