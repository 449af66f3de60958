"""Generates tests/golden/*.npz by running the UNMODIFIED reference (splwany/DeSeNet, imported in
place from /root/reference through oracle/ref_shim.py) on seeded inputs.  Run in the build
container only:  python tests/golden/make_golden.py

Every output array in these files was produced by the reference's own functions:
  z / xperm      core/models/yolo.py:255-277    Detect.forward (Identity convs or the real model)
  nms*           core/utils/general.py:659-750  non_max_suppression (one image per call, so the
                                                10 s wall-clock break never truncates)
  seg_full       core/models/yolo.py:183        nn.Upsample(8, bilinear, align_corners=True)
  mask*          core/utils/plots.py:222-229    segoutput_to_target
  pix / iu       core/utils/metrics.py:350-388  batch_pix_accuracy / batch_intersection_union
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))
from oracle import port, ref_shim  # noqa: E402

NMS_PARAM_SETS = [
    dict(conf_thres=0.25, iou_thres=0.45),
    dict(conf_thres=0.001, iou_thres=0.6, multi_label=True, max_det=300),
    dict(conf_thres=0.1, iou_thres=0.5, classes=[0, 2]),
    dict(conf_thres=0.05, iou_thres=0.3, agnostic=True, max_det=7),
    dict(conf_thres=0.25, iou_thres=0.45, multi_label=True, max_det=1000),
]


def ref_nms_per_image(ref, z, **kw):
    return [ref.non_max_suppression(z[i:i + 1].clone(), **kw)[0].numpy() for i in range(z.shape[0])]


def identity_detect(ref, nc):
    det = ref.Detect(nc, port.DEFAULT_ANCHORS, (8, 8, 8)).eval()
    det.stride = torch.tensor(port.DEFAULT_STRIDES)
    det.m = torch.nn.ModuleList([torch.nn.Identity()] * 3)
    return det


def save(name, **arrs):
    path = os.path.join(HERE, name)
    np.savez_compressed(path, **arrs)
    print(f"{name}: {os.path.getsize(path) / 1024:.0f} KiB, {len(arrs)} arrays")


def synth_case(ref, name, bs, S, nc, seed, param_sets, **synth_kw):
    levels, seg, tgt = port.synth_head_outputs(bs, S, nc, C=2, seed=seed, **synth_kw)
    det = identity_detect(ref, nc)
    with torch.no_grad():
        z, xs = det([l.clone() for l in levels])
    arrs = {f"raw{i}": l.numpy() for i, l in enumerate(levels)}
    arrs["z"] = z.numpy()
    arrs["nc"] = np.int64(nc)
    arrs["S"] = np.int64(S)
    for k, kw in enumerate(param_sets):
        for i, o in enumerate(ref_nms_per_image(ref, z, **kw)):
            arrs[f"nms{k}_{i}"] = o
    return arrs


def main():
    ref = ref_shim.load()
    up8 = torch.nn.Upsample(scale_factor=8, mode="bilinear", align_corners=True)

    # ---- P family: real head outputs of a random-init DeSeNet-s ----
    torch.manual_seed(0)
    model = ref_shim.build_desenet_s(n_segcls=2)
    det_mod = model.model[-1]
    seg_mod = model.model[-2]
    cap = {}
    hooks = [det_mod.m[i].register_forward_hook(lambda m, a, o, i=i: cap.__setitem__(f"raw{i}", o.detach().clone()))
             for i in range(3)]
    hooks.append(seg_mod.out[3].register_forward_hook(lambda m, a, o: cap.__setitem__("seg_low", o.detach().clone())))
    img = torch.rand(2, 3, 128, 128)
    with torch.no_grad():
        (z, xs), seg_full = model(img)
    for h in hooks:
        h.remove()
    g = torch.Generator().manual_seed(7)
    target = torch.randint(0, 2, (2, 128, 128), generator=g, dtype=torch.int64)
    arrs = {k: v.numpy() for k, v in cap.items()}
    arrs.update(z=z.numpy(), seg_full=seg_full.numpy(), target=target.numpy(),
                anchor_grid=det_mod.anchor_grid.view(3, 3, 2).numpy(), stride=det_mod.stride.numpy())
    for i, x in enumerate(xs):
        arrs[f"xperm{i}"] = x.numpy()
    for i, o in enumerate(ref_nms_per_image(ref, z, conf_thres=0.001, iou_thres=0.6, multi_label=True, max_det=300)):
        arrs[f"nms_val_{i}"] = o
    for i, o in enumerate(ref_nms_per_image(ref, z, conf_thres=0.25, iou_thres=0.45)):
        arrs[f"nms_det_{i}"] = o
    arrs["mask"] = ref.segoutput_to_target(seg_full, (128, 128)).numpy()
    arrs["mask_96x170"] = ref.segoutput_to_target(seg_full, (96, 170)).numpy()
    for nclass in (2, 3):
        c, l = ref.batch_pix_accuracy(seg_full, target)
        i_, u_ = ref.batch_intersection_union(seg_full, target, nclass)
        arrs[f"pix_{nclass}"] = np.array([c, l], np.int64)
        arrs[f"inter_{nclass}"] = i_.astype(np.int64)
        arrs[f"union_{nclass}"] = u_.astype(np.int64)
    save("real_desenet_s128.npz", **arrs)

    # ---- T family, nc=6 and nc=80 ----
    save("synthT_s160_nc6.npz", **synth_case(ref, "t6", 2, 160, 6, 1, NMS_PARAM_SETS))
    save("synthT_s64_nc80.npz", **synth_case(ref, "t80", 2, 64, 80, 2, NMS_PARAM_SETS))

    # ---- > 30000 candidates (top-30000 path, general.py:728-729); no score ties allowed near the cut ----
    levels, _, _ = port.synth_head_outputs(1, 160, 80, seed=3, obj_mean=0.0, cls_mean=0.0)
    det = identity_detect(ref, 80)
    with torch.no_grad():
        z, _ = det([l.clone() for l in levels])
    x = z[0][z[0][:, 4] > 0.001]
    conf = (x[:, 5:] * x[:, 4:5])
    sc = conf[conf > 0.001]
    assert sc.numel() > 30000, sc.numel()
    srt = torch.sort(sc, descending=True).values
    assert srt[29999] != srt[30000], "tie at the top-30000 cut: pick another seed"
    out = ref.non_max_suppression(z.clone(), conf_thres=0.001, iou_thres=0.6, multi_label=True, max_det=300)[0]
    save("over30k_s160_nc80.npz", raw0=levels[0].numpy(), raw1=levels[1].numpy(), raw2=levels[2].numpy(),
         nms0_0=out.numpy(), n_pairs=np.int64(sc.numel()), nc=np.int64(80), S=np.int64(160))

    # ---- seg-only cases: odd sizes, C in {2,3,4}, nclass in {2,3,4}, out-of-range targets ----
    arrs = {}
    g = torch.Generator().manual_seed(11)
    k = 0
    for C, h, w in [(2, 20, 20), (3, 9, 13), (4, 16, 24), (2, 7, 5)]:
        low = torch.randn(2, C, h, w, generator=g)
        full = up8(low)
        tgt = torch.randint(-1, C + 2, (2, 8 * h, 8 * w), generator=g, dtype=torch.int64)
        tgt[0, 0, :5] = 255
        arrs[f"low{k}"] = low.numpy()
        arrs[f"full{k}"] = full.numpy()
        arrs[f"target{k}"] = tgt.numpy()
        arrs[f"mask{k}"] = ref.segoutput_to_target(full).numpy()
        arrs[f"mask{k}_resized"] = ref.segoutput_to_target(full, (3 * h + 1, 11 * w - 2)).numpy()
        arrs[f"resized_size{k}"] = np.array([3 * h + 1, 11 * w - 2], np.int64)
        for nclass in (2, 3, 4):
            c, l = ref.batch_pix_accuracy(full, tgt)
            i_, u_ = ref.batch_intersection_union(full, tgt, nclass)
            arrs[f"pix{k}"] = np.array([c, l], np.int64)
            arrs[f"inter{k}_{nclass}"] = i_.astype(np.int64)
            arrs[f"union{k}_{nclass}"] = u_.astype(np.int64)
        k += 1
    arrs["n_cases"] = np.int64(k)
    save("seg_cases.npz", **arrs)


if __name__ == "__main__":
    main()
