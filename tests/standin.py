"""Stand-in for the reference model that honours the attribute contract ``desenet_b200.patch.install`` relies on
(the reference itself cannot travel to the GPU box):

  * a class NAMED ``Detect`` with ``m`` / ``anchor_grid`` / ``stride`` / ``nl`` / ``na`` / ``no`` whose ``forward`` is a
    plain-torch restatement of core/models/yolo.py:255-277 (mutates the input list in place like the original);
  * a class named ``SegMaskPSP`` whose ``out`` is an ``nn.Sequential`` ending in
    ``nn.Upsample(scale_factor=8, mode='bilinear', align_corners=True)`` (yolo.py:178-184);
  * a ``Model`` whose ``forward(img, augment=False, visualize=False)`` returns ``((z, [x_i]), seg_logits)`` in eval
    (yolo.py:344-356).

The "network body" is a lookup: the model holds the raw head tensors of a golden fixture (``real_desenet_s128.npz``:
outputs of the real random-init DeSeNet-s captured at ``Detect.m[i]`` and ``SegMaskPSP.out[3]``) as buffers and the
heads' last convs are ``nn.Identity``, so on any device the raw head outputs are bit-identical to the fixture's and the
fixture's reference results (``nms_val_*``, ``mask*``, ``pix_*`` ...) are the expected outputs of the call sequences of
scripts/val.py:219-234 and scripts/detect.py:158-198.

Also builds a fake ``scripts.val`` / ``scripts.detect`` module pair that imports the hot-path callables BY NAME from a
fake ``core.utils.*`` (CPU implementations = the oracle port), which is what ``patch.install`` has to patch."""
import sys
import types

import torch
import torch.nn as nn


class Detect(nn.Module):
    def __init__(self, nc, anchor_grid, stride):
        super().__init__()
        ag = torch.as_tensor(anchor_grid, dtype=torch.float32).reshape(len(stride), -1, 2)
        self.nc, self.no, self.nl, self.na = nc, nc + 5, ag.shape[0], ag.shape[1]
        self.grid = [torch.zeros(1)] * self.nl
        self.register_buffer("anchors", ag.clone() / torch.as_tensor(stride, dtype=torch.float32).view(-1, 1, 1))
        self.register_buffer("anchor_grid", ag.clone().view(self.nl, 1, -1, 1, 1, 2))
        self.stride = torch.as_tensor(stride, dtype=torch.float32)      # plain attribute as in yolo.py:315 (stays on the CPU)
        self.m = nn.ModuleList(nn.Identity() for _ in range(self.nl))
        self.inplace = True

    def forward(self, x):
        z = []
        for i in range(self.nl):
            x[i] = self.m[i](x[i])
            bs, _, ny, nx = x[i].shape
            x[i] = x[i].view(bs, self.na, self.no, ny, nx).permute(0, 1, 3, 4, 2).contiguous()
            if not self.training:
                yv, xv = torch.meshgrid([torch.arange(ny), torch.arange(nx)], indexing="ij")
                grid = torch.stack((xv, yv), 2).view((1, 1, ny, nx, 2)).to(x[i])
                y = x[i].sigmoid()
                y[..., 0:2] = (y[..., 0:2] * 2. - 0.5 + grid) * self.stride[i]
                y[..., 2:4] = (y[..., 2:4] * 2) ** 2 * self.anchor_grid[i]
                z.append(y.view(bs, -1, self.no))
        return x if self.training else (torch.cat(z, 1), x)


class SegMaskPSP(nn.Module):
    def __init__(self, n_segcls):
        super().__init__()
        self.c_out = n_segcls
        self.out = nn.Sequential(nn.Identity(), nn.Upsample(scale_factor=8, mode="bilinear", align_corners=True))

    def forward(self, x):
        return self.out(x)


class Model(nn.Module):
    def __init__(self, fixture):
        super().__init__()
        nc = int(fixture["raw0"].shape[1]) // 3 - 5
        for i in range(3):
            self.register_buffer(f"raw{i}", torch.from_numpy(fixture[f"raw{i}"]).clone())
        self.register_buffer("seg_low", torch.from_numpy(fixture["seg_low"]).clone())
        self.dummy = nn.Parameter(torch.zeros(1))            # next(model.parameters()) gives the device (val.py:52)
        self.seg = SegMaskPSP(int(fixture["seg_low"].shape[1]))
        self.det = Detect(nc, fixture["anchor_grid"], fixture["stride"].tolist())
        self.stride = self.det.stride
        self.de_names = [str(i) for i in range(nc)]
        self.se_names = [str(i) for i in range(self.seg.c_out)]

    def forward(self, img, augment=False, visualize=False):
        bs = img.shape[0]
        feats = [getattr(self, f"raw{i}")[:bs].clone().to(img.dtype) for i in range(3)]
        seg = self.seg(self.seg_low[:bs].to(img.dtype))
        return self.det(feats), seg


def fake_reference_modules(prefix="dsn_fake_ref"):
    """Modules laid out like the reference: ``<prefix>.core.utils.{general,plots,metrics}`` define the CPU callables
    (oracle port), ``<prefix>.scripts.{val,detect}`` import them by name.  Returns (namespace, list of module names)."""
    from oracle import port

    def mod(name):
        m = types.ModuleType(name)
        sys.modules[name] = m
        return m

    general, plots, metrics = mod(f"{prefix}.core.utils.general"), mod(f"{prefix}.core.utils.plots"), mod(f"{prefix}.core.utils.metrics")

    def non_max_suppression(prediction, conf_thres=0.25, iou_thres=0.45, classes=None, agnostic=False, multi_label=False,
                            labels=(), max_det=300):
        return port.non_max_suppression(prediction.float(), conf_thres, iou_thres, classes, agnostic, multi_label, labels, max_det)

    def scale_coords(img1_shape, coords, img0_shape, ratio_pad=None):      # in place, returns coords (general.py:598-611)
        coords[:, :4] = port.scale_coords(img1_shape, coords, img0_shape, ratio_pad)[:, :4]
        return coords

    def segoutput_to_target(output, size=None):
        return port.segoutput_to_target(output, size)

    general.non_max_suppression = non_max_suppression
    general.scale_coords = scale_coords
    plots.segoutput_to_target = segoutput_to_target
    metrics.batch_pix_accuracy = lambda output, target: port.batch_pix_accuracy(output, target)
    metrics.batch_intersection_union = lambda output, target, nclass: port.batch_intersection_union(output, target, nclass)
    val, detect = mod(f"{prefix}.scripts.val"), mod(f"{prefix}.scripts.detect")
    for m in (val, detect):                       # `from core.utils.general import ...` (val.py:30-38, detect.py:27-32)
        m.non_max_suppression = general.non_max_suppression
        m.scale_coords = general.scale_coords
        m.segoutput_to_target = plots.segoutput_to_target
    val.batch_pix_accuracy = metrics.batch_pix_accuracy
    val.batch_intersection_union = metrics.batch_intersection_union
    names = [general.__name__, plots.__name__, metrics.__name__, val.__name__, detect.__name__]
    return types.SimpleNamespace(general=general, plots=plots, metrics=metrics, val=val, detect=detect), names


def val_sequence(ns, model, img, seg_targets, conf_thres=0.001, iou_thres=0.6, single_cls=False, half=False):
    """scripts/val.py:211-234, verbatim up to names: returns (out, seg_out)."""
    device = next(model.parameters()).device
    img = img.to(device, non_blocking=True)
    img = img.half() if half else img.float()                                   # val.py:213
    img /= 255.0
    seg_targets = seg_targets.to(device)
    (out, train_out), seg_out = model(img, augment=False)                        # val.py:220
    lb = []                                                                      # save_hybrid=False (val.py:228)
    out = ns.val.non_max_suppression(out, conf_thres, iou_thres, labels=lb, multi_label=True, agnostic=single_cls)   # val.py:230
    seg_out = ns.val.segoutput_to_target(seg_out, seg_targets[0].shape)          # val.py:234
    return out, seg_out, train_out


def detect_sequence(ns, model, img, im0_shape, conf_thres=0.25, iou_thres=0.45, classes=None, agnostic_nms=False,
                    max_det=1000, half=False):
    """scripts/detect.py:144-198 + :218, verbatim up to names: returns (pred, seg_pred)."""
    device = next(model.parameters()).device
    img = img.to(device)
    img = img.half() if half else img.float()                                   # detect.py:146
    img = img / 255.0
    (pred, _), seg_pred = model(img, augment=False, visualize=False)             # detect.py:158
    pred = ns.detect.non_max_suppression(pred, conf_thres, iou_thres, classes, agnostic_nms, max_det=max_det)   # detect.py:190
    seg_pred = ns.detect.segoutput_to_target(seg_pred.cpu(), size=im0_shape[:2])                  # detect.py:198
    for det in pred:
        if len(det):
            det[:, :4] = ns.detect.scale_coords(img.shape[2:], det[:, :4], im0_shape).round()    # detect.py:218
    return pred, seg_pred
