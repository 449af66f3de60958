"""NMS parameter sets shared by the golden generator and the tests (index = key in the fixtures)."""
NMS_PARAM_SETS = [
    dict(conf_thres=0.25, iou_thres=0.45),
    dict(conf_thres=0.001, iou_thres=0.6, multi_label=True, max_det=300),
    dict(conf_thres=0.1, iou_thres=0.5, classes=[0, 2]),
    dict(conf_thres=0.05, iou_thres=0.3, agnostic=True, max_det=7),
    dict(conf_thres=0.25, iou_thres=0.45, multi_label=True, max_det=1000),
]
