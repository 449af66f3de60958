import sys, torch, time
sys.path.insert(0,'/root/repo')
from bench import CONFIGS
from desenet_b200 import ops
from desenet_b200.pipeline import PostProcessor
from desenet_b200.synth import DEFAULT_ANCHORS, DEFAULT_STRIDES, synth_head_outputs
cfg=dict(CONFIGS[int(sys.argv[1])]); dev=torch.device('cuda',0)
ev=cfg['mode']=='eval'
levels,seg,tgt=synth_head_outputs(cfg['bs'],cfg['S'],cfg['nc'],cfg['C'],seed=0,with_target=ev)
levels=[l.to(dev) for l in levels]; seg=seg.to(dev); tgt=tgt.to(dev) if ev else None
for overlap in (False, True):
    pp=PostProcessor(cfg['bs'],cfg['S'],cfg['nc'],cfg['C'],DEFAULT_ANCHORS,DEFAULT_STRIDES,conf_thres=cfg['conf'],iou_thres=cfg['iou'],multi_label=cfg['multi_label'],max_det=cfg['max_det'],mode=cfg['mode'],nclass=2,device=dev,overlap_streams=overlap)
    head=ops.Head(levels,DEFAULT_ANCHORS,DEFAULT_STRIDES)
    for i in range(3): pp.run_head(head,seg,tgt)
    torch.cuda.synchronize()
    ts=[]
    for i in range(10):
        a=torch.cuda.Event(enable_timing=True); b=torch.cuda.Event(enable_timing=True)
        a.record(); pp.run_head(head,seg,tgt); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b)*1e3)
    print('overlap',overlap,'direct us',[round(t) for t in ts])
    pp.capture(levels,seg,tgt)
    ts=[]
    for i in range(10):
        a=torch.cuda.Event(enable_timing=True); b=torch.cuda.Event(enable_timing=True)
        a.record(); pp.replay(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b)*1e3)
    print('overlap',overlap,'graph us',[round(t) for t in ts])
    print(pp.counts[:5].tolist(), pp.det_count[:5].tolist())
