import os, sys, ctypes, torch
os.environ['DSN_LIB'] = os.path.abspath(os.path.join(os.path.dirname(__file__), '..', '..', 'desenet_b200', 'lib', 'libdesenet_b200_timing.so'))   # DSN_TIMING=1 python -m desenet_b200.build
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), '..', '..')))
from bench import CONFIGS
from desenet_b200 import ops, _lib
from desenet_b200.synth import DEFAULT_ANCHORS, DEFAULT_STRIDES, synth_head_outputs
cfg=dict(CONFIGS[int(sys.argv[1]) if len(sys.argv)>1 else 3]); dev=torch.device('cuda',0)
bs=int(sys.argv[2]) if len(sys.argv)>2 else 64
levels,seg,tgt=synth_head_outputs(bs,cfg['S'],cfg['nc'],cfg['C'],seed=0,with_target=False)
head=ops.Head([l.to(dev) for l in levels],DEFAULT_ANCHORS,DEFAULT_STRIDES)
lib=_lib.load()
lib.dsn_debug_timing.argtypes=[ctypes.c_void_p]
lib.dsn_debug_log.argtypes=[ctypes.c_void_p]
buf=(ctypes.c_ulonglong*64)()
logbuf=(ctypes.c_ulonglong*2048)()
use_cut=os.environ.get('NOCUT') is None
for it in range(5):
    cut=ops.conf_cut_levels(head,cfg['conf'],cfg['multi_label'],target=ops.cut_target_for(cfg['max_det'])) if use_cut else None
    keys,cnt=ops.filter_levels(head,cfg['conf'],cfg['multi_label'],cut=cut)
    torch.cuda.synchronize()
    a=torch.cuda.Event(enable_timing=True); b=torch.cuda.Event(enable_timing=True)
    a.record()
    det,dc=ops.nms_levels(head,keys,cnt,cfg['conf'],cfg['iou'],False,cfg['max_det'],cut=cut)
    b.record(); torch.cuda.synchronize()
    lib.dsn_debug_timing(buf)
    lib.dsn_debug_log(logbuf)
    t=list(buf)
    ent=[(v>>56, v & ((1<<56)-1)) for v in logbuf if v]
    sums={}
    for (s0,t0),(s1,t1) in zip(ent[:-1], ent[1:]):
        sums[s1]=sums.get(s1,0)+(t1-t0)
    for k in range(64):
        if k>=32 and k<50: t[k]=sums.get(k,0)*(it+1)      # same per-iteration print below
    total_log=(ent[-1][1]-ent[0][1])/1e3 if ent else 0
    if it==4: print('sequence (slot:us):', ' '.join(f"{s1}:{(t1-t0)/1e3:.1f}" for (s0,t0),(s1,t1) in zip(ent[:-1], ent[1:])))
    laps={32:'select',33:'gather',34:'sort',35:'boxes',36:'extent',37:'rsetup',45:'rprev',46:'a_chain',47:'alive',38:'a_kept',39:'b_pairs',40:'c_scan',41:'d_append'}
    cnts={50:'rounds',51:'consumed',52:'chunks',53:'gathered',54:'consumed_sum_all',55:'consumed_max',56:'gathered_sum_all',57:'gathered_max'}
    print('counters/iter:', ' '.join(f"{v}={t[k]/(it+1):.0f}" for k,v in cnts.items()))
    print('iter',it,'nms total us %.1f'%(a.elapsed_time(b)*1e3),'block0 timeline us %.1f (%d laps)'%(total_log,len(ent)),'keys',cnt[0].item(),'det',dc[0].item(),'| laps(us cumulative/iter):', ' '.join(f"{v}={t[k]/1e3/(it+1):.1f}" for k,v in laps.items()))
