O=gpurun_out; mkdir -p $O
B="python bench.py --config 3 --steps 50 --warmup 10 --no-cpu-baseline --no-e2e --no-secondary"
DSN_VARIANT="n256:-DDSN_NMS_THREADS=256:nms.cu,fused.cu" $B > $O/x7_cfg3_n256.json 2>$O/x7_n256.err
DSN_VARIANT="n256e3:-DDSN_NMS_THREADS=256 -DDSN_SEG_EVAL_CTAS_PER_SM=3:nms.cu,fused.cu,seg.cu" $B > $O/x7_cfg3_n256e3.json 2>$O/x7_n256e3.err
DSN_VARIANT="n256e2:-DDSN_NMS_THREADS=256 -DDSN_SEG_EVAL_CTAS_PER_SM=2:nms.cu,fused.cu,seg.cu" $B > $O/x7_cfg3_n256e2.json 2>$O/x7_n256e2.err
DSN_VARIANT="n256:-DDSN_NMS_THREADS=256:nms.cu,fused.cu" python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-e2e --no-secondary > $O/x7_cfg2_n256.json 2>>$O/x7_n256.err
for f in $O/x7_cfg*.json; do python - "$f" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print(sys.argv[1], round(d["value"]), "img/s", round(d["ms_per_step"] * 1e3, 1), "us/step", {k: round(v, 1) for k, v in d["roofline"]["kernels_us"].items()}, d["check"].get("equals_recorded_1gpu_result"))
except Exception as e:
    print(sys.argv[1], "failed", e)
PY
done
DSN_VARIANT="n256e3:-DDSN_NMS_THREADS=256 -DDSN_SEG_EVAL_CTAS_PER_SM=3:nms.cu,fused.cu,seg.cu" python profiles/experiments/overlap_timeline.py --config 3 --order det 2>&1 | tail -9
for f in $O/x7_*.err; do echo $f; tail -n 3 $f; done
