O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x > $O/x3_pytest.log 2>&1; tail -3 $O/x3_pytest.log
B="python bench.py --config 3 --steps 50 --warmup 10 --no-cpu-baseline --no-e2e --no-secondary"
$B > $O/x3_cfg3_base.json 2>$O/x3_base.err
DSN_VARIANT="s2w4:-DDSN_TMA_STAGES=2 -DDSN_TMA_WARPS=4:detect.cu" $B > $O/x3_cfg3_s2w4.json 2>$O/x3_s2w4.err
DSN_VARIANT="s3w4:-DDSN_TMA_STAGES=3 -DDSN_TMA_WARPS=4:detect.cu" $B > $O/x3_cfg3_s3w4.json 2>$O/x3_s3w4.err
DSN_VARIANT="s2w12:-DDSN_TMA_STAGES=2 -DDSN_TMA_WARPS=12:detect.cu" $B > $O/x3_cfg3_s2w12.json 2>$O/x3_s2w12.err
for f in $O/x3_cfg3_*.json; do python - "$f" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print(sys.argv[1], round(d["value"]), "img/s", round(d["ms_per_step"] * 1e3, 1), "us/step", {k: round(v, 1) for k, v in d["roofline"]["kernels_us"].items()}, d["check"].get("equals_recorded_1gpu_result"))
except Exception as e:
    print(sys.argv[1], "failed", e)
PY
done
python profiles/experiments/nms_timeline.py 3 64 2>&1 | tail -4
python profiles/experiments/nms_timeline.py 2 64 2>&1 | tail -3
for f in $O/x3_*.err; do echo $f; tail -n 3 $f; done
