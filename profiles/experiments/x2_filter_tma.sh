O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x > $O/x2_pytest.log 2>&1; tail -3 $O/x2_pytest.log
B="python bench.py --config 3 --steps 50 --warmup 10 --no-cpu-baseline --no-e2e --no-secondary"
$B > $O/x2_cfg3_base.json 2>$O/x2_base.err
DSN_VARIANT="s2:-DDSN_TMA_STAGES=2:detect.cu" $B > $O/x2_cfg3_s2.json 2>$O/x2_s2.err
DSN_VARIANT="s2w8:-DDSN_TMA_STAGES=2 -DDSN_TMA_WARPS=8:detect.cu" $B > $O/x2_cfg3_s2w8.json 2>$O/x2_s2w8.err
DSN_VARIANT="s4:-DDSN_TMA_STAGES=4:detect.cu" $B > $O/x2_cfg3_s4.json 2>$O/x2_s4.err
for f in $O/x2_cfg3_*.json; do python - "$f" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print(sys.argv[1], round(d["value"]), "img/s", round(d["ms_per_step"] * 1e3, 1), "us/step", {k: round(v, 1) for k, v in d["roofline"]["kernels_us"].items()}, d["check"].get("equals_recorded_1gpu_result"))
except Exception as e:
    print(sys.argv[1], "failed", e)
PY
done
python profiles/experiments/overlap_timeline.py --config 3 --order det 2>&1 | tail -12
python profiles/experiments/overlap_timeline.py --config 3 --order seg 2>&1 | tail -12
for f in $O/x2_*.err; do echo $f; tail -n 3 $f; done
