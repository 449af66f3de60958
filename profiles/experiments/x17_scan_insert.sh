O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x > $O/x17_pytest.log 2>&1; tail -3 $O/x17_pytest.log
B="--steps 50 --warmup 10 --no-cpu-baseline --no-e2e --no-secondary"
python bench.py --config 3 $B > $O/x17_cfg3.json 2>$O/x17_a.err
for f in $O/x17_cfg*.json; do python - "$f" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print(sys.argv[1], round(d["value"]), "img/s", round(d["ms_per_step"] * 1e3, 1), "us/step", {k: round(v, 1) for k, v in d["roofline"]["kernels_us"].items()}, d["check"].get("equals_recorded_1gpu_result"))
except Exception as e:
    print(sys.argv[1], "failed", e)
PY
done
tail -n 3 $O/x17_a.err
