O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x > $O/x16_pytest.log 2>&1; tail -3 $O/x16_pytest.log
B="--steps 50 --warmup 10 --no-cpu-baseline --no-e2e --no-secondary"
python bench.py --config 3 $B > $O/x16_cfg3_tensor.json 2>$O/x16_a.err
DSN_VARIANT="bulk:-DDSN_FILTER_TENSORMAP=0:detect.cu" python bench.py --config 3 $B > $O/x16_cfg3_bulk.json 2>$O/x16_b.err
python bench.py --config 1 $B > $O/x16_cfg1_tensor.json 2>>$O/x16_a.err
for f in $O/x16_cfg*.json; do python - "$f" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print(sys.argv[1], round(d["value"]), "img/s", round(d["ms_per_step"] * 1e3, 1), "us/step", {k: round(v, 1) for k, v in d["roofline"]["kernels_us"].items()}, d["check"].get("equals_recorded_1gpu_result"))
except Exception as e:
    print(sys.argv[1], "failed", e)
PY
done
for f in $O/x16_*.err; do echo $f; tail -n 3 $f; done
