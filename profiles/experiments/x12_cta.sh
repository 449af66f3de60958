O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x > $O/x12_pytest.log 2>&1; tail -3 $O/x12_pytest.log
for c in 2 4; do python bench.py --config $c --steps 60 --warmup 10 --no-cpu-baseline --no-e2e --no-secondary > $O/x12_cfg$c.json 2>$O/x12_cfg$c.err; done
for f in $O/x12_cfg?.json; do python - "$f" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print(sys.argv[1], round(d["value"]), "img/s", round(d["ms_per_step"] * 1e3, 1), "us/step", {k: round(v, 1) for k, v in d["roofline"]["kernels_us"].items()}, d["check"].get("equals_recorded_1gpu_result"))
except Exception as e:
    print(sys.argv[1], "failed", e)
PY
done
for f in $O/x12_*.err; do echo $f; tail -n 3 $f; done
