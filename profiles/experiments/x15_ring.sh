O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x -k "seg or Seg or eval or pipeline or smoke or dropin" > $O/x15_pytest.log 2>&1; tail -3 $O/x15_pytest.log
B="--steps 50 --warmup 10 --no-cpu-baseline --no-e2e --no-secondary"
python bench.py --config 3 $B > $O/x15_cfg3_ring.json 2>$O/x15_a.err
python bench.py --config 4 $B > $O/x15_cfg4_ring.json 2>>$O/x15_a.err
DSN_VARIANT="noring:-DDSN_SEG_EVAL_RING=0:seg.cu" python bench.py --config 3 $B > $O/x15_cfg3_noring.json 2>$O/x15_b.err
for f in $O/x15_cfg*.json; do python - "$f" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print(sys.argv[1], round(d["value"]), "img/s", round(d["ms_per_step"] * 1e3, 1), "us/step", {k: round(v, 1) for k, v in d["roofline"]["kernels_us"].items()}, d["check"].get("equals_recorded_1gpu_result"), d["check"].get("cm"))
except Exception as e:
    print(sys.argv[1], "failed", e)
PY
done
python profiles/experiments/overlap_timeline.py --config 3 --order det 2>&1 | tail -9
for f in $O/x15_*.err; do echo $f; tail -n 3 $f; done
