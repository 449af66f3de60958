O=gpurun_out; mkdir -p $O
B="python bench.py --config 3 --steps 50 --warmup 10 --no-cpu-baseline --no-e2e --no-secondary"
for v in "" ev2 ev3; do
  if [ -z "$v" ]; then $B > $O/x1_cfg3_base.json 2>$O/x1_base.err; else DSN_VARIANT="$v:-DDSN_SEG_EVAL_CTAS_PER_SM=${v#ev}:seg.cu" $B > $O/x1_cfg3_$v.json 2>$O/x1_$v.err; fi
done
$B --in-flight 1 > $O/x1_cfg3_base_fl1.json 2>>$O/x1_base.err
DSN_VARIANT="ev2:-DDSN_SEG_EVAL_CTAS_PER_SM=2:seg.cu" $B --in-flight 2 > $O/x1_cfg3_ev2_fl2.json 2>>$O/x1_ev2.err
for f in $O/x1_cfg3_*.json; do python - "$f" <<'PY'
import json, sys
d = json.load(open(sys.argv[1]))
print(sys.argv[1], round(d["value"]), "img/s", round(d["ms_per_step"] * 1e3, 1), "us/step", {k: round(v, 1) for k, v in d["roofline"]["kernels_us"].items()}, d["check"].get("equals_recorded_1gpu_result"))
PY
done
tail -3 $O/x1_*.err
