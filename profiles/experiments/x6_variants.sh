O=gpurun_out; mkdir -p $O
B="python bench.py --config 3 --steps 50 --warmup 10 --no-cpu-baseline --no-e2e --no-secondary"
$B > $O/x6_cfg3_base.json 2>$O/x6_base.err
DSN_VARIANT="lg:-DDSN_FILTER_LDGSTS=1:detect.cu" $B > $O/x6_cfg3_lg.json 2>$O/x6_lg.err
for f in $O/x6_cfg3_*.json; do python - "$f" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print(sys.argv[1], round(d["value"]), "img/s", round(d["ms_per_step"] * 1e3, 1), "us/step", {k: round(v, 1) for k, v in d["roofline"]["kernels_us"].items()}, d["check"].get("equals_recorded_1gpu_result"))
except Exception as e:
    print(sys.argv[1], "failed", e)
PY
done
L=$PWD/desenet_b200/lib
python profiles/experiments/time_seg.py 2>&1 | tail -1
for v in px8 px8r16 r16 r64 c6r8 px8c6r8; do DSN_LIB=$L/libdesenet_b200_$v.so python profiles/experiments/time_seg.py 2>&1 | tail -1; done
for n in 2 4; do python bench.py --in-flight $n --steps 100 --warmup 10 --no-cpu-baseline --no-e2e --no-secondary 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('cfg2 in-flight $n', round(d['value']), round(d['ms_per_step']*1e3,2),'us')"; done
for f in $O/x6_*.err; do echo $f; tail -n 3 $f; done
