O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x > $O/x4_pytest.log 2>&1; tail -3 $O/x4_pytest.log
B="python bench.py --config 3 --steps 50 --warmup 10 --no-cpu-baseline --no-e2e --no-secondary"
$B > $O/x4_cfg3_base.json 2>$O/x4_base.err
DSN_VARIANT="c6:-DDSN_TMA_MIN_CTAS=6:detect.cu" $B > $O/x4_cfg3_c6.json 2>$O/x4_c6.err
DSN_VARIANT="c4:-DDSN_TMA_MIN_CTAS=4:detect.cu" $B > $O/x4_cfg3_c4.json 2>$O/x4_c4.err
DSN_VARIANT="s8:-DDSN_TMA_STAGES=8:detect.cu" $B > $O/x4_cfg3_s8.json 2>$O/x4_s8.err
DSN_VARIANT="w4c12:-DDSN_TMA_WARPS=4 -DDSN_TMA_MIN_CTAS=12:detect.cu" $B > $O/x4_cfg3_w4c12.json 2>$O/x4_w4c12.err
for f in $O/x4_cfg3_*.json; do python - "$f" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print(sys.argv[1], round(d["value"]), "img/s", round(d["ms_per_step"] * 1e3, 1), "us/step", {k: round(v, 1) for k, v in d["roofline"]["kernels_us"].items()}, d["check"].get("equals_recorded_1gpu_result"))
except Exception as e:
    print(sys.argv[1], "failed", e)
PY
done
ncu --set full --clock-control none --import-source on -k regex:'filter_dense_tma' -c 2 -o $O/x4_ncu_filter -f python profiles/prof_kernels.py --config 3 --bs 512 --reps 2 > $O/x4_ncu.log 2>&1
ncu -i $O/x4_ncu_filter.ncu-rep --page raw --csv > $O/x4_ncu_filter_raw.csv 2>/dev/null
ncu -i $O/x4_ncu_filter.ncu-rep --page source --csv --launch-skip 1 --launch-count 1 > $O/x4_ncu_filter_src.csv 2>/dev/null; gzip -f $O/x4_ncu_filter_src.csv
rm -f $O/x4_ncu_filter.ncu-rep
for f in $O/x4_*.err; do echo $f; tail -n 3 $f; done
