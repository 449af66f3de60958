# Re-measures what changed after the round's full suite (profiles/measure_round.sh): the cfg 3 filter.  Same commands, cfg 2 / cfg 3 only.
R=${1:-r2}
O=gpurun_out; mkdir -p $O
timeout 900 python bench.py > $O/${R}_bench_cfg2.json 2> $O/${R}_bench_cfg2.err
timeout 900 python bench.py --config 3 --steps 50 --warmup 10 > $O/${R}_bench_cfg3.json 2> $O/${R}_bench_cfg3.err
timeout 900 python bench.py --config 1 --steps 50 --warmup 10 > $O/${R}_bench_cfg1.json 2> $O/${R}_bench_cfg1.err
for f in $O/${R}_bench_cfg[123].json; do python - "$f" <<'PY'
import json, sys
d = json.load(open(sys.argv[1]))
print(sys.argv[1], round(d["value"]), "img/s", round(d["ms_per_step"] * 1e3, 1), "us/step", "roofline", round(d["roofline"]["frac"], 3),
      {k: round(v, 1) for k, v in d["roofline"]["kernels_us"].items()}, "e2e", round(d["e2e"]["value"]) if d.get("e2e") else None,
      d["clocks"]["reasons"], d["check"].get("equals_recorded_1gpu_result"), "secondary", round((d.get("secondary") or {}).get("value", 0)))
PY
done
c=3; bs=512
python profiles/prof_kernels.py --config $c --bs $bs --reps 2 > $O/${R}_prof_plain_cfg$c.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'filter_|nms_kernel|seg_|conf_cut|detect_fused' -c 10 \
    -o $O/${R}_ncu_full_cfg$c -f python profiles/prof_kernels.py --config $c --bs $bs --reps 2 > $O/${R}_ncu_full_cfg$c.log 2>&1
ncu -i $O/${R}_ncu_full_cfg$c.ncu-rep --page raw --csv > $O/${R}_ncu_full_cfg${c}_raw.csv 2>/dev/null
ncu -i $O/${R}_ncu_full_cfg$c.ncu-rep --page source --csv -k regex:filter_dense_tma --launch-skip 0 --launch-count 1 > $O/${R}_ncu_src_cfg${c}_filter_dense_tma.csv 2>/dev/null
gzip -f $O/${R}_ncu_src_cfg${c}_filter_dense_tma.csv
rm -f $O/${R}_ncu_full_cfg$c.ncu-rep
ls -la $O | grep ${R}_ | wc -l
