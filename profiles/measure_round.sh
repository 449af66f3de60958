#!/bin/bash
# Whole measurement suite of a round on ONE B200:  gpurun -- 'bash profiles/measure_round.sh r2'
# Writes into gpurun_out/; profiles/summarize.py (run in the build container afterwards) turns that into profiles/*.
R=${1:-r2}
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/${R}_pytest_gpu.log 2>&1; tail -2 $O/${R}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/${R}_smoke.log 2>&1; tail -1 $O/${R}_smoke.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/${R}_bench_cfg2_reference_arm.json 2> $O/${R}_ref.err
timeout 900 python bench.py > $O/${R}_bench_cfg2.json 2> $O/${R}_bench_cfg2.err
for c in 1 3 4 5; do timeout 900 python bench.py --config $c --steps 50 --warmup 10 > $O/${R}_bench_cfg$c.json 2> $O/${R}_bench_cfg$c.err; done
for f in $O/${R}_bench_cfg?.json; do python - "$f" <<'PY'
import json, sys
d = json.load(open(sys.argv[1]))
print(sys.argv[1], round(d["value"]), "img/s", round(d["ms_per_step"] * 1e3, 1), "us/step", "roofline", round(d["roofline"]["frac"], 3),
      {k: round(v, 1) for k, v in d["roofline"]["kernels_us"].items()}, "e2e", round(d["e2e"]["value"]) if d.get("e2e") else None,
      d["clocks"]["reasons"], d["check"].get("equals_recorded_1gpu_result"))
PY
done
# launch list of the default bench command (the plain run above exited 0); shares, not absolutes, are comparable
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $O/${R}_launches_cfg2.csv \
    python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-secondary > $O/${R}_ncu_launches.log 2>&1
# one --set full capture per hot kernel and config AT THE BENCH'S BATCH SIZE (L2 flushed between repetitions by prof_kernels.py)
for c in 2 3 4 5; do
  bs=64; [ $c = 3 ] && bs=512; [ $c = 4 ] && bs=32; [ $c = 5 ] && bs=128
  python profiles/prof_kernels.py --config $c --bs $bs --reps 2 > $O/${R}_prof_plain_cfg$c.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:'filter_|nms_kernel|seg_|conf_cut|detect_fused' -c 10 \
      -o $O/${R}_ncu_full_cfg$c -f python profiles/prof_kernels.py --config $c --bs $bs --reps 2 > $O/${R}_ncu_full_cfg$c.log 2>&1
  ncu -i $O/${R}_ncu_full_cfg$c.ncu-rep --page raw --csv > $O/${R}_ncu_full_cfg${c}_raw.csv 2>/dev/null
  # source pages (SASS + source line, per-instruction executed / stall samples) of the hot kernels, then drop the
  # report itself: gpurun only brings back 64 MiB and one report is 30-50 MB
  for k in seg_mask_rows seg_fused_rows nms_kernel filter_dense_tma filter_dense_kernel filter_levels; do
    ncu -i $O/${R}_ncu_full_cfg$c.ncu-rep --page source --csv -k regex:$k --launch-skip 0 --launch-count 1 \
        > $O/${R}_ncu_src_cfg${c}_$k.csv 2>/dev/null
    [ -s $O/${R}_ncu_src_cfg${c}_$k.csv ] && [ $(wc -l < $O/${R}_ncu_src_cfg${c}_$k.csv) -gt 20 ] && gzip -f $O/${R}_ncu_src_cfg${c}_$k.csv || rm -f $O/${R}_ncu_src_cfg${c}_$k.csv
  done
  rm -f $O/${R}_ncu_full_cfg$c.ncu-rep $O/${R}_ncu_full_cfg$c.ncu-rep.tmp
done
# Detect head convs (tcgen05 kernel vs the framework's cuDNN convolution): correctness + timing, then one capture per level
timeout 300 python profiles/experiments/check_detconv.py > $O/${R}_detconv.log 2>&1; tail -12 $O/${R}_detconv.log
timeout 120 python profiles/experiments/prof_detconv.py > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'detect_head_conv_umma' -c 6 -o $O/${R}_ncu_full_detconv -f \
    python profiles/experiments/prof_detconv.py > $O/${R}_ncu_full_detconv.log 2>&1
ncu -i $O/${R}_ncu_full_detconv.ncu-rep --page raw --csv > $O/${R}_ncu_full_detconv_raw.csv 2>/dev/null
ncu -i $O/${R}_ncu_full_detconv.ncu-rep --page source --csv --launch-skip 1 --launch-count 1 > $O/${R}_ncu_src_detconv_umma.csv 2>/dev/null
gzip -f $O/${R}_ncu_src_detconv_umma.csv
rm -f $O/${R}_ncu_full_detconv.ncu-rep
du -sh $O; ls -la $O | grep ${R}_ | head -80
