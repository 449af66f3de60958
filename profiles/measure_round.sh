#!/bin/bash
# Whole measurement suite of a round, run on the GPU box:  gpurun -- 'bash profiles/measure_round.sh r1'
# Writes into gpurun_out/; profiles/summarize.py (run in the build container) turns that into profiles/*.
R=${1:-r1}
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/${R}_pytest_gpu.log 2>&1; tail -2 $O/${R}_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/${R}_smoke.log 2>&1; tail -1 $O/${R}_smoke.log
python bench.py --impl reference --steps 3 --warmup 1 > $O/${R}_bench_cfg2_reference_arm.json 2> $O/${R}_ref.err
python bench.py > $O/${R}_bench_cfg2.json 2> $O/${R}_bench_cfg2.err
for c in 3 4 5; do python bench.py --config $c --steps 50 --warmup 10 > $O/${R}_bench_cfg$c.json 2> $O/${R}_bench_cfg$c.err; done
for f in $O/${R}_bench_cfg?.json; do python - "$f" <<'PY'
import json, sys
d = json.load(open(sys.argv[1]))
print(sys.argv[1], round(d["value"]), "img/s", round(d["ms_per_step"] * 1e3, 1), "us/step", "roofline", round(d["roofline"]["frac"], 3),
      "e2e", round(d["e2e"]["value"]) if d.get("e2e") else None, d["clocks"]["reasons"])
PY
done
# launch list of the default bench command (plain run above exited 0); shares, not absolutes, are comparable
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/${R}_launches_cfg2.csv \
    python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > $O/${R}_ncu_launches.log 2>&1
# one --set full capture per hot kernel and config (L2 flushed between repetitions by prof_kernels.py)
for c in 2 3 5; do
  bs=64; [ $c = 5 ] && bs=32
  python profiles/prof_kernels.py --config $c --bs $bs --reps 2 > $O/${R}_prof_plain_cfg$c.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:'filter_|nms_kernel|seg_fused|conf_cut' -c 8 \
      -o $O/${R}_ncu_full_cfg$c -f python profiles/prof_kernels.py --config $c --bs $bs --reps 2 > $O/${R}_ncu_full_cfg$c.log 2>&1
  ncu -i $O/${R}_ncu_full_cfg$c.ncu-rep --page raw --csv > $O/${R}_ncu_full_cfg${c}_raw.csv 2>/dev/null
done
ls -la $O | grep ${R}_ | head -40
