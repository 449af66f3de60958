O=gpurun_out; mkdir -p $O
timeout 120 python profiles/experiments/prof_detconv.py > $O/x9_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'detect_head_conv_umma' -c 6 -o $O/x9_detconv -f python profiles/experiments/prof_detconv.py > $O/x9_ncu.log 2>&1
ncu -i $O/x9_detconv.ncu-rep --page raw --csv > $O/x9_detconv_raw.csv 2>/dev/null
ncu -i $O/x9_detconv.ncu-rep --page source --csv --launch-skip 1 --launch-count 1 > $O/x9_detconv_src.csv 2>/dev/null; gzip -f $O/x9_detconv_src.csv
rm -f $O/x9_detconv.ncu-rep
tail -3 $O/x9_ncu.log
