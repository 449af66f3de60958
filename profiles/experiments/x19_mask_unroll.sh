L=$PWD/desenet_b200/lib
python profiles/experiments/time_seg.py 2>&1 | tail -1
for v in u1 u4 u4r16; do DSN_LIB=$L/libdesenet_b200_$v.so python profiles/experiments/time_seg.py 2>&1 | tail -1; done
python profiles/experiments/time_seg.py 2>&1 | tail -1
