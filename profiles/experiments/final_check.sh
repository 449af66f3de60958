O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/final_pytest.log 2>&1; tail -2 $O/final_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | cut -c1-200
timeout 900 python bench.py 2>/dev/null > $O/final_bench.json; python -c "
import json; d=json.load(open('gpurun_out/final_bench.json')); print(round(d['value']), round(d['ms_per_step']*1e3,2), d['roofline']['frac'], d['e2e']['value'], d['clocks'], d['gpu_launches'], d['check'])"
