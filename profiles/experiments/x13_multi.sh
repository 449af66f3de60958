O=gpurun_out; mkdir -p $O
nvidia-smi -L | head -4
timeout 600 python -m pytest tests/test_distributed_nccl.py -m gpu -q > $O/x13_nccl_test.log 2>&1; tail -2 $O/x13_nccl_test.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 100 --warmup 10 > $O/r2_bench_cfg2_n2.json 2> $O/x13_n2.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2_bench_cfg2_n2.json"))
print("n2", round(d["value"]), round(d["ms_per_step"]*1e3, 2), "us", d["check"], "e2e", round(d["e2e"]["value"]), {k: round(v["value"]) for k, v in d["e2e"].items() if isinstance(v, dict) and "value" in v})
s = d.get("secondary") or {}
print("secondary", s.get("config", {}).get("workload"), round(s.get("value", 0)), round(s.get("ms_per_step", 0)*1e3, 1), "us", s.get("check"))
PY
timeout 600 python bench.py --config 4 --steps 50 --warmup 10 > $O/r2_bench_cfg4.json 2> $O/x13_cfg4.err
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_cfg4.json')); print('cfg4', round(d['value']), round(d['ms_per_step']*1e3,1), d['roofline']['kernels_us'])"
tail -n 3 $O/x13_n2.err
