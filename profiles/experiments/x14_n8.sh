O=gpurun_out; mkdir -p $O
nvidia-smi -L | wc -l
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 100 --warmup 10 > $O/r2_bench_cfg2_n8.json 2> $O/x14_n8.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2_bench_cfg2_n8.json"))
print("n8", round(d["value"]), round(d["ms_per_step"]*1e3, 2), "us", d["check"], "e2e", round(d["e2e"]["value"]), {k: round(v["value"]) for k, v in d["e2e"].items() if isinstance(v, dict) and "value" in v}, d["e2e"].get("host"))
s = d.get("secondary") or {}
print("secondary", round(s.get("value", 0)), round(s.get("ms_per_step", 0)*1e3, 1), "us", s.get("check"), s.get("roofline", {}).get("kernels_us"))
PY
tail -n 2 $O/x14_n8.err
