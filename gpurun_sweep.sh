timeout 300 python -m pytest tests/test_gpu_icp.py -x -q 2>&1 | tail -2
for rep in 1 2; do timeout 300 python bench.py --steps 128 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value'],1), round(d['e2e']['value'],1), round(d['latency']['ms_per_alignment_device'],3), d['roofline']['per_alignment_ms'])"; done
