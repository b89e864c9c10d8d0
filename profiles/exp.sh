timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -5
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-ref-gpu --e2e-steps 1 > gpurun_out/s2_b.json 2> gpurun_out/s2_b.err
python -c "
import json; d=json.load(open('gpurun_out/s2_b.json')); print(d['value'], d['ms_per_step'], d['stage_ms'], d['roofline']['frac'])"
