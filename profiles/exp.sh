timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
for v in 0 1 2; do
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-ref-gpu --e2e-steps 1 --opt map_prefetch=$v > gpurun_out/s2_v$v.json 2> gpurun_out/s2_v$v.err
python -c "
import json; d=json.load(open('gpurun_out/s2_v$v.json')); print('prefetch', $v, d['value'], d['ms_per_step'], d['stage_ms'], d['roofline']['frac'])"
done
