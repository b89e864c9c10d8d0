for e in 16; do
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-ref-gpu --e2e-steps 1 --opt rank_slots=$e > gpurun_out/s2_e$e.json 2> gpurun_out/s2_e$e.err
python -c "
import json; d=json.load(open('gpurun_out/s2_e$e.json')); print($e, d['value'], d['ms_per_step'], d['stage_ms'], d['roofline']['frac'])"
done
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
