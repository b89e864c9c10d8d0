timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -8
for m in 1; do
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-ref-gpu --e2e-steps 1 --opt rank_map=$m > gpurun_out/s2_m$m.json 2> gpurun_out/s2_m$m.err
python -c "
import json; d=json.load(open('gpurun_out/s2_m$m.json')); print($m, d['value'], d['ms_per_step'], d['stage_ms'], d['roofline']['frac'])"
done
timeout 300 bash profiles/launch_list.sh s2_map3 --no-ref-gpu
