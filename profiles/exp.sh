timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
timeout 600 python profiles/compare_ref.py --no-ref --out gpurun_out/s2_cmp_rmat24b.jsonl --iters 5 --warmup 3 rmat24 > /dev/null 2>&1
python -c "
import json
for l in open('gpurun_out/s2_cmp_rmat24b.jsonl'):
    d=json.loads(l); o=d['ours']; print(' ', d['workload'], round(o['mean_ms'],3), round(o['gflops_mean'],1), {k:round(x,3) for k,x in o['stage_ms'].items()}, o['idx_bit_exact_vs_oracle'])"
