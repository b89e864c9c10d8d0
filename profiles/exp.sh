for v in 1 2 4; do
echo "num_streams=$v"
timeout 300 python profiles/compare_ref.py --no-ref --out gpurun_out/s2_cmp_s$v.jsonl --opt num_streams=$v econ_like circuit_like webbase_like cant_like banded_like > /dev/null 2>&1
python -c "
import json
for l in open('gpurun_out/s2_cmp_s$v.jsonl'):
    d=json.loads(l); o=d['ours']; print(' ', d['workload'], round(o['mean_ms'],3), {k:round(x,3) for k,x in o['stage_ms'].items()})"
done
timeout 120 bash profiles/launch_list.sh s2_webbase --no-ref-gpu --workload webbase_like --seed 3
