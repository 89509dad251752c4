set -x
cd $GRAFT_REPO_ROOT
( time timeout 900 python -m pytest tests -x -q -m gpu --durations=5 ) > gpurun_out/c5_gputests.txt 2>&1
echo "tests rc=$?" >> gpurun_out/c5_gputests.txt
grep -E "^FAILED|^ERROR|passed|failed|rc=" gpurun_out/c5_gputests.txt
( time timeout 300 python -c "import __graft_entry__ as g; g.build(); g.smoke()" ) > gpurun_out/c5_smoke.txt 2>&1
tail -4 gpurun_out/c5_smoke.txt
( time timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/c5_bench_default.json 2> gpurun_out/c5_bench_default.err
tail -c 600 gpurun_out/c5_bench_default.err
B="--steps 20 --warmup 5 --cpu-baseline 0 --gpu-eager-baseline 0 --through-trainer 0"
timeout 300 python bench.py --config cfg-2 $B > gpurun_out/c5_bench_cfg2.json 2>/dev/null
DALM_B200_BENCH_NF4=1 timeout 300 python bench.py $B > gpurun_out/c5_bench_cfg3_nf4.json 2>/dev/null
for f in gpurun_out/c5_bench_*.json; do echo $f; python - "$f" <<'P'
import json,sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['share_of_step'], d['roofline'].get('traffic'), d.get('peak_mem_gb'), (d.get('trainer') or d.get('through_trainer') or {}), d['clocks'])
P
done
