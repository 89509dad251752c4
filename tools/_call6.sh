set -x
cd $GRAFT_REPO_ROOT
nvidia-smi -L
( time timeout 600 python -m pytest tests/test_dp_nccl_gpu.py -m gpu -q ) > gpurun_out/c6_nccl_tests.txt 2>&1
grep -E "^FAILED|passed|failed|skipped" gpurun_out/c6_nccl_tests.txt; tail -15 gpurun_out/c6_nccl_tests.txt | cut -c1-400
B="--gpus 2 --config cfg-2 --steps 10 --warmup 3 --cpu-baseline 0 --gpu-eager-baseline 0 --through-trainer 0"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py $B > gpurun_out/c6_bench_cfg2_n2.json 2> gpurun_out/c6_bench_cfg2_n2.err
DALM_B200_CROSS_RANK_NEGATIVES=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py $B > gpurun_out/c6_bench_cfg2_n2_xneg.json 2> gpurun_out/c6_bench_cfg2_n2_xneg.err
for f in gpurun_out/c6_bench_*.json; do echo $f; python - "$f" <<'P'
import json,sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], d['ms_per_step'], d['config'].get('launch'), d['config'].get('negatives'), d['config'].get('loss_last'), d['clocks'])
P
done
tail -3 gpurun_out/c6_bench_cfg2_n2_xneg.err | cut -c1-300
