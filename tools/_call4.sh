set -x
cd $GRAFT_REPO_ROOT
( DALM_B200_EXPERIMENTAL=1 timeout 600 python -m pytest tests/test_generate_gpu.py -m gpu -q ) > gpurun_out/c4_decode_exp_tests.txt 2>&1
grep -E "^FAILED|^ERROR|passed|failed" gpurun_out/c4_decode_exp_tests.txt
timeout 300 python tools/bench_decode.py > gpurun_out/c4_decode_v1.json 2> gpurun_out/c4_decode_v1.err
DALM_B200_DECODE_ATTN=2 timeout 300 python tools/bench_decode.py > gpurun_out/c4_decode_v2.json 2> gpurun_out/c4_decode_v2.err
cat gpurun_out/c4_decode_v1.json gpurun_out/c4_decode_v2.json | cut -c1-600
B="--cpu-baseline 0 --gpu-eager-baseline 0 --through-trainer 0"
DALM_B200_BENCH_NF4=1 timeout 300 python bench.py --steps 20 --warmup 5 $B > gpurun_out/c4_bench_cfg3_nf4.json 2> gpurun_out/c4_bench_cfg3_nf4.err
timeout 400 python bench.py --config cfg-5 --steps 3 --warmup 3 $B > gpurun_out/c4_bench_cfg5_default.json 2> gpurun_out/c4_bench_cfg5_default.err
DALM_B200_FUSE_GELU_MIN_K=1000000 DALM_B200_CHUNKED_HEAD=0 timeout 400 python bench.py --config cfg-5 --steps 3 --warmup 3 $B > gpurun_out/c4_bench_cfg5_r2a.json 2>/dev/null
for f in gpurun_out/c4_bench_*.json; do echo $f; python - "$f" <<'P'
import json,sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['share_of_step'], d.get('peak_mem_gb'), d['config'].get('nf4_store_gb'), d['clocks'])
P
done
tail -3 gpurun_out/c4_bench_cfg3_nf4.err
