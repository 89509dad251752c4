set -x
cd $GRAFT_REPO_ROOT
( time timeout 600 python -m pytest tests/test_head_chunked_gpu.py tests/test_kernels_gpu.py -m gpu -x -q -k "head or chunk or gelu or l2_hints or raster or ce_rows" ) > gpurun_out/c2_newtests.txt 2>&1
echo "new tests rc=$?" >> gpurun_out/c2_newtests.txt
tail -5 gpurun_out/c2_newtests.txt
( time timeout 900 python -m pytest tests -m gpu -q --durations=8 ) > gpurun_out/c2_gputests.txt 2>&1
echo "tests rc=$?" >> gpurun_out/c2_gputests.txt
tail -4 gpurun_out/c2_gputests.txt
timeout 300 python tools/probe_gemm_l2.py > gpurun_out/c2_l2probe_timing.jsonl 2> gpurun_out/c2_l2probe_timing.err
timeout 400 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:gemm_bf16 --csv --log-file gpurun_out/c2_l2probe_ncu.csv python tools/probe_gemm_l2.py --once > gpurun_out/c2_l2probe_cases.jsonl 2>&1
B="--steps 20 --warmup 5 --cpu-baseline 0 --gpu-eager-baseline 0 --through-trainer 0"
timeout 300 python bench.py $B > gpurun_out/c2_bench_cfg3_default.json 2> gpurun_out/c2_bench_cfg3_default.err
DALM_B200_GEMM_RASTER=-2 timeout 300 python bench.py $B > gpurun_out/c2_bench_cfg3_raster_old.json 2>/dev/null
DALM_B200_CHUNKED_HEAD=0 timeout 300 python bench.py $B > gpurun_out/c2_bench_cfg3_unchunked.json 2>/dev/null
timeout 300 python bench.py --config cfg-2 $B > gpurun_out/c2_bench_cfg2_default.json 2> gpurun_out/c2_bench_cfg2_default.err
DALM_B200_FUSE_GELU=0 timeout 300 python bench.py --config cfg-2 $B > gpurun_out/c2_bench_cfg2_nogelufuse.json 2>/dev/null
for f in gpurun_out/c2_bench_*.json; do echo $f; python - "$f" <<'P'
import json,sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['share_of_step'], d['clocks'])
P
done
