set -x
cd $GRAFT_REPO_ROOT
( time timeout 900 python -m pytest tests -m gpu -q --durations=6 ) > gpurun_out/c3_gputests.txt 2>&1
echo "tests rc=$?" >> gpurun_out/c3_gputests.txt
grep -E "^FAILED|^ERROR|passed|failed" gpurun_out/c3_gputests.txt
timeout 300 python tools/probe_epilogue_fusions.py > gpurun_out/c3_fusion_probe.jsonl 2> gpurun_out/c3_fusion_probe.err
cat gpurun_out/c3_fusion_probe.jsonl
timeout 400 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm_bf16_tn -s 195 -c 4 -f -o gpurun_out/r02b_gemm_v5_full python tools/profile_step.py cfg-3 > gpurun_out/c3_ncu_gemm.log 2>&1
echo "ncu rc=$?"
B="--steps 20 --warmup 5 --cpu-baseline 0 --gpu-eager-baseline 0 --through-trainer 0"
timeout 300 python bench.py $B > gpurun_out/c3_bench_cfg3_default.json 2> gpurun_out/c3_bench_cfg3_default.err
DALM_B200_FUSE_SWIGLU_BWD=0 timeout 300 python bench.py $B > gpurun_out/c3_bench_cfg3_noswiglubwd.json 2>/dev/null
DALM_B200_GEMM_L2_HINTS=0 timeout 300 python bench.py $B > gpurun_out/c3_bench_cfg3_nohints.json 2>/dev/null
for f in gpurun_out/c3_bench_*.json; do echo $f; python - "$f" <<'P'
import json,sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['share_of_step'], d['clocks'])
P
done
