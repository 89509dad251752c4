set -x
cd $GRAFT_REPO_ROOT
( time timeout 900 python -m pytest tests -m gpu -x -q --durations=15 ) > gpurun_out/c1_gputests.txt 2>&1
echo "tests rc=$?" >> gpurun_out/c1_gputests.txt
timeout 400 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm_bf16_tn -s 195 -c 4 -f -o gpurun_out/r02_gemm_v4_full python tools/profile_step.py cfg-3 > gpurun_out/c1_ncu_gemm.log 2>&1
echo "ncu gemm rc=$?" >> gpurun_out/c1_ncu_gemm.log
timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:attn_ -s 46 -c 8 -f -o gpurun_out/r02_attn64_full python tools/profile_step.py cfg-2 > gpurun_out/c1_ncu_attn.log 2>&1
echo "ncu attn rc=$?" >> gpurun_out/c1_ncu_attn.log
tail -5 gpurun_out/c1_gputests.txt; tail -3 gpurun_out/c1_ncu_gemm.log; tail -3 gpurun_out/c1_ncu_attn.log
ls -la gpurun_out
