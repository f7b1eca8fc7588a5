set -x
tests/microbench/exp_loop2 > gpurun_out/exp_loop2.txt 2>&1
for p in 0 2 3 4; do
  LIDM_ATTN_POLY=$p python tests/attn_cases.py 2,8,2048 2,16,512 > gpurun_out/attn_poly$p.txt 2>&1
  LIDM_ATTN_POLY=$p python tests/op_profile.py 64 2>&1 | grep -E "U-Net|attn" > gpurun_out/ops_poly$p.txt
done
cat gpurun_out/exp_loop2.txt gpurun_out/attn_poly*.txt gpurun_out/ops_poly*.txt
