# The command sequence behind profiles/r1_train_step.jsonl, r1q_train_launches.csv and r1q_ncu_*bwd_gemms*.csv
# (BASELINE configs[3]; run from the repository root on a B200 box: gpurun -- 'bash scripts/profile_train_step.sh').
set -x
mkdir -p gpurun_out
rm -f gpurun_out/r1_train_step.jsonl
for a in "--no-fusion" "" "--no-fusion --graph" "--graph"; do
    python scripts/train_step_bench.py $a 2>/dev/null | tail -1 >> gpurun_out/r1_train_step.jsonl
done
cat gpurun_out/r1_train_step.jsonl
# launch list of this library's kernels (3 warm-up steps + 1 timed = 4 identical steps):
#   python scripts/per_step_launches.py gpurun_out/r1q_train_launches.csv 4
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:^k_ --csv \
    --log-file gpurun_out/r1q_train_launches.csv python scripts/train_step_bench.py --steps 1 > gpurun_out/train_ncu.log 2>&1
# every metric of the 37 tensor-core GEMM launches of one backward (≈ 4 minutes: each launch is replayed ~40 times)
ncu --set full --clock-control none --import-source on -k regex:k_bwd_gemm -c 37 -o gpurun_out/r1q_bwd_gemm -f \
    python scripts/train_step_bench.py --steps 1 > gpurun_out/r1q_bwd_ncu.log 2>&1
# two GPUs, DDP (gpurun --gpus 2):
#   python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/train_step_bench.py
