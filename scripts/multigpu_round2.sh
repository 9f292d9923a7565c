# BASELINE configs[3] (train step, DDP) and configs[4] (batch sweep incl. NMS) at 1 / 2 / 4 / 8 GPUs of one box.
# gpurun --gpus 8 -- 'bash scripts/multigpu_round2.sh'
set -x
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
: > gpurun_out/r2g_train_ddp.jsonl
: > gpurun_out/r2g_sweep.jsonl
python bench.py --workload train --steps 10 --no-graph 2>/dev/null | tail -1 >> gpurun_out/r2g_train_ddp.jsonl
for N in 2 4 8; do
  $T --nproc-per-node $N --master-port $((29500 + N)) bench.py --workload train --gpus $N --steps 10 2>/dev/null | grep '^{' | tail -1 >> gpurun_out/r2g_train_ddp.jsonl
done
python scripts/sweep_batch.py --batches 1,8,64 --steps 5 2>/dev/null | grep '^{' >> gpurun_out/r2g_sweep.jsonl
for N in 2 4 8; do
  $T --nproc-per-node $N --master-port $((29600 + N)) scripts/sweep_batch.py --batches 1,8,64 --steps 5 2>/dev/null | grep '^{' >> gpurun_out/r2g_sweep.jsonl
done
cat gpurun_out/r2g_train_ddp.jsonl gpurun_out/r2g_sweep.jsonl
