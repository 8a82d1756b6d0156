#!/bin/bash
# PPO on the GPU vec-env, Solo12, 4096 envs, bounded wall-clock.  Usage: TASK=stand|walk SECONDS_MAX=200 tools/gpu_train.sh
cd "$(dirname "$0")/.."
TASK=${TASK:-stand}
mkdir -p gpurun_out/runs
python training/train_ppo.py --config-file configs/bench12_walk.yaml --task $TASK --num-agents 4096 --num-steps 32 \
  --mini-batch-size 16384 --ppo-epoch 5 --lr 3e-4 --use-gae --entropy-coef 0.0 --num-env-steps 1e9 --log-interval 20 \
  --save-interval 100 --max-seconds ${SECONDS_MAX:-200} --logdir gpurun_out/runs --timestamp $TASK > gpurun_out/train_$TASK.log 2>&1
grep -A1 "^Updates" gpurun_out/train_$TASK.log | grep -E "Updates|training" | paste - - | awk 'NR%6==1{print $2, $5, $7, $12, $13}'
tail -3 gpurun_out/train_$TASK.log
