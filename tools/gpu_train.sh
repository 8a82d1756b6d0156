#!/bin/bash
# PPO on the GPU vec-env: Solo12 Stand then Walk, 4096 envs, bounded wall-clock. Run on a GPU box.
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/runs
python training/train_ppo.py --config-file configs/bench12_walk.yaml --task stand --num-agents 4096 --num-steps 32 \
  --mini-batch-size 16384 --ppo-epoch 5 --lr 3e-4 --use-gae --entropy-coef 0.0 --num-env-steps 2e8 --log-interval 10 \
  --save-interval 50 --max-seconds ${STAND_SECONDS:-200} --logdir gpurun_out/runs --timestamp stand > gpurun_out/train_stand.log 2>&1
tail -12 gpurun_out/train_stand.log
