#!/bin/bash
# train.py under torchrun with --dist_mode featpart on N GPUs: 3 epochs + checkpoints, then resume for a 4th; the checkpoint must be full-width
N=${NGPU:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
ARGS="--synthetic tiny --bpr_batch 256 --checkpoint_dir gpurun_out/ck_feat --save_every 1 --eval_every 2 --dist_mode featpart"
timeout 300 $TR --master-port 29541 -m lgcn_b200.train $ARGS --epochs 3 > gpurun_out/train_feat_a.log 2>&1; echo "train rc=$?"; grep "EPOCH\|Error\|error" gpurun_out/train_feat_a.log | tail -5
timeout 300 $TR --master-port 29543 -m lgcn_b200.train $ARGS --epochs 4 --resume gpurun_out/ck_feat/last.pth.tar > gpurun_out/train_feat_b.log 2>&1; echo "resume rc=$?"; grep "EPOCH\|Error\|error" gpurun_out/train_feat_b.log | tail -3
python - <<'PY'
import torch
ck = torch.load('gpurun_out/ck_feat/last.pth.tar', map_location='cpu', weights_only=False)
print('epoch', ck['epoch'], {k: tuple(v.shape) for k, v in ck['model_state'].items()},
      'opt', tuple(ck['optimizer_state']['state'][0]['exp_avg'].shape), float(ck['optimizer_state']['state'][0]['step']), 'best', ck['best_metric'])
PY
rm -rf gpurun_out/ck_feat
