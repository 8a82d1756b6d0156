"""Helpers with the semantics of the reference's ``agents/utils.py``."""
import os

import numpy as np
import torch.nn as nn


def init_layer(layer):
    """Orthogonal weights with gain sqrt(2), zero bias (agents/utils.py:122-132)."""
    nn.init.orthogonal_(layer.weight.data, gain=2 ** 0.5)
    if layer.bias is not None:
        nn.init.zeros_(layer.bias.data)
    return layer


def update_linear_schedule(optimizer, epoch, total_num_epochs, initial_lr):
    """lr = lr0 * (1 - epoch/total) (agents/utils.py:14-18)."""
    lr = initial_lr * (1.0 - epoch / float(total_num_epochs))
    for group in optimizer.param_groups:
        if hasattr(group["lr"], "fill_"):      # tensor learning rate of a capturable optimizer: keep the tensor
            group["lr"].fill_(lr)
        else:
            group["lr"] = lr


def init_logging(logdir):
    if logdir is None:
        return None
    savedir = os.path.join(logdir, "checkpoints")
    os.makedirs(savedir, exist_ok=True)
    return savedir


def log(writer, values, name, step):
    """TensorBoard scalar / min-max-mean of a sequence / dict of sequences (agents/utils.py:20-35)."""
    if writer is None:
        return
    if np.isscalar(values):
        writer.add_scalar(name, values, step)
    elif isinstance(values, dict):
        for k, v in values.items():
            writer.add_scalar(name + k, float(np.mean(v)), step)
    elif len(values):
        writer.add_scalar(name + "/min", float(np.min(values)), step)
        writer.add_scalar(name + "/max", float(np.max(values)), step)
        writer.add_scalar(name + "/mean", float(np.mean(values)), step)
