"""Clipped-surrogate PPO with clipped value loss (agents/ppo/ppo.py:6-89), data-parallel over GPUs.

Multi-GPU (one process per GPU, each with its own env shard): the only collectives of the whole
system are here — a 3-scalar all-reduce (count, sum, sum of squares) so that advantage
normalisation uses the statistics of ALL ranks' samples (ppo.py:35-37; torch.std there is the
unbiased N-1 estimator), and one flat-bucket all-reduce of the ~19k-parameter gradient per
mini-batch, inserted before clip_grad_norm_ (ppo.py:72-77).  Over NVLink/NVSwitch both are
latency-bound (76 KB), so the gradient travels as ONE contiguous buffer, one NCCL call."""
import torch
import torch.nn as nn
import torch.optim as optim


def dist_ready():
    return torch.distributed.is_available() and torch.distributed.is_initialized() and \
        torch.distributed.get_world_size() > 1


def global_mean_std(x):
    """Mean and unbiased std of x over all ranks' elements (one 3-scalar all-reduce)."""
    x = x.reshape(-1).double()
    stats = torch.stack([torch.tensor(float(x.numel()), device=x.device, dtype=torch.float64), x.sum(), (x * x).sum()])
    if dist_ready():
        torch.distributed.all_reduce(stats)
    n, s, ss = stats[0], stats[1], stats[2]
    mean = s / n
    var = (ss - n * mean * mean) / (n - 1)
    return mean.float(), var.clamp_min(0).sqrt().float()


class FlatParameters:
    """Every parameter of a module becomes a view of ONE contiguous buffer, and so does every gradient
    (``p.grad`` is pre-set to a view of ``self.grad``; autograd accumulates into an existing ``.grad`` in place and
    ``zero_grad(set_to_none=False)`` keeps it).  Gradients are therefore born contiguous: the data-parallel
    all-reduce is one NCCL call on ``self.grad`` with no pack / unpack kernels around it (round 1 copied 16
    parameter tensors into a bucket and back, ~32 small launches per mini-batch).  ``state_dict`` keys, shapes and
    ``load_state_dict`` (an in-place copy) are unaffected."""

    def __init__(self, module):
        self.params = [p for p in module.parameters() if p.requires_grad]
        p0 = self.params[0]
        self.numel = sum(p.numel() for p in self.params)
        self.data = torch.zeros(self.numel, dtype=p0.dtype, device=p0.device)
        self.grad = torch.zeros(self.numel, dtype=p0.dtype, device=p0.device)
        off = 0
        with torch.no_grad():
            for p in self.params:
                n = p.numel()
                self.data[off:off + n].copy_(p.data.reshape(-1))
                p.data = self.data[off:off + n].view_as(p)
                p.grad = self.grad[off:off + n].view_as(p)
                off += n

    def intact(self):
        """True while every parameter and gradient still aliases the flat buffers (a ``.to()`` / ``set_to_none``
        would silently detach them)."""
        off = 0
        for p in self.params:
            n = p.numel()
            if p.data_ptr() != self.data[off:off + n].data_ptr() or p.grad is None or \
                    p.grad.data_ptr() != self.grad[off:off + n].data_ptr():
                return False
            off += n
        return True

    def zero_grad(self):
        self.grad.zero_()                         # one launch instead of one per tensor

    def all_reduce_mean(self):
        """Average the gradient over the ranks: ONE collective, in place, no copies."""
        if dist_ready():
            torch.distributed.all_reduce(self.grad)
            self.grad.div_(torch.distributed.get_world_size())


class FlatGradAllReduce:
    """Average gradients across ranks through one contiguous bucket (pack / all-reduce / unpack): the fallback
    for parameter lists that are not views of a FlatParameters buffer."""

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        self.numel = sum(p.numel() for p in self.params)
        self.flat = None

    def __call__(self):
        if not dist_ready():
            return
        p0 = self.params[0]
        if self.flat is None or self.flat.device != p0.device:
            self.flat = torch.zeros(self.numel, dtype=p0.dtype, device=p0.device)
        off = 0
        for p in self.params:
            n = p.numel()
            if p.grad is None:
                self.flat[off:off + n].zero_()
            else:
                self.flat[off:off + n].copy_(p.grad.reshape(-1))
            off += n
        torch.distributed.all_reduce(self.flat)
        self.flat.div_(torch.distributed.get_world_size())
        off = 0
        for p in self.params:
            n = p.numel()
            if p.grad is None:
                p.grad = torch.empty_like(p)
            p.grad.copy_(self.flat[off:off + n].view_as(p))
            off += n


def broadcast_parameters(module, src=0):
    if dist_ready():
        for t in list(module.parameters()) + list(module.buffers()):
            torch.distributed.broadcast(t.data, src)


class PPO:
    def __init__(self, actor_critic, clip_param, ppo_epoch, mini_batch_size, value_loss_coef, entropy_coef,
                 lr=None, l2_coef=0.0, max_grad_norm=None, use_clipped_value_loss=True, use_graph=True,
                 flat_parameters=True):
        self.actor_critic = actor_critic
        self.clip_param = clip_param
        self.ppo_epoch = ppo_epoch
        self.mini_batch_size = mini_batch_size
        self.value_loss_coef = value_loss_coef
        self.entropy_coef = entropy_coef
        self.max_grad_norm = max_grad_norm
        self.use_clipped_value_loss = use_clipped_value_loss
        p0 = next(actor_critic.parameters())
        self.use_graph = bool(use_graph and p0.is_cuda)
        # parameters and gradients as views of two flat buffers (built after the module sits on its device)
        self.flat = FlatParameters(actor_critic) if flat_parameters else None
        self.sync_enabled = True                  # bench.py switches the collective off to measure its share
        if p0.is_cuda:
            # fused Adam: one multi-tensor launch instead of ~10 per parameter tensor (same arithmetic); capturable
            # with a tensor learning rate so that the mini-batch step can be replayed as a CUDA graph while the
            # linear schedule keeps changing the rate in place
            self.optimizer = optim.Adam(actor_critic.parameters(), lr=torch.tensor(float(lr), device=p0.device),
                                        weight_decay=l2_coef, fused=True, capturable=True)
        else:
            self.optimizer = optim.Adam(actor_critic.parameters(), lr=lr, weight_decay=l2_coef)
        self.grad_sync = FlatGradAllReduce(actor_critic.parameters())
        self._graph = None
        self._graph_key = None
        self._warm = 0
        # multi-GPU: the gradient all-reduce is captured inside the graph (NCCL collectives are capturable);
        # SOLO_PPO_GRAPH_NCCL=0 keeps the multi-GPU update eager
        import os
        self.graph_with_nccl = os.environ.get("SOLO_PPO_GRAPH_NCCL", "1") != "0"

    # ---- one mini-batch step (ppo.py:47-81) on given batch tensors ----------------------------------------
    def _step(self, obs, actions, old_values, returns, old_logp, adv_b, sums):
        values, logp, entropy = self.actor_critic.evaluate_actions(obs, actions)
        ratio = torch.exp(logp - old_logp)
        surr = torch.min(ratio * adv_b, ratio.clamp(1.0 - self.clip_param, 1.0 + self.clip_param) * adv_b)
        action_loss = -surr.mean()
        if self.use_clipped_value_loss:
            clipped = old_values + (values - old_values).clamp(-self.clip_param, self.clip_param)
            value_loss = 0.5 * torch.max((values - returns).pow(2), (clipped - returns).pow(2)).mean()
        else:
            value_loss = 0.5 * (returns - values).pow(2).mean()
        if self.flat is not None:
            self.flat.zero_grad()
        else:
            self.optimizer.zero_grad(set_to_none=False)
        (value_loss * self.value_loss_coef + action_loss - entropy * self.entropy_coef).backward()
        if self.sync_enabled:
            if self.flat is not None:
                self.flat.all_reduce_mean()       # gradients are born contiguous: one NCCL call, no copies
            else:
                self.grad_sync()
        nn.utils.clip_grad_norm_(self.actor_critic.parameters(), self.max_grad_norm)
        self.optimizer.step()
        sums += torch.stack([value_loss.detach(), action_loss.detach(), entropy.detach()])

    def update(self, storage):
        adv = storage.returns[:-1] - storage.value_preds[:-1]
        mean, std = global_mean_std(adv)
        adv = (adv - mean) / (std + 1e-5)
        if self.use_graph and adv.is_cuda and (not dist_ready() or self.graph_with_nccl):
            return self._update_graphed(storage, adv)
        sums = torch.zeros(3, device=adv.device)
        n_updates = 0
        for _ in range(self.ppo_epoch):
            for obs, actions, old_values, returns, _masks, old_logp, adv_b in \
                    storage.batch_generator(adv, self.mini_batch_size):
                self._step(obs, actions, old_values, returns, old_logp, adv_b, sums)
                n_updates += 1
        v, a, e = (sums / max(n_updates, 1)).tolist()      # one host sync per update, not per mini-batch
        return v, a, e

    # ---- the same loop with the mini-batch step replayed as ONE CUDA graph (single process) -----------------
    def _update_graphed(self, storage, adv):
        """~100 small launches per mini-batch step (two 3-layer MLPs forward and backward, losses, gradient clip,
        Adam) cost more CPU time than GPU time; captured once, a step is one graph launch.  The batch is
        selected by an index tensor that is overwritten before every replay (random mini-batches without
        replacement, last partial batch dropped: storage.py:57-71)."""
        ns, mb, dev = storage.num_samples, self.mini_batch_size, adv.device
        key = (id(storage), ns, mb)
        if self._graph_key != key:
            self._graph, self._graph_key, self._warm = None, key, 0
            self._idx = torch.zeros(mb, dtype=torch.long, device=dev)
            self._adv = torch.zeros(ns, 1, device=dev)
            self._sums = torch.zeros(3, device=dev)
            self._flat = (storage.obs[:-1].reshape(ns, *storage.obs.shape[2:]), storage.actions.reshape(ns, -1),
                          storage.value_preds[:-1].reshape(ns, 1), storage.returns[:-1].reshape(ns, 1),
                          storage.action_log_probs.reshape(ns, 1))
        self._adv.copy_(adv.reshape(ns, 1))
        self._sums.zero_()

        def body():
            o, a, v, r, lp = (t.index_select(0, self._idx) for t in self._flat)
            self._step(o, a, v, r, lp, self._adv.index_select(0, self._idx), self._sums)

        n_updates = 0
        for _ in range(self.ppo_epoch):
            perm = torch.randperm(ns, device=dev)
            for start in range(0, ns - mb + 1, mb):
                self._idx.copy_(perm[start:start + mb])
                if self._graph is not None:
                    self._graph.replay()
                elif self._warm < 3:                       # eager warm-up steps (they are real updates) on a side stream
                    s = torch.cuda.Stream()
                    s.wait_stream(torch.cuda.current_stream())
                    with torch.cuda.stream(s):
                        body()
                    torch.cuda.current_stream().wait_stream(s)
                    self._warm += 1
                else:
                    try:
                        g = torch.cuda.CUDAGraph()
                        with torch.cuda.graph(g):
                            body()
                        self._graph = g
                        g.replay()
                    except Exception as e:                 # pragma: no cover - capture is best effort
                        print(f"[ppo] CUDA graph capture of the update failed ({e}); continuing eagerly", flush=True)
                        self.use_graph = False
                        body()
                n_updates += 1
        v, a, e = (self._sums / max(n_updates, 1)).tolist()
        return v, a, e
