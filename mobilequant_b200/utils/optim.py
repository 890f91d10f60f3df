"""Optimiser glue (mirror of mobilellm/utils/optim.py:5-46).

The reference wraps torch.cuda.amp.GradScaler: scale(loss).backward -> unscale_ -> grad-norm -> step (skipped when a
gradient is inf/nan) -> update.  Calibration here always runs in fp32 (--deactive_amp), where the loss scale is a
power of two and therefore numerically neutral, so only the observable behaviour is kept: backward, global L2
grad-norm, skip-step-on-non-finite."""
import torch


@torch.no_grad()
def ampscaler_get_grad_norm(parameters, norm_type=2.0):
    if isinstance(parameters, torch.Tensor):
        parameters = [parameters]
    grads = [p.grad.detach() for p in parameters if p.grad is not None]
    if len(grads) == 0:
        return torch.tensor(0.)
    if norm_type == float("inf"):
        return max(g.abs().max() for g in grads)
    return torch.norm(torch.stack([torch.norm(g, norm_type) for g in grads]), norm_type)


class NativeScalerWithGradNormCount:
    state_dict_key = "amp_scaler"

    def __init__(self):
        self.skipped_steps = 0

    def step_only(self, optimizer, clip_grad=None, parameters=None):
        parameters = list(parameters) if parameters is not None else None
        if clip_grad is not None:
            assert parameters is not None
            norm = torch.nn.utils.clip_grad_norm_(parameters, clip_grad)
        else:
            norm = ampscaler_get_grad_norm(parameters)
        if torch.isfinite(norm):                  # GradScaler.step semantics: skip the update on inf/nan
            optimizer.step()
        else:
            self.skipped_steps += 1
        return norm

    def __call__(self, loss, optimizer, clip_grad=None, parameters=None, create_graph=False, update_grad=True,
                 retain_graph=False):
        loss.backward(create_graph=create_graph, retain_graph=retain_graph)
        if not update_grad:
            return None
        return self.step_only(optimizer, clip_grad, parameters)

    def state_dict(self):
        return {"skipped_steps": self.skipped_steps}

    def load_state_dict(self, state_dict):
        self.skipped_steps = state_dict.get("skipped_steps", 0)


class FlatAdamW:
    """torch.optim.AdamW as the reference constructs it (alg:513, 716-722) + the scaler's grad-norm / skip-on-non-finite
    (optim.py:28-41), executed by ONE fused kernel pair (mq_adamw_step) over a flat parameter buffer.

    Every learnable is re-pointed to a view of `flat` (16-byte aligned slots, grouped by learning-rate group) and its
    .grad to a view of `flat_grad`, so autograd accumulates straight into the buffer the kernel reads; the learning rates
    and the step counter live on the device, which lets a whole training step be replayed as a CUDA graph.
    param_groups mirrors the torch attribute for callers that set `param_groups[i]["lr"]`."""

    ALIGN = 4      # floats

    def __init__(self, groups, weight_decay=0.0, betas=(0.9, 0.999), eps=1e-8, device=None):
        import torch
        from .. import kernels as K
        self.K = K
        self.param_groups = [{"params": list(g["params"]), "lr": float(g["lr"])} for g in groups]
        self.betas, self.eps, self.weight_decay = betas, eps, weight_decay
        params = [p for g in self.param_groups for p in g["params"]]
        if not params:
            raise ValueError("FlatAdamW: no parameters")
        self.device = device if device is not None else params[0].device
        off, self.slots, self.seg_end = 0, [], []
        for g in self.param_groups:
            for p in g["params"]:
                if p.dtype != torch.float32:
                    raise TypeError("FlatAdamW: fp32 learnables only")
                self.slots.append((p, off, p.numel()))
                off += (p.numel() + self.ALIGN - 1) // self.ALIGN * self.ALIGN
            self.seg_end.append(off)
        self.n = off
        z = lambda: torch.zeros(self.n, dtype=torch.float32, device=self.device)
        self.flat, self.flat_grad, self.exp_avg, self.exp_avg_sq = z(), z(), z(), z()
        with torch.no_grad():
            for p, o, n in self.slots:
                self.flat[o:o + n].copy_(p.detach().reshape(-1))
                p.data = self.flat[o:o + n].view(p.shape)
                p.grad = self.flat_grad[o:o + n].view(p.shape)
        self.lr_host = torch.tensor([g["lr"] for g in self.param_groups], dtype=torch.float32).pin_memory() \
            if self.device.type == "cuda" else torch.tensor([g["lr"] for g in self.param_groups], dtype=torch.float32)
        self.lr_dev = self.lr_host.to(self.device)
        self.state = torch.zeros(8, dtype=torch.float32, device=self.device)      # step, norm, found_inf, bc1, sqrt(bc2), skipped

    def set_lr(self, idx, value):
        self.param_groups[idx]["lr"] = float(value)

    def sync_lr(self):
        """Host learning rates -> device (one tiny copy per step, outside any captured graph)."""
        for i, g in enumerate(self.param_groups):
            self.lr_host[i] = g["lr"]
        self.lr_dev.copy_(self.lr_host, non_blocking=True)

    def zero_grad(self, set_to_none=False):
        self.flat_grad.zero_()                  # the views stay attached: autograd accumulates into the flat buffer
        for p, o, n in self.slots:
            if p.grad is None or p.grad.data_ptr() != self.flat_grad.data_ptr() + 4 * o:
                p.grad = self.flat_grad[o:o + n].view(p.shape)

    def allreduce_grads(self, world):
        """Data-parallel exchange: SUM of the flat gradient buffer over NCCL, then the batch mean (alg:459)."""
        import torch.distributed as dist
        if world > 1:
            dist.all_reduce(self.flat_grad, op=dist.ReduceOp.SUM)
            self.flat_grad.div_(world)

    def step(self):
        """Returns the global gradient norm (0-d device tensor; the update is skipped on the device when it is not finite)."""
        self.K.adamw_step(self.flat, self.flat_grad, self.exp_avg, self.exp_avg_sq, self.seg_end, self.lr_dev, self.state,
                          self.betas, self.eps, self.weight_decay)
        return self.state[1]

    @property
    def skipped_steps(self):
        return int(self.state[5].item())
