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
