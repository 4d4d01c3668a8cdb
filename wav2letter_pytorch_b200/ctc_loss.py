"""Drop-in for ``torch.nn.CTCLoss`` as the reference constructs it (base_asr_models.py:23):
``CTCLoss(blank=0, reduction='mean', zero_infinity=True)`` called as
``criterion(out.transpose(0,1), targets, output_lengths, target_lengths)`` (base_asr_models.py:81,90)."""
import torch
import torch.nn as nn

from . import functional as F


class _CTCLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, log_probs_tnc, targets, input_lengths, target_lengths, blank, zero_infinity, reduction):
        x = log_probs_tnc.transpose(0, 1)                    # [N,T,C] view; the kernel takes arbitrary N/T strides
        need_grad = log_probs_tnc.requires_grad
        loss, nll, grad = F.ctc_loss_raw(x, targets, input_lengths, target_lengths, blank=blank, zero_infinity=zero_infinity,
                                         reduction_mean=(reduction == "mean"), need_grad=need_grad)
        ctx.reduction = reduction
        if need_grad:
            ctx.save_for_backward(grad)
        if reduction == "none":
            return nll
        return loss[0]

    @staticmethod
    def backward(ctx, grad_out):
        (grad,) = ctx.saved_tensors                          # [N,T,C]; 'mean' scaling already applied by the kernel
        if ctx.reduction == "none":
            g = grad * grad_out.view(-1, 1, 1)
        else:
            g = grad * grad_out
        return g.transpose(0, 1), None, None, None, None, None, None


class CTCLoss(nn.Module):
    def __init__(self, blank=0, reduction="mean", zero_infinity=False):
        super().__init__()
        if reduction not in ("mean", "sum", "none"):
            raise ValueError("%s is not a valid value for reduction" % reduction)
        self.blank, self.reduction, self.zero_infinity = blank, reduction, zero_infinity

    def forward(self, log_probs, targets, input_lengths, target_lengths):
        """log_probs [T,N,C] (any strides), targets [N,S] int, lengths [N] int -- same as torch.nn.CTCLoss."""
        if not torch.is_tensor(input_lengths):
            input_lengths = torch.as_tensor(input_lengths, dtype=torch.int32)
        if not torch.is_tensor(target_lengths):
            target_lengths = torch.as_tensor(target_lengths, dtype=torch.int32)
        return _CTCLossFn.apply(log_probs, targets, input_lengths, target_lengths, self.blank, self.zero_infinity, self.reduction)
