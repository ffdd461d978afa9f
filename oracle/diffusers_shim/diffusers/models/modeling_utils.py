import torch


class ModelMixin(torch.nn.Module):
    """Only what the reference touches: .device and .dtype of the first parameter (modeling_utils.py)."""
    _supports_gradient_checkpointing = False

    @property
    def device(self):
        return next(self.parameters()).device

    @property
    def dtype(self):
        return next(self.parameters()).dtype
