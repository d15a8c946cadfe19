import torch


class ModelMixin(torch.nn.Module):
    """nn.Module with diffusers' ``dtype`` / ``device`` properties."""
    _supports_gradient_checkpointing = False

    @property
    def dtype(self):
        for p in self.parameters():
            return p.dtype
        return torch.float32

    @property
    def device(self):
        for p in self.parameters():
            return p.device
        return torch.device("cpu")
