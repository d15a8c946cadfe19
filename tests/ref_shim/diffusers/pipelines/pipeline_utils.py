"""The slice of DiffusionPipeline the reference pipelines use: register_modules, _execution_device, progress_bar,
maybe_free_model_hooks."""
import torch


class _Bar:
    def __init__(self, total):
        self.total, self.n = total, 0

    def update(self, k=1):
        self.n += k

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


class DiffusionPipeline:
    def register_modules(self, **modules):
        for name, m in modules.items():
            setattr(self, name, m)

    @property
    def _execution_device(self):
        for name in ("transformer", "controlnet", "vae"):
            m = getattr(self, name, None)
            if isinstance(m, torch.nn.Module):
                for p in m.parameters():
                    return p.device
        return torch.device("cpu")

    def progress_bar(self, iterable=None, total=None):
        return _Bar(total)

    def maybe_free_model_hooks(self):
        pass
