import logging as _pylogging
from dataclasses import fields

import torch
from packaging import version

USE_PEFT_BACKEND = False


class _Logging:
    @staticmethod
    def get_logger(name):
        return _pylogging.getLogger(name)


logging = _Logging()


class BaseOutput:
    """Dataclass-style output, indexable like a tuple or by field name (diffusers.utils.BaseOutput)."""

    def to_tuple(self):
        return tuple(getattr(self, f.name) for f in fields(self) if getattr(self, f.name) is not None)

    def __getitem__(self, k):
        if isinstance(k, str):
            return getattr(self, k)
        return self.to_tuple()[k]


def is_torch_version(op, v):
    cur = version.parse(version.parse(torch.__version__).base_version)
    ops = {">": cur > version.parse(v), ">=": cur >= version.parse(v), "<": cur < version.parse(v),
           "<=": cur <= version.parse(v), "==": cur == version.parse(v)}
    return ops[op]


def is_torch_xla_available():
    return False


def replace_example_docstring(example):
    def deco(fn):
        return fn
    return deco


def scale_lora_layers(model, weight):
    pass


def unscale_lora_layers(model, weight=None):
    pass
