import logging as _logging
from collections import OrderedDict
from dataclasses import fields, is_dataclass

WEIGHTS_NAME = "diffusion_pytorch_model.bin"
SAFETENSORS_WEIGHTS_NAME = "diffusion_pytorch_model.safetensors"


class BaseOutput(OrderedDict):
    """dataclass + dict hybrid (diffusers utils/outputs.py): fields are reachable as attributes, keys and by index."""

    def __post_init__(self):
        if is_dataclass(self):
            for f in fields(self):
                v = getattr(self, f.name)
                if v is not None:
                    self[f.name] = v

    def __getitem__(self, k):
        if isinstance(k, str):
            return dict(self.items())[k]
        return self.to_tuple()[k]

    def to_tuple(self):
        return tuple(self[k] for k in self.keys())


class _Logging:
    @staticmethod
    def get_logger(name=None):
        return _logging.getLogger(name)


logging = _Logging()
