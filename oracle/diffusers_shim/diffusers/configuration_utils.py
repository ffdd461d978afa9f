"""ConfigMixin / register_to_config restated (diffusers 0.29.2 configuration_utils.py): the decorator records
the constructor's keyword arguments (with defaults) into `self.config`, an attribute-style frozen dict."""
import functools
import inspect


class FrozenDict(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e


class ConfigMixin:
    config_name = "config.json"

    def register_to_config(self, **kwargs):
        cfg = dict(getattr(self, "_internal_dict", {}))
        cfg.update(kwargs)
        self._internal_dict = FrozenDict(cfg)

    @property
    def config(self):
        return self._internal_dict

    @classmethod
    def from_config(cls, config, **kwargs):
        sig = inspect.signature(cls.__init__).parameters
        args = {k: v for k, v in dict(config).items() if k in sig and not k.startswith("_")}
        args.update(kwargs)
        return cls(**args)


def register_to_config(init):
    @functools.wraps(init)
    def inner(self, *args, **kwargs):
        sig = inspect.signature(init)
        names = [n for n in sig.parameters if n != "self"]
        cfg = {n: p.default for n, p in sig.parameters.items() if n != "self" and p.default is not inspect._empty}
        cfg.update(dict(zip(names, args)))
        cfg.update({k: v for k, v in kwargs.items() if not k.startswith("_")})
        init(self, *args, **kwargs)
        self.register_to_config(**cfg)

    return inner
