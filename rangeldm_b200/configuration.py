"""Minimal mirror of diffusers' ConfigMixin / ModelMixin surface that the reference touches:
`load_config`, `from_config`, `.config.<key>`, `register_to_config`, `save_pretrained`,
`from_pretrained` (`ldm/inference.py:84-88,126-127`, `ldm/pipelines.py:139`,
`ldm/train_unconditional.py:153,171-173,675`)."""
import inspect
import json
import os


class FrozenDict(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        raise AttributeError("config is frozen; use register_to_config")


class ConfigMixin:
    config_name = "config.json"
    ignore_for_config = ()

    def register_to_config(self, **kw):
        cfg = dict(getattr(self, "_internal_dict", {}))
        cfg.update(kw)
        object.__setattr__(self, "_internal_dict", FrozenDict(cfg))

    @property
    def config(self):
        return self._internal_dict

    def _capture_init(self, local_vars):
        """Record the constructor arguments (call as `self._capture_init(locals())` first thing)."""
        sig = inspect.signature(self.__init__)
        cfg = {}
        for name in sig.parameters:
            if name in ("self", "kwargs") or name in self.ignore_for_config:
                continue
            v = local_vars[name]
            cfg[name] = list(v) if isinstance(v, tuple) else v
        cfg["_class_name"] = type(self).__name__
        self.register_to_config(**cfg)

    @classmethod
    def load_config(cls, path, subfolder=None, **kw):
        if isinstance(path, dict):
            return dict(path)
        path = os.fspath(path)
        if subfolder:
            path = os.path.join(path, subfolder)
        if os.path.isdir(path):
            path = os.path.join(path, cls.config_name)
        with open(path) as f:
            return json.load(f)

    @classmethod
    def from_config(cls, config, **kw):
        cfg = dict(config)
        cfg.update(kw)
        params = inspect.signature(cls.__init__).parameters
        accepted = {k: v for k, v in cfg.items() if k in params and k != "self"}
        return cls(**accepted)

    def save_config(self, save_directory):
        os.makedirs(save_directory, exist_ok=True)
        with open(os.path.join(save_directory, self.config_name), "w") as f:
            json.dump(dict(self.config), f, indent=2, sort_keys=True)


class ModelMixin(ConfigMixin):
    weights_name = "diffusion_pytorch_model.safetensors"

    @property
    def device(self):
        return next(self.parameters()).device

    @property
    def dtype(self):
        return next(self.parameters()).dtype

    def save_pretrained(self, save_directory, **kw):
        import safetensors.torch
        self.save_config(save_directory)
        sd = {k: v.detach().contiguous().cpu() for k, v in self.state_dict().items()}
        safetensors.torch.save_file(sd, os.path.join(save_directory, self.weights_name))

    @classmethod
    def from_pretrained(cls, path, subfolder=None, **kw):
        import safetensors.torch
        if subfolder:
            path = os.path.join(path, subfolder)
        model = cls.from_config(cls.load_config(path))
        model.load_state_dict(safetensors.torch.load_file(os.path.join(path, cls.weights_name)))
        return model
