"""YAML configuration with ``_base_`` inheritance -- API-compatible with the reference's
``src/utils/options.py`` (BaseConfig._file2dict_yaml :63-112, _merge_a_into_b :115-130, attribute access,
``opt.get(key, default)``), so ``config/crdr.yaml`` and ``scripts/compress.py`` work unchanged."""
import argparse
import copy
import os

import yaml

BASE_KEY = "_base_"
DELETE_KEY = "_delete_"
RESERVED_KEYS = ("filename", "text")


class ConfigDict(dict):
    """dict with attribute access; nested dicts are wrapped on the way in; missing keys raise."""

    def __init__(self, *args, **kwargs):
        super().__init__()
        for k, v in dict(*args, **kwargs).items():
            self[k] = v

    @classmethod
    def _wrap(cls, v):
        if isinstance(v, dict) and not isinstance(v, ConfigDict):
            return cls(v)
        if isinstance(v, (list, tuple)):
            return type(v)(cls._wrap(e) for e in v)
        return v

    def __setitem__(self, key, value):
        super().__setitem__(key, self._wrap(value))

    def __setattr__(self, key, value):
        self[key] = value

    def __getattr__(self, key):
        try:
            return self[key]
        except KeyError:
            raise AttributeError(f"'{type(self).__name__}' object has no attribute '{key}'") from None

    def __delattr__(self, key):
        del self[key]

    def to_dict(self):
        plain = lambda v: v.to_dict() if isinstance(v, ConfigDict) else (
            type(v)(plain(e) for e in v) if isinstance(v, (list, tuple)) else v)
        return {k: plain(v) for k, v in self.items()}

    def __deepcopy__(self, memo):
        out = type(self)()
        memo[id(self)] = out
        for k, v in self.items():
            dict.__setitem__(out, copy.deepcopy(k, memo), copy.deepcopy(v, memo))
        return out

    def __getstate__(self):
        return dict(self)

    def __setstate__(self, state):
        self.update(state)


class BaseConfig:
    @staticmethod
    def _file2dict_yaml(filename):
        filename = os.path.abspath(os.path.expanduser(filename))
        if not os.path.isfile(filename):
            raise FileNotFoundError(f'file "{filename}" does not exist')
        if os.path.splitext(filename)[1] != ".yaml":
            raise IOError("Only yaml type are supported now!")
        with open(filename, encoding="utf-8") as f:
            raw = f.read()
        cfg = yaml.safe_load(raw) or {}
        text = filename + "\n" + raw
        loaded = [filename]
        if BASE_KEY in cfg:
            bases = cfg.pop(BASE_KEY)
            bases = bases if isinstance(bases, list) else [bases]
            merged, texts = {}, []
            for rel in bases:
                sub, sub_text, sub_loaded = BaseConfig._file2dict_yaml(os.path.join(os.path.dirname(filename), rel))
                clash = merged.keys() & sub.keys()
                if clash:
                    raise KeyError(f"Duplicate key is not allowed among bases. Duplicate keys: {clash}")
                merged.update(sub)
                texts.append(sub_text)
                loaded.extend(sub_loaded)
            cfg = BaseConfig._merge_a_into_b(cfg, merged)
            text = "\n".join(texts + [text])
        return cfg, text, loaded

    @staticmethod
    def _merge_a_into_b(a, b):
        """Recursive override of b by a; a dict carrying ``_delete_: True`` replaces instead of merging."""
        out = b.copy()
        for k, v in a.items():
            if isinstance(v, dict) and k in out and not v.pop(DELETE_KEY, False):
                if not isinstance(out[k], dict):
                    raise TypeError(f"{k}={v} in child config cannot inherit from base because {k} is a dict in "
                                    f"the child config but is of type {type(out[k])} in base config. You may set "
                                    f"`{DELETE_KEY}=True` to ignore the base config")
                out[k] = BaseConfig._merge_a_into_b(v, out[k])
            else:
                out[k] = v
        return out

    def __init__(self, cfg_dict=None, cfg_text=None, filename=None):
        cfg_dict = {} if cfg_dict is None else cfg_dict
        if not isinstance(cfg_dict, dict):
            raise TypeError(f"cfg_dict must be a dict, but got {type(cfg_dict)}")
        for key in cfg_dict:
            if key in RESERVED_KEYS:
                raise KeyError(f"{key} is reserved for config file")
        object.__setattr__(self, "_cfg_dict", ConfigDict(cfg_dict))
        object.__setattr__(self, "_filename", filename)
        if not cfg_text and filename:
            with open(filename) as f:
                cfg_text = f.read()
        object.__setattr__(self, "_text", cfg_text or "")

    @classmethod
    def fromfile(cls, filename, **overrides):
        cfg, text, _ = cls._file2dict_yaml(filename)
        return cls(cls._merge_a_into_b(overrides, cfg), cfg_text=text, filename=filename)

    filename = property(lambda self: self._filename)
    text = property(lambda self: self._text)

    def __repr__(self):
        return f"Config (path: {self.filename}): {self._cfg_dict!r}"

    def __len__(self):
        return len(self._cfg_dict)

    def __getattr__(self, name):
        if name.startswith("_"):
            raise AttributeError(name)
        return getattr(self._cfg_dict, name)

    def __getitem__(self, name):
        return self._cfg_dict[name]

    def __setattr__(self, name, value):
        self._cfg_dict[name] = value

    def __setitem__(self, name, value):
        self._cfg_dict[name] = value

    def __iter__(self):
        return iter(self._cfg_dict)

    def __contains__(self, name):
        return name in self._cfg_dict

    def __getstate__(self):
        return (self._cfg_dict, self._filename, self._text)

    def __setstate__(self, state):
        for k, v in zip(("_cfg_dict", "_filename", "_text"), state):
            object.__setattr__(self, k, v)

    def __deepcopy__(self, memo):
        obj = type(self).__new__(type(self))
        obj.__setstate__((copy.deepcopy(self._cfg_dict, memo), self._filename, self._text))
        return obj

    def dump(self, filename):
        with open(filename, "w") as f:
            yaml.dump(self._cfg_dict.to_dict(), f)


class TestConfig(BaseConfig):
    """Inference-side config; ``scripts/compress.py`` subclasses this and overrides get_opt/arg_parse
    exactly as the reference script does (compress.py:23-45)."""

    @classmethod
    def get_opt(cls, config_dir=None, arg_dict=None):
        arg_dict = arg_dict or cls.arg_parse()
        filename = arg_dict["config_path"]
        cfg, text, loaded = cls._file2dict_yaml(filename)
        cfg["loaded_yamls"] = loaded
        arg_dict = cls._merge_a_into_b(arg_dict, cfg)
        arg_dict["exp"] = os.path.basename(filename).split(".")[0]
        arg_dict["is_train"] = False
        return cls(arg_dict, cfg_text=text, filename=filename)

    @staticmethod
    def arg_parse():
        p = argparse.ArgumentParser()
        p.add_argument("config_path", type=str)
        p.add_argument("-d", "--device", type=str, default="cuda:0")
        return {k: v for k, v in vars(p.parse_args()).items() if v is not None}
