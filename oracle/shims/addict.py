"""ORACLE shim: the slice of `addict.Dict` the reference's options.py uses."""
import copy


class Dict(dict):
    def __init__(self, *args, **kwargs):
        super().__init__()
        for a in args:
            if not a:
                continue
            for k, v in (a.items() if isinstance(a, dict) else a):
                self[k] = self._hook(v)
        for k, v in kwargs.items():
            self[k] = self._hook(v)

    @classmethod
    def _hook(cls, v):
        if isinstance(v, dict) and not isinstance(v, cls):
            return cls(v)
        if isinstance(v, (list, tuple)):
            return type(v)(cls._hook(e) for e in v)
        return v

    def __getattr__(self, name):
        return self.__getitem__(name)

    def __setattr__(self, name, value):
        self[name] = value

    def __setitem__(self, name, value):
        super().__setitem__(name, value)

    def __missing__(self, name):
        v = type(self)()
        self[name] = v
        return v

    def __delattr__(self, name):
        del self[name]

    def to_dict(self):
        out = {}
        for k, v in self.items():
            if isinstance(v, Dict):
                out[k] = v.to_dict()
            elif isinstance(v, (list, tuple)):
                out[k] = type(v)(e.to_dict() if isinstance(e, Dict) else e for e in v)
            else:
                out[k] = v
        return out

    def copy(self):
        return copy.copy(self)

    def __deepcopy__(self, memo):
        other = type(self)()
        memo[id(self)] = other
        for k, v in self.items():
            dict.__setitem__(other, copy.deepcopy(k, memo), copy.deepcopy(v, memo))
        return other

    def __getstate__(self):
        return dict(self)

    def __setstate__(self, state):
        self.update(state)
