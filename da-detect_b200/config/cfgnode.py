"""Minimal attribute-dict config node with the subset of the yacs behaviour the
reference relies on (yacs itself is a dependency of the reference,
maskrcnn_benchmark/config/defaults.py:4, and is not a dependency here).

Semantics kept so the reference YAMLs under configs/da_faster_rcnn/ load unmodified:
  * merge_from_file / merge_from_list / merge_from_other_cfg, clone, freeze/defrost;
  * string values are passed through ``ast.literal_eval`` (so ``(600,)`` in YAML
    becomes a tuple), and list<->tuple are coerced to the default's type;
  * a key that is absent from the defaults raises KeyError (typo protection);
  * a type mismatch raises ValueError.
"""
import ast
import copy

import yaml


class CfgNode(dict):
    _IMMUTABLE = "__immutable__"

    def __init__(self, init=None):
        super().__init__()
        self.__dict__[CfgNode._IMMUTABLE] = False
        for k, v in (init or {}).items():
            self[k] = CfgNode(v) if isinstance(v, dict) and not isinstance(v, CfgNode) else v

    # attribute access -------------------------------------------------------
    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError:
            raise AttributeError(name)

    def __setattr__(self, name, value):
        if self.is_frozen():
            raise AttributeError("attempted to set {} on a frozen CfgNode".format(name))
        self[name] = value

    # freezing ---------------------------------------------------------------
    def is_frozen(self):
        return self.__dict__[CfgNode._IMMUTABLE]

    def _set_frozen(self, flag):
        self.__dict__[CfgNode._IMMUTABLE] = flag
        for v in self.values():
            if isinstance(v, CfgNode):
                v._set_frozen(flag)

    def freeze(self):
        self._set_frozen(True)

    def defrost(self):
        self._set_frozen(False)

    def clone(self):
        return copy.deepcopy(self)

    def __deepcopy__(self, memo):
        out = CfgNode()
        for k, v in self.items():
            out[k] = copy.deepcopy(v, memo)
        out.__dict__[CfgNode._IMMUTABLE] = self.is_frozen()
        return out

    # merging ----------------------------------------------------------------
    @staticmethod
    def _decode(value):
        if isinstance(value, dict):
            return CfgNode(value)
        if not isinstance(value, str):
            return value
        try:
            return ast.literal_eval(value)
        except (ValueError, SyntaxError):
            return value

    @staticmethod
    def _coerce(new, old, key):
        if old is None or type(new) is type(old):
            return new
        for a, b in ((list, tuple), (tuple, list)):
            if isinstance(new, a) and isinstance(old, b):
                return b(new)
        if isinstance(old, float) and isinstance(new, int) and not isinstance(new, bool):
            return float(new)
        if isinstance(old, int) and not isinstance(old, bool) and isinstance(new, float):
            return new  # the reference YAMLs put floats into int defaults (e.g. BIAS_LR_FACTOR)
        raise ValueError("type mismatch for config key {}: {} vs default {}".format(
            key, type(new).__name__, type(old).__name__))

    def _merge(self, other, path):
        for k, v in other.items():
            full = ".".join(path + [k])
            if k not in self:
                raise KeyError("non-existent config key: {}".format(full))
            v = self._decode(copy.deepcopy(v))
            if isinstance(self[k], CfgNode):
                if not isinstance(v, dict):
                    raise ValueError("config key {} must be a mapping".format(full))
                self[k]._merge(v, path + [k])
            else:
                self[k] = self._coerce(v, self[k], full)

    def merge_from_other_cfg(self, other):
        if self.is_frozen():
            raise AttributeError("cannot merge into a frozen CfgNode")
        self._merge(other, [])

    def merge_from_file(self, filename):
        with open(filename, "r") as f:
            loaded = yaml.safe_load(f) or {}
        self.merge_from_other_cfg(loaded)

    def merge_from_list(self, opts):
        if len(opts) % 2:
            raise ValueError("override list must have an even length: {}".format(opts))
        if self.is_frozen():
            raise AttributeError("cannot merge into a frozen CfgNode")
        for full, v in zip(opts[0::2], opts[1::2]):
            node, keys = self, full.split(".")
            for k in keys[:-1]:
                if k not in node:
                    raise KeyError("non-existent config key: {}".format(full))
                node = node[k]
            if keys[-1] not in node:
                raise KeyError("non-existent config key: {}".format(full))
            node[keys[-1]] = self._coerce(self._decode(v), node[keys[-1]], full)

    def dump(self):
        def plain(n):
            return {k: plain(v) if isinstance(v, CfgNode) else (list(v) if isinstance(v, tuple) else v)
                    for k, v in n.items()}
        return yaml.safe_dump(plain(self))
