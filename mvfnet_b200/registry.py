"""Class registries + config-dict construction.

API contract kept from the reference (codes/utils/registry.py:7-81): `Registry(name)`, `.name`,
`.module_dict`, `.get(key)`, the `@REG.register_module` class decorator, and
`build_from_cfg(cfg, registry, default_args)` with its error types (KeyError for unknown / duplicate
names, TypeError for non-class registrations or a bad `type` value).
"""
import inspect


class Registry:
    def __init__(self, name):
        self._name, self._module_dict = name, {}

    name = property(lambda self: self._name)
    module_dict = property(lambda self: self._module_dict)

    def __repr__(self):
        return '%s(name=%s, items=%s)' % (type(self).__name__, self._name, list(self._module_dict))

    def __contains__(self, key):
        return key in self._module_dict

    def __len__(self):
        return len(self._module_dict)

    def get(self, key):
        return self._module_dict.get(key)

    def register_module(self, cls):
        """Decorator: file `cls` under its class name."""
        if not inspect.isclass(cls):
            raise TypeError('module must be a class, but got %s' % type(cls))
        key = cls.__name__
        if key in self._module_dict:
            raise KeyError('%s is already registered in %s' % (key, self._name))
        self._module_dict[key] = cls
        return cls


def _resolve(kind, registry):
    if isinstance(kind, str):
        found = registry.get(kind)
        if found is None:
            raise KeyError('%s is not in the %s registry' % (kind, registry.name))
        return found
    if inspect.isclass(kind):
        return kind
    raise TypeError('type must be a str or valid type, but got %s' % type(kind))


def build_from_cfg(cfg, registry, default_args=None):
    """Instantiate cfg['type'] (a registered name or a class) with the other keys as kwargs;
    `default_args` fill in keys the config does not set.  The config dict is not modified."""
    if not (isinstance(cfg, dict) and 'type' in cfg):
        raise AssertionError('cfg must be a dict with a "type" key')
    if not (default_args is None or isinstance(default_args, dict)):
        raise AssertionError('default_args must be a dict or None')
    kwargs = {k: v for k, v in cfg.items() if k != 'type'}
    for k, v in (default_args or {}).items():
        kwargs.setdefault(k, v)
    return _resolve(cfg['type'], registry)(**kwargs)
