"""Loader for the reference's plain-python config files (stands in for mmcv.Config.fromfile, which
train_recognizer.py:52 uses): executes the file and exposes its top-level names as attributes, with
nested dicts reachable both as attributes and as items."""
import os


class ConfigDict(dict):
    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError:
            raise AttributeError(name)

    def __setattr__(self, name, value):
        self[name] = value


def _wrap(v):
    if isinstance(v, dict):
        return ConfigDict((k, _wrap(x)) for k, x in v.items())
    if isinstance(v, list):
        return [_wrap(x) for x in v]
    if isinstance(v, tuple):
        return tuple(_wrap(x) for x in v)
    return v


class Config(ConfigDict):
    @staticmethod
    def fromfile(filename):
        filename = os.path.abspath(os.path.expanduser(filename))
        if not os.path.isfile(filename):
            raise FileNotFoundError(filename)
        if not filename.endswith('.py'):
            raise IOError('Only py type are supported now!')
        scope = {'__file__': filename}
        import numpy
        legacy = not hasattr(numpy, 'Inf')          # the reference configs spell np.Inf (NumPy < 2)
        if legacy:
            numpy.Inf = numpy.inf
        try:
            with open(filename, 'r') as f:
                exec(compile(f.read(), filename, 'exec'), scope)
        finally:
            if legacy:
                del numpy.Inf
        cfg = Config((k, _wrap(v)) for k, v in scope.items() if not k.startswith('__') and not callable(v)
                     and not isinstance(v, type(os)))
        cfg['filename'] = filename
        return cfg
