"""Python-file configs with attribute access: the subset of ``mmcv.Config`` that tools/train.py:70-72 and
tools/test.py:82-84 of the reference rely on (``Config.fromfile``, ``cfg.model.backbone.depth``,
``cfg.merge_from_dict({'a.b': 1})``, ``cfg.get``)."""
import os
import runpy


class ConfigDict(dict):
    """dict whose items are also attributes; nested dicts are converted recursively."""

    def __init__(self, *args, **kwargs):
        super().__init__()
        for k, v in dict(*args, **kwargs).items():
            self[k] = v

    @staticmethod
    def _wrap(v):
        if isinstance(v, ConfigDict):
            return v
        if isinstance(v, dict):
            return ConfigDict(v)
        if isinstance(v, list):
            return [ConfigDict._wrap(x) for x in v]
        if isinstance(v, tuple):
            return tuple(ConfigDict._wrap(x) for x in v)
        return v

    def __setitem__(self, k, v):
        super().__setitem__(k, ConfigDict._wrap(v))

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError:
            raise AttributeError(f"'{type(self).__name__}' object has no attribute '{name}'")

    def __setattr__(self, name, value):
        self[name] = value

    def __delattr__(self, name):
        try:
            del self[name]
        except KeyError:
            raise AttributeError(name)

    def copy(self):
        return ConfigDict(self)

    def to_dict(self):
        def un(v):
            if isinstance(v, dict):
                return {k: un(x) for k, x in v.items()}
            if isinstance(v, list):
                return [un(x) for x in v]
            if isinstance(v, tuple):
                return tuple(un(x) for x in v)
            return v

        return un(self)


class Config:
    """``cfg = Config.fromfile('configs/r50_nc_sgd_cos_100e_r5_1xNx2_k400.py')``."""

    def __init__(self, cfg_dict=None, filename=None, text=''):
        if cfg_dict is None:
            cfg_dict = {}
        if not isinstance(cfg_dict, dict):
            raise TypeError(f'cfg_dict must be a dict, but got {type(cfg_dict)}')
        object.__setattr__(self, '_cfg_dict', ConfigDict(cfg_dict))
        object.__setattr__(self, '_filename', filename)
        object.__setattr__(self, '_text', text)

    @staticmethod
    def fromfile(filename):
        filename = os.path.abspath(os.path.expanduser(filename))
        if not os.path.isfile(filename):
            raise FileNotFoundError(f'file "{filename}" does not exist')
        if not filename.endswith('.py'):
            raise IOError('Only py type configs are supported')
        ns = runpy.run_path(filename)
        cfg = {k: v for k, v in ns.items() if not k.startswith('__') and not callable(v)
               and not isinstance(v, type(os))}
        with open(filename) as fh:
            text = fh.read()
        return Config(cfg, filename=filename, text=text)

    @property
    def filename(self):
        return self._filename

    @property
    def text(self):
        return self._text

    def merge_from_dict(self, options):
        for full_key, v in options.items():
            d = self._cfg_dict
            keys = full_key.split('.')
            for sub in keys[:-1]:
                if sub not in d or not isinstance(d[sub], dict):
                    d[sub] = ConfigDict()
                d = d[sub]
            d[keys[-1]] = v

    def get(self, key, default=None):
        return self._cfg_dict.get(key, default)

    def __getattr__(self, name):
        return getattr(self._cfg_dict, name)

    def __getitem__(self, name):
        return self._cfg_dict[name]

    def __setattr__(self, name, value):
        self._cfg_dict[name] = value

    def __setitem__(self, name, value):
        self._cfg_dict[name] = value

    def __contains__(self, name):
        return name in self._cfg_dict

    def __iter__(self):
        return iter(self._cfg_dict)

    def __len__(self):
        return len(self._cfg_dict)

    def __repr__(self):
        return f'Config (path: {self._filename}): {dict.__repr__(self._cfg_dict)}'
