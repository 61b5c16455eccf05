"""The part of the `fastargs` 1.2 API that the reference's driver uses (main.py:57,106-189,1170-1177), for boxes
where the pip package is absent: `Section(...).params(...)`, `Param`, `get_current_config()`,
`@param('section.key')`, `OneOf` / `And`, and the config object's `augment_argparse`, `collect_argparse_args`,
`validate`, `summary`, `get`, `__getitem__` and `entries`.  So the `--section.key value` / `--section.key=value`
command line of run_train.sh (and a `--config-file` YAML) reaches the train step under the reference's key names.
Host-side configuration only; nothing here touches the GPU path.
"""
import argparse
import functools
import sys
from types import SimpleNamespace


class _Checker:
    def check(self, value):
        raise NotImplementedError

    def help(self):
        return ''


class _Type(_Checker):
    def __init__(self, t):
        self.t = t

    def check(self, value):
        if self.t is bool and isinstance(value, str):
            if value.lower() in ('1', 'true', 'yes'):
                return True
            if value.lower() in ('0', 'false', 'no'):
                return False
            raise ValueError(f'{value!r} is not a boolean')
        return self.t(value)

    def help(self):
        return self.t.__name__


class OneOf(_Checker):
    def __init__(self, possible_values):
        self.possible_values = list(possible_values)

    def check(self, value):
        if value not in self.possible_values:
            raise ValueError(f'{value!r} not in {self.possible_values}')
        return value

    def help(self):
        return 'one of [' + ', '.join(map(str, self.possible_values)) + ']'


class And(_Checker):
    def __init__(self, *checkers):
        self.checkers = [_as_checker(c) for c in checkers]

    def check(self, value):
        for c in self.checkers:
            value = c.check(value)
        return value

    def help(self):
        return ' and '.join(c.help() for c in self.checkers)


def _as_checker(c):
    return c if isinstance(c, _Checker) else _Type(c)


class Param:
    def __init__(self, checker, desc='', default=None, required=False):
        self.checker = _as_checker(checker)
        self.desc = desc
        self.default = default
        self.required = required


class Config:
    def __init__(self):
        self.entries = {}            # ('section', 'key') -> Param
        self.sections = {}           # ('section',) -> description
        self.content = {}            # ('section', 'key') -> raw value given by the user
        self.conditions = {}

    # ------------------------------------------------------------------ declaration
    def declare(self, path, parameter):
        self.entries[tuple(path)] = parameter

    # ------------------------------------------------------------------ sources
    def collect(self, mapping, prefix=()):
        for k, v in mapping.items():
            path = prefix + tuple(str(k).split('.'))
            if isinstance(v, dict):
                self.collect(v, path)
            else:
                self.content[path] = v
        return self

    def collect_config_file(self, fname):
        import yaml
        with open(fname) as fh:
            self.collect(yaml.safe_load(fh) or {})
        return self

    def augment_argparse(self, parser):
        parser.add_argument('--config-file', '-C', action='append', default=[], help='YAML file(s) with section: {key: value}')
        for path, p in self.entries.items():
            name = '.'.join(path)
            extra = ' (required)' if p.required else f' (default: {p.default})'
            parser.add_argument(f'--{name}', dest=name, default=None, metavar='',
                                help=f'{p.desc}; {p.checker.help()}{extra}')

    def collect_argparse_args(self, parser, argv=None):
        args = parser.parse_args(argv)
        for fname in getattr(args, 'config_file', None) or []:
            self.collect_config_file(fname)
        for path in self.entries:
            v = getattr(args, '.'.join(path), None)
            if v is not None:
                self.content[path] = v               # the command line wins over files
        return self

    # ------------------------------------------------------------------ access
    def _path(self, key):
        return tuple(key.split('.')) if isinstance(key, str) else tuple(key)

    def __getitem__(self, key):
        path = self._path(key)
        p = self.entries[path]
        if path in self.content:
            return p.checker.check(self.content[path])
        if p.required:
            raise KeyError(f'missing required parameter {".".join(path)}')
        return p.default

    def get(self):
        root = SimpleNamespace()
        for path in self.entries:
            node = root
            for part in path[:-1]:
                if not hasattr(node, part):
                    setattr(node, part, SimpleNamespace())
                node = getattr(node, part)
            try:
                setattr(node, path[-1], self[path])
            except KeyError:
                setattr(node, path[-1], None)
        return root

    def validate(self, mode='stderr'):
        errors = {}
        for path, p in self.entries.items():
            try:
                self[path]
            except (KeyError, ValueError, TypeError) as e:
                errors['.'.join(path)] = str(e).strip("'\"")
        for path in self.content:
            if path not in self.entries:
                errors['.'.join(path)] = 'unknown parameter'
        if errors and mode == 'stderr':
            for k, v in errors.items():
                print(f'config error: {k}: {v}', file=sys.stderr)
            sys.exit(1)
        if errors and mode == 'errordict':
            return errors
        if errors:
            raise ValueError(errors)
        return {}

    def summary(self, stream=None):
        stream = stream or sys.stdout
        rows = [('.'.join(path), self[path]) for path in self.entries]
        w = max(len(r[0]) for r in rows) if rows else 0
        print('Parameter'.ljust(w) + ' | Value', file=stream)
        for k, v in rows:
            print(k.ljust(w) + f' | {v}', file=stream)


_CURRENT = [Config()]


def get_current_config():
    return _CURRENT[0]


def set_current_config(config):
    _CURRENT[0] = config


class Section:
    def __init__(self, ns, desc=''):
        self.ns = tuple(ns.split('.')) if isinstance(ns, str) else tuple(ns)
        self.desc = desc
        get_current_config().sections[self.ns] = desc

    def params(self, **kwargs):
        for k, p in kwargs.items():
            get_current_config().declare(self.ns + (k,), p)
        return self

    def enable_if(self, condition):
        get_current_config().conditions[self.ns] = condition
        return self


def param(parameter, alias=None):
    """Decorator: fills the keyword argument named after the last component of `section.key` (or `alias`) from the
    current config unless the caller passed it (fastargs.decorators.param)."""
    path = tuple(parameter.split('.'))
    name = alias or path[-1]

    def wrap(fn):
        @functools.wraps(fn)
        def inner(*args, **kwargs):
            if name not in kwargs:
                kwargs[name] = get_current_config()[path]
            return fn(*args, **kwargs)
        return inner
    return wrap


def make_config(argv=None, quiet=False, description='Fast imagenet training'):
    """main.py:1162-1169."""
    config = get_current_config()
    parser = argparse.ArgumentParser(description=description)
    config.augment_argparse(parser)
    config.collect_argparse_args(parser, argv)
    config.validate(mode='stderr')
    if not quiet:
        config.summary()
    return config
