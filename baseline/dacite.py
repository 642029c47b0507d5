"""Stand-in for the third-party ``dacite`` package (absent from this image and its wheelhouse).

The reference imports ``dacite`` at module scope in ``deepfm/config.py`` but only calls ``dacite.from_dict`` inside
``load_config``.  This minimal ``from_dict`` covers what that call needs (nested dataclasses, lists, scalars) so the
UNMODIFIED reference in ``baseline/_ref`` can be imported and timed by ``bench.py --impl reference``."""

import dataclasses
import typing


def from_dict(data_class, data, config=None):
    hints = typing.get_type_hints(data_class)
    kwargs = {}
    for f in dataclasses.fields(data_class):
        if f.name not in data:
            continue
        v, t = data[f.name], hints.get(f.name)
        if dataclasses.is_dataclass(t) and isinstance(v, dict):
            v = from_dict(t, v)
        kwargs[f.name] = v
    return data_class(**kwargs)
