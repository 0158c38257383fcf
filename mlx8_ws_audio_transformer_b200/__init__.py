"""Importable name of the package that lives in ``mlx8-ws-audio-transformer_b200/``.

The product directory carries the repository's name, which is not a valid Python identifier;
this shim makes ``import mlx8_ws_audio_transformer_b200`` resolve to it without symlinks.
"""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      "mlx8-ws-audio-transformer_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _os, _f
