"""Importable alias of the product package, whose directory name (`self-play-ai_b200/`, mandated by the
repo layout) is not a valid Python identifier.  `import selfplay_b200` executes
`self-play-ai_b200/__init__.py` with this package's `__path__` pointing at that directory."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "self-play-ai_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _f
