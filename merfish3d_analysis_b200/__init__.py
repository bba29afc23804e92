"""Import alias: the product package lives in ``merfish3d-analysis_b200/`` (a directory
name that is not a Python identifier); this stub makes it importable as
``merfish3d_analysis_b200`` by pointing ``__path__`` at it."""

import pathlib as _pathlib

_real = _pathlib.Path(__file__).resolve().parent.parent / "merfish3d-analysis_b200"
__path__ = [str(_real)]
exec(compile((_real / "__init__.py").read_text(), str(_real / "__init__.py"), "exec"))
