"""Import alias: the package lives in ``gcn-vae_b200/`` (a hyphen is not importable).

``import gcn_vae_b200`` resolves submodules from that directory and executes its __init__.
"""
import os

_REAL = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gcn-vae_b200")
__path__ = [_REAL]
with open(os.path.join(_REAL, "__init__.py")) as _f:
    exec(compile(_f.read(), os.path.join(_REAL, "__init__.py"), "exec"))
