"""Import shim: the package directory is `omni-swarm_b200/` (not a valid Python identifier), so
`import omniswarm_b200` resolves its submodules from that directory."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "omni-swarm_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
