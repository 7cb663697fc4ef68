"""Import shim: exposes the package directory `micro-regime-propogation_b200/` (hyphenated, not a
valid Python identifier) as the importable package `mrp_b200`."""
import os as _os

__path__ = [_os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "micro-regime-propogation_b200")]
__package__ = __name__
__file__ = _os.path.join(__path__[0], "__init__.py")
with open(__file__) as _f:
    exec(compile(_f.read(), __file__, "exec"))
