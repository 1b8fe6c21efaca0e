"""Import alias for the package directory `arithmetic-circuits_b200/` (hyphen: not importable by name).
All code lives there; this module only redirects the package path."""
import os as _os

__path__.insert(0, _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                                 "arithmetic-circuits_b200"))
from .qap import *  # noqa: E402,F401,F403
from . import qap, _lib, json_io  # noqa: E402,F401
