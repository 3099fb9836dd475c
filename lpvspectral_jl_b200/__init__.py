"""Import alias: the product package lives in ``lpvspectral.jl_b200/`` (a directory name Python cannot import
directly because of the dot).  This shim points ``lpvspectral_jl_b200`` at that directory."""
import os as _os

__path__.insert(0, _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "..", "lpvspectral.jl_b200"))

from ._api import *  # noqa: F401,F403,E402
from ._api import __all__  # noqa: E402
