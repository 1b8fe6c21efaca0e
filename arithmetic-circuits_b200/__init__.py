"""arithmetic-circuits_b200 -- B200-native R1CS / QAP hot path behind the reference's QAP API.

The directory name follows the project name and is not a valid Python identifier; import it as
`arithmetic_circuits_b200` (the alias package at the repository root extends its __path__ here)."""
from .qap import *  # noqa: F401,F403
from . import qap, _lib, json_io  # noqa: F401
