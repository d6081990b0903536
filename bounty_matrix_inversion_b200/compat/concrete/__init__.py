"""Import shim: put this directory's parent on sys.path and the reference's own
`from concrete import fhe` resolves to the B200 engine."""
import os
import sys

_root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
if _root not in sys.path:
    sys.path.insert(0, _root)

from bounty_matrix_inversion_b200 import fhe  # noqa: E402,F401

sys.modules[__name__ + ".fhe"] = fhe
