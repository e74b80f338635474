"""Test infrastructure: shims that let the UNMODIFIED reference (/root/reference, build
container only) be imported and run as the ground truth for tests/golden/make_golden.py.

Nothing here is reference code; see SURVEY.md Appendix A for why each shim is needed.
"""
import os
import sys
import types

REFERENCE_ROOT = "/root/reference"


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "beetroots"))


def install():
    """Make ``import beetroots`` (reference) and ``import nnbma`` (stand-in) work."""
    import numpy as np

    if not reference_available():
        raise RuntimeError("the reference tree is only present in the build container")
    if not hasattr(np, "infty"):
        np.infty = np.inf                                   # my_sampler.py:277,372,379
    sys.modules.setdefault("h5py", types.ModuleType("h5py"))  # abstract_saver.py:5
    here = os.path.dirname(os.path.abspath(__file__))
    for p in (here, REFERENCE_ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    from beetroots.sampler.abstract_sampler import Sampler   # noqa: E402

    def _state(self):                                        # numpy>=2: abstract_sampler.py:63-72 breaks
        from sys import byteorder as bo
        st = self.rng.bit_generator.state["state"]
        return (np.array(bytearray(st["state"].to_bytes(32, bo))),
                np.array(bytearray(st["inc"].to_bytes(32, bo))))

    Sampler.get_rng_state = _state
