"""Test infrastructure: shims that let the UNMODIFIED reference (/root/reference, build
container only) be imported and run as the ground truth for tests/golden/make_golden.py.

Nothing here is reference code; see SURVEY.md Appendix A for why each shim is needed.
"""
import os
import sys
import types

_REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
# the reference tree (build container) or its unmodified git-ignored copy that travels to the GPU box
# (baseline/_ref, made by oracle/install_reference.py)
_CANDIDATES = ("/root/reference", os.path.join(_REPO, "baseline", "_ref"))


def reference_root():
    for p in _CANDIDATES:
        if os.path.isdir(os.path.join(p, "beetroots")):
            return p
    return None


REFERENCE_ROOT = reference_root() or _CANDIDATES[0]


def reference_available() -> bool:
    return reference_root() is not None


def install():
    """Make ``import beetroots`` (reference) and ``import nnbma`` (stand-in) work."""
    import numpy as np

    if not reference_available():
        raise RuntimeError("no reference tree: neither /root/reference nor baseline/_ref (oracle/install_reference.py)")
    if not hasattr(np, "infty"):
        np.infty = np.inf                                   # my_sampler.py:277,372,379
    sys.modules.setdefault("h5py", types.ModuleType("h5py"))  # abstract_saver.py:5
    here = os.path.dirname(os.path.abspath(__file__))
    for p in (here, reference_root()):
        if p not in sys.path:
            sys.path.insert(0, p)
    from beetroots.sampler.abstract_sampler import Sampler   # noqa: E402

    def _state(self):                                        # numpy>=2: abstract_sampler.py:63-72 breaks
        from sys import byteorder as bo
        st = self.rng.bit_generator.state["state"]
        return (np.array(bytearray(st["state"].to_bytes(32, bo))),
                np.array(bytearray(st["inc"].to_bytes(32, bo))))

    Sampler.get_rng_state = _state
