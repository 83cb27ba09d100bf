"""desta_b200 — B200-native waveform -> LLM-embedding front-end for DeSTA2.5-Audio (import package).

Host side (Python, mirrors the reference's interfaces for this path) over ``libdesta_b200.so`` (hand-written sm_100a
CUDA behind a C ABI, see include/desta_b200.h).  There is no CPU fallback: every op raises if the library is missing.

The modules of this package live in ``../desta2.5-audio_b200/`` - the directory name this repository's layout
prescribes is not a legal Python identifier, so this importable package points its ``__path__`` there; ``import
desta_b200.ops`` etc. load the files of that directory (and ``libdesta_b200.so`` next to them).
"""
import os as _os

__path__ = [_os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "desta2.5-audio_b200")]
__version__ = "0.2.0"


def install(model, **kw):
    """Swap the B200 path into a live reference ``DeSTA25AudioModel`` (see desta_b200.patch.install)."""
    from .patch import install as _install
    return _install(model, **kw)
