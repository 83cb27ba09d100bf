"""desta_b200 — B200-native waveform -> LLM-embedding front-end for DeSTA2.5-Audio.

Host side (Python, mirrors the reference's interfaces for this path) over ``libdesta_b200.so`` (hand-written sm_100a
CUDA behind a C ABI, see include/desta_b200.h).  There is no CPU fallback: every op raises if the library is missing.
"""
__version__ = "0.1.0"


def install(model, **kw):
    from .patch import install as _install
    return _install(model, **kw)
