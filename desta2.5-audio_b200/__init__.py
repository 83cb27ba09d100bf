"""Implementation directory of the import package ``desta_b200`` (see ../desta_b200/__init__.py, whose ``__path__`` points
here): host-side mirrors of the reference's interfaces for the waveform -> LLM-embedding path, the ctypes binding, and
``csrc/`` with the sm_100a kernels behind the C ABI of include/desta_b200.h."""
