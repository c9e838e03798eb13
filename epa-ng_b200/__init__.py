"""epa-ng_b200: B200-native implementation of EPA-ng's per-query placement hot path.

The product is the C-ABI shared library `libepa_b200.so` (csrc/, declared in include/epa_b200.h)
plus the C++ host program built from csrc/host/. This Python package only holds the ctypes
binding of that ABI (capi.py), the host-side session wrapper (session.py) and the synthetic
data generator used by bench.py and the tests (synth.py).

The directory name contains a hyphen (it mirrors the reference's name), so it is imported
through `__graft_entry__.load_package()`, which registers it as `epa_ng_b200`.
"""
