"""pyxrd_b200 -- B200-native (sm_100a CUDA) implementation of PyXRD's matrix-method
hot path, drop-in behind the function names of ``pyxrd.calculations``.

``pyxrd_b200.calculations``  mirror of the reference interface (device-backed)
``pyxrd_b200.packing``       DataObject graph -> packed arrays of include/pyxrd_b200.h
``pyxrd_b200.runtime``       ctypes binding of csrc/libpyxrd_b200.so (no CPU fallback)
``pyxrd_b200.batched``       whole-population entry points + multi-GPU sharding
``pyxrd_b200.install``       rebinding of ``pyxrd.calculations.*`` inside a PyXRD process
"""
__version__ = "0.1.0"
