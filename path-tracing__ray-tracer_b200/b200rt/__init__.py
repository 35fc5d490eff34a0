"""b200rt — B200-native path-tracing core behind the reference renderer API.

Python here only builds/packs scenes and calls the C-ABI library
(``libb200rt.so``, hand-written sm_100a CUDA); torch tensors are device buffers.
"""
__version__ = "0.1.0"
