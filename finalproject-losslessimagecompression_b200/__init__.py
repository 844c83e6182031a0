"""B200-native (sm_100a) entropy-coding hot path of lym01803/FinalProject-LosslessImageCompression.

Import as `flic_b200` (see flic_b200/__init__.py).  Modules:
  _lib       ctypes binding of libflic_b200.so (include/flic_b200.h); fails loudly if absent
  rans       drop-in `encode` / `decode` (reference: rans/rans.pyx) + tensor stream API
  coder      `Encode` / `Decode` wrappers (reference: coder.py:18-38)
  roundlib, couplelib, priorlib, distlib, invertible, extenddim, nnblock, nnlayer, flows
             host-side mirror of the reference modules that drive the path, with the elementwise
             hot spots routed to the CUDA kernels and `compress` / `decompress` filled in
  container  wire format of a compressed batch
  sharding   image-range sharding across ranks
"""
__version__ = "0.1.0"
