"""B200-native similar-face-filtering hot path (drop-in for the reference's
similar_face_filtering/filter_faces_using_reference.py).  See DESIGN.md.

Layout:
  csrc/                              hand-written sm_100a CUDA kernels + the C ABI (include/ffr.h)
  _build.py / _lib.py                in-tree nvcc build, ctypes binding (fails loudly without the .so)
  ops.py                             torch-facing wrappers (device memory + streams only)
  filter_faces_using_reference.py    host-side mirror of the reference entry point
  embeddings_io.py                   on-disk embedding formats of the reference's extraction scripts
  sharding.py                        candidate-axis sharding + NCCL gather across the GPUs of one box
"""
__version__ = "0.1.0"
