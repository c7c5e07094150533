"""Minimal stand-in for the ``tensorflow`` names that
``similar_face_filtering/filter_faces_using_reference.py`` touches, so that the UNMODIFIED
reference script can be imported and its ``get_ref_mean_vec_and_thres_from_imgs`` / ``main``
executed in a container without TensorFlow.  TEST INFRASTRUCTURE ONLY (used by
``oracle/gen_golden.py``); never imported by the product.

Only I/O is replaced (jpeg decode, resize, standardisation, dataset batching, model loading).
The arithmetic being pinned -- ``np.mean`` (:86), ``np.linalg.norm`` (:92, :189), the ``<=``
test (:189) -- is NumPy inside the reference's own code and is not touched.
"""
from __future__ import annotations

import sys
import types

import numpy as np
from PIL import Image


class _Dataset:
    def __init__(self, items, fn=None, batch=None):
        self._items, self._fn, self._batch = list(items), fn, batch

    @staticmethod
    def from_tensor_slices(items):
        return _Dataset(items)

    def map(self, fn):
        return _Dataset(self._items, fn, self._batch)

    def batch(self, b):
        return _Dataset(self._items, self._fn, b)

    def __len__(self):
        n = len(self._items)
        return n if self._batch is None else (n + self._batch - 1) // self._batch

    def __iter__(self):
        fn = self._fn or (lambda x: x)
        if self._batch is None:
            for it in self._items:
                yield fn(it)
            return
        for s in range(0, len(self._items), self._batch):
            yield np.stack([fn(it) for it in self._items[s:s + self._batch]])


def _read_file(path):
    with open(path, "rb") as f:
        return f.read()


def _decode_jpeg(data, channels=3, dct_method=""):
    import io
    return np.asarray(Image.open(io.BytesIO(data)).convert("RGB"))


def _convert_image_dtype(img, dtype):
    return np.asarray(img, dtype=np.float32) / np.float32(255.0)


def _resize(img, size):
    import torch
    import torch.nn.functional as F
    t = torch.from_numpy(np.ascontiguousarray(img)).permute(2, 0, 1)[None]
    t = F.interpolate(t, size=tuple(size), mode="bilinear", align_corners=False, antialias=False)
    return t[0].permute(1, 2, 0).contiguous().numpy()


def _per_image_standardization(img):
    img = np.asarray(img, dtype=np.float32)
    n = img.size
    adj = max(float(img.std()), 1.0 / np.sqrt(n))
    return ((img - img.mean()) / adj).astype(np.float32)


def install(model_factory):
    """Register a fake ``tensorflow`` in ``sys.modules``; ``tf.keras.models.load_model`` returns
    ``model_factory(path)`` -- any object with ``predict(batch, verbose=0) -> np.ndarray`` plus
    ``inputs`` / ``outputs`` attributes (the reference prints them, :135-136)."""
    tf = types.ModuleType("tensorflow")
    tf.float32 = np.float32
    tf.Tensor = np.ndarray
    tf.random = types.SimpleNamespace(set_seed=lambda s: None)
    tf.data = types.SimpleNamespace(Dataset=_Dataset)
    tf.io = types.SimpleNamespace(read_file=_read_file)
    tf.image = types.SimpleNamespace(decode_jpeg=_decode_jpeg, convert_image_dtype=_convert_image_dtype,
                                     resize=_resize, per_image_standardization=_per_image_standardization)
    tf.keras = types.SimpleNamespace(Model=object,
                                     models=types.SimpleNamespace(load_model=lambda p, compile=False: model_factory(p)))
    sys.modules["tensorflow"] = tf
    return tf
