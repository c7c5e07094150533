"""CPU oracle for the similar-face-filtering hot path.  TEST INFRASTRUCTURE ONLY.

This module is the checker, never the product: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it.  The shipped package
(``face_detection_and_recognition_b200``) never imports anything from
``oracle/`` and has no CPU fallback.

It is a NumPy fp32 restatement of the arithmetic the reference does for this
path.  All ``path:line`` citations are relative to the reference checkout
(SamSamhuns/face_detection_and_recognition):

* ``similar_face_filtering/filter_faces_using_reference.py:85-99``  mean vector + max-distance threshold
* ``similar_face_filtering/filter_faces_using_reference.py:186-189`` per-row Euclid keep test (inclusive ``<=``)
* ``face_detection_and_extraction/face_extraction/extract_and_label_faces_from_dataset.py:101-116``
  per-pair cosine distance (``:106``) / Euclid distance (``:104``), threshold rule (``:110``), first-match scan
* ``face_detection_and_extraction/modules/mobile_facenet/mobile_facenet.py:30-33`` ``l2_norm``
* ``face_detection_and_extraction/face_extraction/extract_and_clean_imdb_wiki_faces.py:146`` NumPy L2-normalise

The arithmetic itself lives in third-party NumPy (``np.mean``, ``np.linalg.norm``,
``np.inner``; the reference pins numpy 1.26.4 in ``poetry.lock``), so the
restatement calls the same NumPy entry points at the same call-site shapes.

Pinning status (see tests/golden/README.md and DESIGN.md §3):
  * mean-vector / threshold / keep decision: pinned against the reference's own
    ``get_ref_mean_vec_and_thres_from_imgs`` and ``main()`` executed in the
    authoring container (TensorFlow I/O shimmed, arithmetic untouched) ->
    ``tests/golden/ref_main_*.npz``.
  * ``l2_norm``: pinned against the reference's importable torch function ->
    ``tests/golden/l2norm_*.npz``.
  * cosine rule: pinned against ``Net.check_if_face_exists`` executed through an
    unbound call -> ``tests/golden/label_scan_*.npz``.
  * gallery max / first-argmax over references: NOT in the reference (north-star
    requirement); defined here via ``np.argmax`` first-occurrence.  "parity
    unpinned" for that reduction only.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import numpy as np

METRIC_COSINE = 0
METRIC_EUCLID = 1


# --------------------------------------------------------------------------- #
# literal restatements (same expressions, same shapes, Python loops)
# --------------------------------------------------------------------------- #

def l2_norm(x: np.ndarray, axis: int = 1) -> np.ndarray:
    """mobile_facenet.py:30-33 -- ``norm = torch.norm(x, 2, axis, True); x / norm`` (no epsilon)."""
    x = np.asarray(x, dtype=np.float32)
    norm = np.sqrt(np.sum(x * x, axis=axis, keepdims=True, dtype=np.float32)).astype(np.float32)
    return (x / norm).astype(np.float32)


def l2_normalise_vec(face_feat: np.ndarray) -> np.ndarray:
    """extract_and_clean_imdb_wiki_faces.py:146 -- ``face_feat / np.linalg.norm(face_feat)``."""
    face_feat = np.asarray(face_feat, dtype=np.float32)
    return face_feat / np.linalg.norm(face_feat)


def ref_mean_vec_and_thres(ref_feat: np.ndarray, max_ref_img_count: int = 32) -> Tuple[np.ndarray, np.float32]:
    """filter_faces_using_reference.py:79-99.

    ``ref_feat`` is what the reference builds at :80-85: one ``model.predict``
    output of shape (1, D) per reference image, stacked to (R, 1, D), where
    R = min(max_ref_img_count, #images).  Returns (mean (1, D) f32, thres).
    """
    ref_feat = np.asarray(ref_feat, dtype=np.float32)
    if ref_feat.ndim == 2:
        ref_feat = ref_feat[:, None, :]
    ref_feat = ref_feat[:max_ref_img_count]
    ref_num = ref_feat.shape[0]
    ref_mean_vec = np.mean(ref_feat, axis=0)                      # :86
    max_dist_from_mean = 0
    for i in range(ref_num):                                      # :90-92
        max_dist_from_mean = max(max_dist_from_mean,
                                 np.linalg.norm(ref_mean_vec - ref_feat[i]))
    return ref_mean_vec, max_dist_from_mean


def euclid_keep_literal(output_batch: np.ndarray, ref_mean_vec: np.ndarray, thres) -> np.ndarray:
    """filter_faces_using_reference.py:186-189 -- per-row ``np.linalg.norm(out - mu) <= thres``."""
    keep = np.zeros(len(output_batch), dtype=np.uint8)
    for i, out in enumerate(output_batch):
        if np.linalg.norm(out - ref_mean_vec) <= thres:
            keep[i] = 1
    return keep


def cosine_dist_literal(feat: np.ndarray, new_feat: np.ndarray) -> np.float32:
    """extract_and_label_faces_from_dataset.py:106."""
    return 1 - (np.inner(feat, new_feat) / (np.linalg.norm(feat) * np.linalg.norm(new_feat)))


def euclid_dist_literal(feat: np.ndarray, new_feat: np.ndarray) -> np.float32:
    """extract_and_label_faces_from_dataset.py:104."""
    return np.linalg.norm(feat - new_feat)


def first_match_scan(gallery: List[np.ndarray], new_feat: np.ndarray, metric: int,
                     normal_thres: float = 1.0, harsh_thres: float = 0.72,
                     ious: Optional[List[float]] = None) -> Tuple[bool, int]:
    """extract_and_label_faces_from_dataset.py:101-116 -- linear scan, FIRST hit wins.

    Rule (:110): ``(dist < normal_thres and iou > 0.1) or dist < harsh_thres``.
    Returns (found, gallery position or -1).  The in-place gallery update
    (:113-114) is the caller's job.
    """
    for i, feat in enumerate(gallery):
        dist = euclid_dist_literal(feat, new_feat) if metric == METRIC_EUCLID else cosine_dist_literal(feat, new_feat)
        iou = 0.0 if ious is None else ious[i]
        if (dist < normal_thres and iou > 0.1) or dist < harsh_thres:
            return True, i
    return False, -1


# --------------------------------------------------------------------------- #
# vectorised fp32 restatement (what the CUDA path is compared with at scale)
# --------------------------------------------------------------------------- #

def row_norms(x: np.ndarray) -> np.ndarray:
    x = np.asarray(x, dtype=np.float32)
    return np.sqrt(np.einsum("ij,ij->i", x, x, dtype=np.float32)).astype(np.float32)


def filter_cosine(ref: np.ndarray, cand: np.ndarray, thr: float, block: int = 8192,
                  dtype=np.float32) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Gallery cosine filter: for every candidate, max / first-argmax over references of
    ``(r.c)/(|r||c|)`` (extract_and_label...:106 rearranged as similarity = 1 - dist) and
    ``keep = best >= thr``.  Returns (keep u8[M], best_idx i32[M], best_sim f32[M]).

    ``dtype=np.float64`` gives the fp64 shadow used to compute the tolerance band.
    """
    ref = np.asarray(ref, dtype=dtype)
    cand = np.asarray(cand, dtype=dtype)
    rn = ref / np.sqrt(np.einsum("ij,ij->i", ref, ref, dtype=dtype))[:, None].astype(dtype)
    m = cand.shape[0]
    best = np.empty(m, dtype=dtype)
    idx = np.empty(m, dtype=np.int32)
    for s in range(0, m, block):
        c = cand[s:s + block]
        cn = c / np.sqrt(np.einsum("ij,ij->i", c, c, dtype=dtype))[:, None].astype(dtype)
        sim = rn @ cn.T                                            # [N, b]
        idx[s:s + block] = np.argmax(sim, axis=0)                  # first occurrence
        best[s:s + block] = np.max(sim, axis=0)
    keep = (best >= dtype(thr)).astype(np.uint8)
    return keep, idx, best.astype(dtype)


def filter_cosine_torch(ref: np.ndarray, cand: np.ndarray, thr: float, block: int = 16384):
    """Same restatement as ``filter_cosine`` with torch-CPU doing the sgemm / max / argmax (multi-threaded MKL): the
    fastest CPU form of the reference's arithmetic, used as the CPU baseline in bench.py.  torch.max over dim 0 returns
    the first maximal index like np.argmax."""
    import torch
    with torch.no_grad():
        r = torch.from_numpy(np.ascontiguousarray(ref, dtype=np.float32))
        rn = r / torch.linalg.vector_norm(r, dim=1, keepdim=True)
        m = cand.shape[0]
        best = torch.empty(m, dtype=torch.float32)
        idx = torch.empty(m, dtype=torch.int64)
        for s in range(0, m, block):
            c = torch.from_numpy(np.ascontiguousarray(cand[s:s + block], dtype=np.float32))
            cn = c / torch.linalg.vector_norm(c, dim=1, keepdim=True)
            sim = rn @ cn.T
            b, i = torch.max(sim, dim=0)
            best[s:s + block], idx[s:s + block] = b, i
        keep = (best >= thr).to(torch.uint8)
    return keep.numpy(), idx.numpy().astype(np.int32), best.numpy()


def filter_euclid(ref: np.ndarray, cand: np.ndarray, thr: float, block: int = 4096,
                  dtype=np.float32) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Gallery Euclid filter: ``best = min_i |c - r_i|`` (filter_faces...:189 when ref = {mean}),
    ``keep = best <= thr``.  Direct differences (no norm expansion) so it matches the
    reference's cancellation behaviour.  Returns (keep, best_idx, best_dist)."""
    ref = np.asarray(ref, dtype=dtype)
    cand = np.asarray(cand, dtype=dtype)
    n = ref.shape[0]
    m = cand.shape[0]
    best = np.empty(m, dtype=dtype)
    idx = np.empty(m, dtype=np.int32)
    step = max(1, block // max(1, n // 64))
    for s in range(0, m, step):
        c = cand[s:s + step]
        diff = c[None, :, :] - ref[:, None, :]                     # [N, b, D]
        d = np.sqrt(np.einsum("nbd,nbd->nb", diff, diff, dtype=dtype))
        idx[s:s + step] = np.argmin(d, axis=0)
        best[s:s + step] = np.min(d, axis=0)
    keep = (best <= dtype(thr)).astype(np.uint8)
    return keep, idx, best.astype(dtype)


def tolerance_band(best64: np.ndarray, thr: float, tol: float = 1e-3) -> np.ndarray:
    """Rows whose (fp64-shadow) best similarity lies within ``tol`` of the threshold: the
    only rows whose keep bit / index may legitimately differ (BASELINE.md parity bar)."""
    return np.nonzero(np.abs(np.asarray(best64, dtype=np.float64) - thr) <= tol)[0]


# --------------------------------------------------------------------------- #
# synthetic workloads
# --------------------------------------------------------------------------- #

def make_synthetic(n_ref: int, n_cand: int, dim: int, seed: int = 42, planted: float = 0.5,
                   n_adversarial: int = 0, n_dup_refs: int = 0, thr: float = 0.5,
                   unit_norm: bool = True) -> Tuple[np.ndarray, np.ndarray]:
    """Unit-norm references; candidates: a ``planted`` fraction are noisy copies of a random
    reference with cos in [0.55, 0.95], the rest independent noise.  ``n_adversarial`` rows are
    placed within +-2e-3 of ``thr``; ``n_dup_refs`` references are exact duplicates of earlier
    ones (first-argmax rule).  Seed 42 = the reference's seed (filter_faces...:24)."""
    rng = np.random.default_rng(seed)
    ref = rng.standard_normal((n_ref, dim), dtype=np.float32)
    ref /= np.linalg.norm(ref, axis=1, keepdims=True)
    if n_dup_refs and n_ref > 1:
        n_dup_refs = min(n_dup_refs, n_ref // 2)
        src = rng.integers(0, n_ref // 2, n_dup_refs)
        dst = rng.choice(np.arange(n_ref // 2, n_ref), n_dup_refs, replace=False)
        ref[dst] = ref[src]
    cand = rng.standard_normal((n_cand, dim), dtype=np.float32)
    cand /= np.linalg.norm(cand, axis=1, keepdims=True)
    n_pl = int(planted * n_cand)
    if n_pl:
        rows = rng.choice(n_cand, n_pl, replace=False)
        k = rng.integers(0, n_ref, n_pl)
        cos = rng.uniform(0.55, 0.95, n_pl).astype(np.float32)
        if n_adversarial:
            na = min(n_adversarial, n_pl)
            cos[:na] = thr + rng.uniform(-2e-3, 2e-3, na).astype(np.float32)
        g = cand[rows]
        r = ref[k]
        g = g - np.sum(g * r, axis=1, keepdims=True) * r           # orthogonal part
        g /= np.linalg.norm(g, axis=1, keepdims=True)
        cand[rows] = cos[:, None] * r + np.sqrt(1 - cos[:, None] ** 2) * g
    if not unit_norm:
        cand *= rng.uniform(0.5, 12.0, (n_cand, 1)).astype(np.float32)
        ref *= rng.uniform(0.5, 12.0, (n_ref, 1)).astype(np.float32)
    return ref.astype(np.float32), cand.astype(np.float32)
