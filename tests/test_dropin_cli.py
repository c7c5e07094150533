"""End to end through the drop-in entry point (same flags and output tree as the reference's
similar_face_filtering/filter_faces_using_reference.py :103-199): synthetic jpgs -> embeddings (stub model with the
reference's ``predict`` contract) -> GPU filter -> clean/unclean tree + summary line, checked against the oracle."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import oracle

pytestmark = pytest.mark.gpu


class StubModel:
    """``predict(batch, verbose=0) -> [b, 128] float32`` (reference :84,184): fixed random projection of the pooled image."""
    inputs, outputs = "[b,160,160,3]", "[b,128]"

    def __init__(self):
        rng = np.random.default_rng(42)
        self.w = (rng.standard_normal((16 * 16 * 3, 128)) * 0.035).astype(np.float32)

    def predict(self, batch, verbose=0):
        b = np.asarray(batch, dtype=np.float32)
        pooled = b.reshape(b.shape[0], 16, 10, 16, 10, 3).mean(axis=(2, 4)).reshape(b.shape[0], -1)
        return (pooled @ self.w).astype(np.float32)


def _make_dataset(root, rng, classes=("alice", "bob"), n_ref=12, n_cand=40):
    from PIL import Image
    for cls_i, cls in enumerate(classes):
        base = rng.integers(0, 255, (24, 20, 3))
        for kind, n in (("ref", n_ref), ("unf", n_cand)):
            d = os.path.join(root, kind, cls)
            os.makedirs(d)
            for i in range(n):
                # same-identity images = the class pattern plus noise; a third of the candidates are off-class
                pattern = base if (kind == "ref" or i % 3) else rng.integers(0, 255, (24, 20, 3))
                img = np.clip(pattern + rng.normal(0, 25, pattern.shape), 0, 255).astype(np.uint8)
                img = np.kron(img, np.ones((8, 8, 1), dtype=np.uint8))            # 192 x 160
                Image.fromarray(img).save(os.path.join(d, f"{i}.jpg"), quality=95)
    return os.path.join(root, "unf"), os.path.join(root, "ref")


def _embed(model, paths):
    from face_detection_and_recognition_b200.filter_faces_using_reference import read_and_preprocess_img
    return model.predict(torch.stack([read_and_preprocess_img(p) for p in paths]).numpy())


def test_main_reproduces_reference_semantics(tmp_path, capsys, ffr_lib, cuda_dev):
    from face_detection_and_recognition_b200.filter_faces_using_reference import main
    rng = np.random.default_rng(1)
    ud, rd = _make_dataset(str(tmp_path), rng)
    td = str(tmp_path / "out")
    model = StubModel()
    main(["--ud", ud, "--rd", rd, "--td", td, "-b", "16", "-r", "8"], model=model)
    out = capsys.readouterr().out
    for cls in ("alice", "bob"):
        refs = sorted(glob.glob(os.path.join(rd, cls, "*.jpg")))[:8]
        cands = sorted(glob.glob(os.path.join(ud, cls, "*.jpg")))
        mu, thres = oracle.ref_mean_vec_and_thres(_embed(model, refs)[:, None, :], 8)
        emb = _embed(model, cands)
        keep = oracle.euclid_keep_literal(emb, mu, thres)
        d = np.linalg.norm(emb - mu, axis=1)
        for p, k, dist in zip(cands, keep, d):
            name = os.path.basename(p)
            in_clean = os.path.exists(os.path.join(td, "clean", cls, name))
            in_unclean = os.path.exists(os.path.join(td, "unclean", cls, name))
            assert in_clean != in_unclean, f"{cls}/{name} must be copied to exactly one side"
            if abs(dist - thres) > 1e-4:
                assert in_clean == bool(k), f"{cls}/{name}: dist {dist} thres {thres}"
        assert 0 < keep.sum() < len(keep)
        pos = len(os.listdir(os.path.join(td, "clean", cls)))
        assert f"Similar images percentage={pos / len(cands):2.2f}%, positive={pos}, total={len(cands)}" in out


def test_gallery_mode(tmp_path, ffr_lib, cuda_dev):
    from face_detection_and_recognition_b200.filter_faces_using_reference import main
    rng = np.random.default_rng(2)
    ud, rd = _make_dataset(str(tmp_path), rng, classes=("carol",), n_ref=12, n_cand=30)
    td = str(tmp_path / "out")
    model = StubModel()
    main(["--ud", ud, "--rd", rd, "--td", td, "--gallery", "--threshold", "0.6", "-r", "12"], model=model)
    refs = sorted(glob.glob(os.path.join(rd, "carol", "*.jpg")))
    cands = sorted(glob.glob(os.path.join(ud, "carol", "*.jpg")))
    ko, io, so = oracle.filter_cosine(_embed(model, refs), _embed(model, cands), 0.6)
    for p, k, s in zip(cands, ko, so):
        if abs(s - 0.6) > 1e-3:
            assert os.path.exists(os.path.join(td, "clean" if k else "unclean", "carol", os.path.basename(p)))
