#!/usr/bin/env python
"""Embedding-producer throughput: images per second from JPEG files to device-resident embeddings,
the reference's per-image host pipeline + NumPy hand-over (filter_faces_using_reference.py:60-68,168-184; PIL standing in
for tf.io.decode_jpeg) against the device-side pipeline (thread-pool file reads, batched nvJPEG decode, GPU resize /
standardise, embeddings stay on the device).  Synthetic JPEG set with the bundled faces' size range; the model is a small
conv net with the MobileFaceNet adapter's ``embed`` / ``predict`` contract (the reference's weights are not in the repo).

    python tools/bench_embed.py [--images 2048] [--batch 128]        # prints one JSON line
"""
import argparse
import json
import os
import sys
import tempfile
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from face_detection_and_recognition_b200.filter_faces_using_reference import _embed_paths_device  # noqa: E402


class ConvEmbed(torch.nn.Module):
    def __init__(self, dev):
        super().__init__()
        self.net = torch.nn.Sequential(
            torch.nn.Conv2d(3, 32, 3, stride=2, padding=1), torch.nn.ReLU(), torch.nn.Conv2d(32, 64, 3, stride=2, padding=1),
            torch.nn.ReLU(), torch.nn.Conv2d(64, 128, 3, stride=2, padding=1), torch.nn.ReLU(),
            torch.nn.AdaptiveAvgPool2d(1), torch.nn.Flatten(), torch.nn.Linear(128, 512))
        self.dev = torch.device(dev)
        self.to(self.dev).eval()

    @torch.no_grad()
    def embed(self, batch):
        return torch.nn.functional.normalize(self.net(batch.permute(0, 3, 1, 2)))

    def predict(self, batch, verbose=0):
        return self.embed(torch.as_tensor(np.asarray(batch), dtype=torch.float32, device=self.dev)).cpu().numpy()


class HostOnly:
    """the same model seen through the reference's contract only (no ``embed``): per-image PIL decode, NumPy batches"""
    def __init__(self, m):
        self.predict = m.predict


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--images", type=int, default=2048)
    ap.add_argument("--batch", type=int, default=128)
    args = ap.parse_args()
    from PIL import Image
    rng = np.random.default_rng(0)
    dev = torch.device("cuda:0")
    with tempfile.TemporaryDirectory() as d:
        paths = []
        for i in range(args.images):
            h, w = int(rng.integers(135, 721)), int(rng.integers(100, 515))          # bundled faces: ~100x135 ... 514x720
            small = rng.integers(0, 255, (h // 8 + 1, w // 8 + 1, 3)).astype(np.uint8)
            img = np.kron(small, np.ones((8, 8, 1), dtype=np.uint8))[:h, :w]
            p = os.path.join(d, f"{i:05d}.jpg")
            Image.fromarray(img).save(p, quality=90)
            paths.append(p)
        model = ConvEmbed(dev)
        out = {}
        for name, m, n in (("device_pipeline", model, args.images), ("host_pipeline", HostOnly(model), min(args.images, 512))):
            _embed_paths_device(m, paths[:args.batch], args.batch, dev)               # warm-up (nvJPEG handles, cudnn)
            torch.cuda.synchronize()
            st = {}
            t0 = time.perf_counter()
            emb = _embed_paths_device(m, paths[:n], args.batch, dev, stats=st)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            out[name] = {"images": n, "seconds": dt, "imgs_per_s": n / dt, "decoder": st["decoder"], "embedding_shape": list(emb.shape)}
        out["speedup"] = out["device_pipeline"]["imgs_per_s"] / out["host_pipeline"]["imgs_per_s"]
        out["batch"] = args.batch
        out["host_threads"] = len(os.sched_getaffinity(0))
        print(json.dumps(out))


if __name__ == "__main__":
    main()
