#!/usr/bin/env python
"""Benchmark of the similar-face-filtering hot path: face pairs compared per second.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg3] [--impl reference]

One "step" = one pass of the hot path over one batch of synthetic embeddings that are already resident in HBM:
K1 (row L2-normalise + fp16 cast of the references; of the candidates too unless K2 does it in-kernel) -> K2 (tcgen05
cosine GEMM fused with the candidates' normalisation, the threshold and the running max/argmax) -> K3 (fp32 re-check
of near-tie / near-threshold rows) [-> K4 one NCCL allgather of the packed
{best_idx, keep} when N > 1].  Prints ONE JSON line (rank 0).  See DESIGN.md §6 for how every field is derived.

Workloads (BASELINE.json configs): the default, cfg3, is the per-GPU shard of configs[3] (10k references x 10M
candidates x 512-d, candidate-sharded over 8 GPUs -> 1.25M candidates per GPU, weak scaling: at --gpus 8 the job IS
configs[3]); cfg1 = configs[1] (1k x 100k x 128), cfg2 = configs[2] (10k x 1M x 512), cfg4 = per-GPU shard of
configs[4] (100k x 1.25M x 128).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    "cfg1": dict(n_ref=1_000, n_cand=100_000, dim=128, name="configs[1]: 1k ref x 100k cand x 128-d, threshold filter"),
    "cfg2": dict(n_ref=10_000, n_cand=1_000_000, dim=512, name="configs[2]: 10k ref x 1M cand x 512-d, max/argmax"),
    "cfg3": dict(n_ref=10_000, n_cand=1_250_000, dim=512,
                 name="configs[3] per-GPU shard: 10k ref x 1.25M cand x 512-d (10M candidates over 8 GPUs)"),
    "cfg4": dict(n_ref=100_000, n_cand=1_250_000, dim=128,
                 name="configs[4] per-GPU shard: 100k ref x 1.25M cand x 128-d (10M candidates over 8 GPUs)"),
    # the reference's literal mode (filter_faces_using_reference.py:186-189) at scale: ONE mean vector, Euclid keep test.
    # 0.5 FLOP/byte: the HBM-bound end of the path (exact fp32 streaming kernel K2s, no tensor cores)
    "n1": dict(n_ref=1, n_cand=10_000_000, dim=128, metric="euclid", thr=1.2,
               name="reference literal mode: 1 mean vector x 10M cand x 128-d, Euclid keep test (HBM-bound)"),
}
THR = 0.5
BLOCK = 62_500            # rows per synthetic block; data of a global row never depends on the GPU count
L2_BYTES = 126 * 1024 * 1024


def env_rank():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    return int(os.environ.get("RANK", "0")), world, int(os.environ.get("LOCAL_RANK", "0"))


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(tflops=d["bf16_tflops"], tflops_sustained=d.get("bf16_tflops_sustained"), hbm=d["hbm_gbs"],
                    source="MEASURED_PEAKS.json (of measured)")
    return dict(tflops=1590.0, tflops_sustained=1400.0, hbm=6650.0, source="B200_PROFILING.md fallback (of fallback)")


# ------------------------------------------------------------------------------------------------ data
def make_refs(n_ref, dim, device):
    import torch
    g = torch.Generator(device=device).manual_seed(42)           # the reference's seed (filter_faces...:24)
    return torch.nn.functional.normalize(torch.randn(n_ref, dim, device=device, generator=g))


def make_cands(ref, first_row, n_rows, device):
    """Unit-norm candidates for global rows [first_row, first_row + n_rows): even rows are planted matches of a random
    reference with cos in [0.55, 0.95], odd rows independent noise; every BLOCK rows has its own seed."""
    import torch
    n_ref, dim = ref.shape
    out = torch.empty(n_rows, dim, device=device)
    done = 0
    while done < n_rows:
        row = first_row + done
        blk, off = divmod(row, BLOCK)
        g = torch.Generator(device=device).manual_seed(1_000_003 + blk)
        noise = torch.nn.functional.normalize(torch.randn(BLOCK, dim, device=device, generator=g))
        k = torch.randint(0, n_ref, (BLOCK,), device=device, generator=g)
        c = torch.rand(BLOCK, device=device, generator=g) * 0.4 + 0.55
        planted = torch.nn.functional.normalize(c[:, None] * ref[k] + torch.sqrt(1 - c * c)[:, None] * noise)
        sel = (torch.arange(BLOCK, device=device) % 2 == 0)[:, None]
        block = torch.where(sel, planted, noise)
        take = min(BLOCK - off, n_rows - done)
        out[done:done + take] = block[off:off + take]
        done += take
    return out


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, uuid):
        self.rows, self.proc, self.uuid = [], None, uuid

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", self.uuid, f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [x.strip() for x in line.split(",")]))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        deadline = time.perf_counter() + 1.5                 # very short timed regions: wait for at least one sample
        while not self.rows and time.perf_counter() < deadline:
            time.sleep(0.02)
        time.sleep(0.06)
        self.proc.terminate()
        inside = [r for (t, r) in self.rows if t0 <= t <= t1 + 0.05] or [r for (_, r) in self.rows[-3:]]
        try:
            sm = [float(r[0]) for r in inside]
            reasons = set()
            for r in inside:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            return {"sm_mhz": statistics.median(sm), "sm_max_mhz": float(inside[0][1]),
                    "power_w_max": max(float(r[2]) for r in inside), "samples": len(inside), "reasons": sorted(reasons)}
        except Exception as e:                                                     # pragma: no cover
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [f"parse error: {e}"]}


# ------------------------------------------------------------------------------------------------ CPU baseline
def cpu_port_rate(n_ref, dim, n_sample, seconds_target=None, threads=None, metric="cosine", thr=THR):
    """Times the oracle (the CPU restatement of the reference's arithmetic) on a bounded candidate-axis sample of the
    workload: vectorised fp32 normalise + R @ C^T + max/argmax + threshold, all host threads, NumPy and torch-CPU
    (the faster one is reported), plus the reference's literal per-pair Python loop on a tiny sub-sample."""
    import numpy as np
    import torch
    from oracle import oracle
    if threads is None:                       # torchrun exports OMP_NUM_THREADS=1: use every core this process may run on
        threads = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    torch.set_num_threads(threads)
    rng = np.random.default_rng(42)
    ref = rng.standard_normal((n_ref, dim), dtype=np.float32)
    if metric == "euclid":
        # filter_faces_using_reference.py:186-189: vectorised (diff, row norm, <=) on all threads and the literal loop
        n_sample = int(min(n_sample, 4_000_000))
        cand = rng.standard_normal((n_sample, dim), dtype=np.float32)
        tc = torch.from_numpy(cand); tr = torch.from_numpy(ref)
        t = time.perf_counter(); k = (torch.linalg.vector_norm(tc - tr, dim=1) <= thr); t_torch = time.perf_counter() - t
        t = time.perf_counter(); k = (torch.linalg.vector_norm(tc - tr, dim=1) <= thr); t_torch = min(t_torch, time.perf_counter() - t)
        n_np = n_sample // 4
        t = time.perf_counter(); kn = oracle.filter_euclid(ref, cand[:n_np], thr)[0]; t_np = time.perf_counter() - t
        assert (k[:n_np].numpy().astype(np.uint8) == kn).mean() > 0.9999
        t = time.perf_counter(); oracle.euclid_keep_literal(cand[:100_000], ref, thr); r_lit = 100_000 / (time.perf_counter() - t)
        r_torch, r_np = n_ref * n_sample / t_torch, n_ref * n_np / t_np
        return dict(value=max(r_torch, r_np), torch_cpu=r_torch, numpy=r_np, literal_loop=r_lit, threads=threads,
                    n_sample=n_sample, seconds=t_torch)
    probe = rng.standard_normal((min(2048, n_sample), dim), dtype=np.float32)
    t = time.perf_counter(); oracle.filter_cosine_torch(ref, probe, THR); t_probe = time.perf_counter() - t
    t = time.perf_counter(); oracle.filter_cosine_torch(ref, probe, THR); t_probe = min(t_probe, time.perf_counter() - t)
    if seconds_target is not None:
        n_sample = int(max(2048, min(n_sample, probe.shape[0] * seconds_target / max(t_probe, 1e-6))))
    cand = rng.standard_normal((n_sample, dim), dtype=np.float32)
    t = time.perf_counter(); out_t = oracle.filter_cosine_torch(ref, cand, THR); t_torch = time.perf_counter() - t
    n_np = max(1024, n_sample // 8)
    t = time.perf_counter(); out_n = oracle.filter_cosine(ref, cand[:n_np], THR); t_np = time.perf_counter() - t
    assert (out_t[1][:n_np] == out_n[1]).mean() > 0.999
    r_torch, r_np = n_ref * n_sample / t_torch, n_ref * n_np / t_np
    # literal reference expression (extract_and_label_faces_from_dataset.py:106), Python double loop, 1 thread
    nl_r, nl_c = min(n_ref, 50), 200
    t = time.perf_counter()
    for j in range(nl_c):
        for i in range(nl_r):
            oracle.cosine_dist_literal(ref[i], cand[j])
    r_lit = nl_r * nl_c / (time.perf_counter() - t)
    return dict(value=max(r_torch, r_np), torch_cpu=r_torch, numpy=r_np, literal_loop=r_lit, threads=threads,
                n_sample=n_sample, seconds=t_torch)


def run_reference(args, rank, world, emit):
    """--impl reference: the reference's CPU arithmetic for this path (oracle port; the reference is pure Python/NumPy and
    its TensorFlow front end cannot run here, so there is no oracle/_ref) on the host cores, same config and metric."""
    if rank != 0:
        return
    import torch
    w = WORKLOADS[args.workload]
    n_ref, dim = w["n_ref"], w["dim"]
    metric, thr = w.get("metric", "cosine"), w.get("thr", THR)
    import numpy as np
    from oracle import oracle
    probe = cpu_port_rate(n_ref, dim, 4096 if metric == "cosine" else 400_000, metric=metric, thr=thr)
    n_sample = int(max(2048, min(w["n_cand"], 1.5 * probe["value"] / n_ref)))       # ~1.5 s per step
    rng = np.random.default_rng(42)
    ref = rng.standard_normal((n_ref, dim), dtype=np.float32)
    cand = rng.standard_normal((n_sample, dim), dtype=np.float32)
    if metric == "euclid":
        tr, tc = torch.from_numpy(ref), torch.from_numpy(cand)
        use_torch = probe["torch_cpu"] >= probe["numpy"]
        fn = (lambda r, c, t: (torch.linalg.vector_norm(tc - tr, dim=1) <= t)) if use_torch else oracle.filter_euclid
        how = ("torch-CPU" if use_torch else "NumPy") + " (diff, row norm, <=)"
    else:
        use_torch = probe["torch_cpu"] >= probe["numpy"]
        fn = oracle.filter_cosine_torch if use_torch else oracle.filter_cosine
        how = ("torch-CPU" if use_torch else "NumPy") + " sgemm + max/argmax + threshold"
    for _ in range(args.warmup):
        fn(ref, cand, thr)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        fn(ref, cand, thr)
    dt = time.perf_counter() - t0
    value = n_ref * n_sample * args.steps / dt
    cores = torch.get_num_threads()
    sample = (f"{n_sample} of {w['n_cand']} candidates per step x all {n_ref} references x {dim}-d, vectorised fp32 "
              f"restatement ({how}), {cores} threads; the reference's literal per-row Python loop = "
              f"{probe['literal_loop']:.3g} pairs/s on 1 thread")
    emit({
        "impl": "reference", "metric": "face pairs compared/sec (ref x cand cosine+filter)", "value": value,
        "unit": "pairs/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": w["name"], "n_ref": n_ref, "n_cand_per_gpu": w["n_cand"], "dim": dim, "threshold": thr,
                   "metric": metric},
        "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    })


# ------------------------------------------------------------------------------------------------ GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="cfg3", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    rank, world, local_rank = env_rank()
    # the contract is ONE JSON line on stdout: anything libraries print while we run (NCCL's version banner, warnings)
    # is sent to stderr by pointing fd 1 at fd 2 until the result line is written
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        print(json.dumps(obj), flush=True)
        os.dup2(2, 1)

    if args.impl == "reference":
        return run_reference(args, rank, world, emit)

    import torch
    import torch.distributed as dist
    from face_detection_and_recognition_b200 import _lib, ops

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the filter has no CPU path (use --impl reference for the CPU arm)")
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    w = WORKLOADS[args.workload]
    n_ref, n_cand, dim = w["n_ref"], w["n_cand"], w["dim"]
    metric, thr = w.get("metric", "cosine"), w.get("thr", THR)
    pk = peaks()

    ref = make_refs(n_ref, dim, dev)
    in_bytes = (n_cand + n_ref) * dim * 4
    n_buf = 1 if in_bytes > 2 * L2_BYTES else int(-(-3 * L2_BYTES // in_bytes))   # rotate copies when L2 could hold the input
    cands = [make_cands(ref, rank * n_cand + (b * world * n_cand if b else 0), n_cand, dev) for b in range(n_buf)]
    keep = torch.empty(n_cand, dtype=torch.uint8, device=dev)
    idx = torch.empty(n_cand, dtype=torch.int32, device=dev)
    val = torch.empty(n_cand, dtype=torch.float32, device=dev)
    gather = ops.ResultGather(rank, world, local_rank) if world > 1 else None

    def step(i):
        r = ops.face_filter(ref, cands[i % n_buf], thr, metric=metric, out=(keep, idx, val))
        if gather is not None:
            return gather.all_gather(r.keep, r.best_idx)
        return r.keep, r.best_idx

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler("GPU-" + str(torch.cuda.get_device_properties(local_rank).uuid)) if rank == 0 else None
    if sampler:
        sampler.start()
    for i in range(args.warmup):
        step(i)
    k2_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for a, b in k2_ev:
        a.record(); b.record()                                      # materialise the cudaEvent_t handles
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    l0 = ops.launch_count()
    t_host0 = time.perf_counter()
    ev0.record()
    for i in range(args.steps):
        lib.ffr_debug_set_k2_events(k2_ev[i][0].cuda_event, k2_ev[i][1].cuda_event)
        step(i)
    ev1.record()
    lib.ffr_debug_set_k2_events(None, None)
    barrier()
    t_host1 = time.perf_counter()
    launches = ops.launch_count() - l0
    ms_total = ev0.elapsed_time(ev1)
    tensor_path = metric == "cosine" and n_ref > 8
    k2_ms = statistics.mean(a.elapsed_time(b) for a, b in k2_ev) if tensor_path else ms_total / args.steps
    stats = ops.face_filter(ref, cands[0], thr, metric=metric, out=(keep, idx, val), want_stats=True).stats
    keep_frac = float(keep.float().mean())
    clocks = sampler.stop(t_host0, t_host1) if sampler else None

    # ---- end to end through the host-buffer C-ABI call (pinned host memory, H2D + D2H inside the timed region)
    e2e = None
    if not args.no_e2e:
        ref_h = ref.cpu().pin_memory()
        cand_h = torch.empty((n_cand, dim), dtype=torch.float32).pin_memory()
        cand_h.copy_(cands[0])
        out_h = (torch.empty(n_cand, dtype=torch.uint8).pin_memory(), torch.empty(n_cand, dtype=torch.int32).pin_memory(),
                 torch.empty(n_cand, dtype=torch.float32).pin_memory())
        hf = ops.HostFilter(device=local_rank, max_ref=n_ref, chunk_cand=min(n_cand, 1 << 17), max_dim=dim)
        hf(ref_h, cand_h, thr, metric=metric, out=out_h)
        e2e_steps = max(2, min(args.steps, 10))
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            hf(ref_h, cand_h, thr, metric=metric, out=out_h)
        t_e2e = time.perf_counter() - t0
        same = bool(torch.equal(out_h[0], keep.cpu()) and torch.equal(out_h[1], idx.cpu()))
        hf.close()
        e2e = dict(seconds=t_e2e / e2e_steps, same_as_device_path=same, steps=e2e_steps)

    # ---- reduce over ranks: slowest rank's time, total launches
    if world > 1:
        t = torch.tensor([ms_total, k2_ms, e2e["seconds"] if e2e else 0.0], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total, k2_ms = float(t[0]), float(t[1])
        if e2e:
            e2e["seconds"] = float(t[2])
        lt = torch.tensor([launches], device=dev, dtype=torch.int64)
        dist.all_reduce(lt)
        launches = int(lt[0])

    if rank == 0:
        ms_step = ms_total / args.steps
        pairs_job = n_ref * n_cand * world
        flops_k2 = 2.0 * n_ref * n_cand * dim
        hbm_bytes = 4.0 * dim * (n_ref + n_cand) + 9.0 * n_cand
        ach = flops_k2 / (k2_ms * 1e-3) / 1e12
        traffic = None
        tp = os.path.join(ROOT, "profiles", "k2_traffic.json")
        if os.path.exists(tp):
            traffic = json.load(open(tp)).get(args.workload)
        out = {
            "metric": "face pairs compared/sec (ref x cand cosine+filter)",
            "value": pairs_job / (ms_step * 1e-3), "unit": "pairs/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f16->f32" if tensor_path else "f32", "data": "synthetic",
            "config": {"workload": w["name"], "n_ref": n_ref, "n_cand_per_gpu": n_cand, "dim": dim, "threshold": thr,
                       "metric": metric,
                       "arithmetic": ("tcgen05 kind::f16: fp16 operands, fp32 accumulate in TMEM; fp32 re-check of near-tie / "
                                      "near-threshold rows") if tensor_path else "fp32 CUDA cores (exact streaming kernel)",
                       "sharding": "candidate axis; references replicated; one NCCL allgather of 5 B/candidate"
                       if world > 1 else "single GPU",
                       "l2": (f"inputs {in_bytes / 1e6:.0f} MB per GPU > L2 (126 MB)" if n_buf == 1 else
                              f"rotating {n_buf} input copies ({n_buf * in_bytes / 1e6:.0f} MB > L2 126 MB)"),
                       "keep_fraction": keep_frac, "recheck": stats},
            "roofline": {"bound": "tensor", "kernel": "filter_mma_kernel (K2)", "achieved": ach, "peak": pk["tflops"],
                         "unit": "TFLOP/s", "frac": ach / pk["tflops"],
                         "frac_of_sustained": ach / pk["tflops_sustained"] if pk["tflops_sustained"] else None,
                         "peak_source": pk["source"], "k2_ms": k2_ms, "k2_share_of_step": k2_ms / ms_step,
                         "flops_per_launch": flops_k2, "traffic": traffic,
                         "step_hbm": {"algorithmic_bytes": hbm_bytes, "gbs_over_step": hbm_bytes / (ms_step * 1e-3) / 1e9,
                                      "frac_of_hbm_peak": hbm_bytes / (ms_step * 1e-3) / 1e9 / pk["hbm"]}},
            "gpu_launches": launches, "clocks": clocks,
        }
        if not tensor_path:            # one HBM-bound kernel: the roofline is the measured copy bandwidth
            gbs = hbm_bytes / (ms_step * 1e-3) / 1e9
            out["roofline"] = {"bound": "hbm", "kernel": "filter_fp32_kernel (K2s)", "achieved": gbs, "peak": pk["hbm"],
                               "unit": "GB/s", "frac": gbs / pk["hbm"], "peak_source": pk["source"],
                               "bytes_per_launch": hbm_bytes, "traffic": traffic}
        if e2e:
            out["e2e"] = {"value": pairs_job / e2e["seconds"], "unit": "pairs/s",
                          "h2d_bytes_per_step": int((n_cand + n_ref) * dim * 4), "d2h_bytes_per_step": int(9 * n_cand),
                          "ms_per_step": e2e["seconds"] * 1e3, "steps": e2e["steps"],
                          "api": "ffr_ctx_filter_host (pinned host buffers, chunked H2D overlapped with compute)",
                          "same_result_as_device_path": e2e["same_as_device_path"]}
        if world == 1 and not args.no_cpu_baseline:
            cb = cpu_port_rate(n_ref, dim, n_cand, seconds_target=10.0, metric=metric, thr=thr)
            out["cpu_baseline"] = {
                "value": cb["value"], "unit": "pairs/s", "cores": cb["threads"], "kind": "port",
                "sample": (f"{cb['n_sample']} of {n_cand} candidates x all {n_ref} references x {dim}-d "
                           f"({cb['seconds']:.1f} s): vectorised fp32 oracle, torch-CPU {cb['torch_cpu']:.3g} / NumPy "
                           f"{cb['numpy']:.3g} pairs/s on {cb['threads']} threads; the reference's literal per-pair Python "
                           f"loop: {cb['literal_loop']:.3g} pairs/s on 1 thread")}
        emit(out)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
