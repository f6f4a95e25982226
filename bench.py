#!/usr/bin/env python
"""bench.py — decode tokens/s of the B200 engine.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload 1b-bf16|1b-w4|8b-bf16|70b-bf16|*-prefill] [--batch B]
                    [--per-op] [--prompt S] [--replicas] [--no-also]

N = 1 (default): BASELINE.json configs[1] -- Llama-3.2-1B-shaped bf16, batch 1, 512-token KV cache, random-init weights, synthetic
ids -- is the headline line; the other single-GPU workloads of the metric (1B int4 batch 1 = the north-star target, 1B bf16 batch 32,
1B prefill 2048, 8B bf16 batch 1) are measured in the SAME run and ride in `also` so that the driver observes them too.
N > 1: ONE Llama-3.1-8B-shaped bf16 model (configs[3], decode phase) sharded tensor-parallel over the N GPUs, the same workload at
every N ("scaling": "strong"); its single-GPU figure is `also[workload == 8b-bf16]` of the N = 1 line and, measured again on rank 0 of
the N > 1 run, `strong_scaling_base`.  `--replicas` runs N independent copies of the N = 1 workload instead (weak scaling).

A "step" is one decode step (one token per sequence).  Prints ONE JSON line (see DESIGN.md "Measurement").  `value` is device-timed
(CUDA events on the engine's stream, token fed back on the device); `e2e` goes through mc_llama_decode with host buffers every step
(H2D ids/pos, D2H id).  `--impl reference` times the CPU oracle port of the reference path (the reference is Metal-only).
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
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

SHAPES = {
    "1b": dict(dim=2048, n_layers=16, n_heads=32, n_kv_heads=8, head_dim=64, ffn_dim=8192, vocab=128256),
    "8b": dict(dim=4096, n_layers=32, n_heads=32, n_kv_heads=8, head_dim=128, ffn_dim=14336, vocab=128256),
    "70b": dict(dim=8192, n_layers=80, n_heads=64, n_kv_heads=8, head_dim=128, ffn_dim=28672, vocab=128256),  # 141 GB in bf16: needs --tp
}
KV_LEN = 512
METRIC = "decode_tokens_per_s"


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples nvidia-smi clocks and throttle reasons during the timed region."""

    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.index, self.rows, self.stop = index, [], threading.Event()
        self.t = threading.Thread(target=self.run, daemon=True)

    def run(self):
        while not self.stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self.stop.wait(0.1)

    def __enter__(self):
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.t.join(timeout=6)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = [float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[2 + i].lower().startswith("active") for r in self.rows if len(r) > 2 + i)]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": float(self.rows[0][1]) if self.rows[0][1].isdigit() else None,
                "reasons": reasons, "samples": len(self.rows)}


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


def host_threads_for_reference():
    """torchrun exports OMP_NUM_THREADS=1 to every rank; the reference arm runs on rank 0 alone and may use every host core it is
    allowed to (set before the oracle library -- and with it the OpenMP runtime -- is loaded)."""
    if "TORCHELASTIC_RUN_ID" in os.environ or int(os.environ.get("WORLD_SIZE", "1")) > 1:
        try:
            n = len(os.sched_getaffinity(0))
        except AttributeError:
            n = os.cpu_count() or 1
        os.environ["OMP_NUM_THREADS"] = str(n)


def cpu_baseline(shape: dict, quant: int, steps: int):
    """The oracle port of the reference path timed on this host's cores (checker code, timed as the baseline)."""
    host_threads_for_reference()
    from oracle import orc

    try:  # torchrun exports OMP_NUM_THREADS=1 and an OpenMP runtime may already be up: size the pool explicitly
        orc.set_num_threads(len(os.sched_getaffinity(0)))
    except AttributeError:
        orc.set_num_threads(os.cpu_count() or 1)

    cfg = orc.make_cfg(**shape, max_seq_len=1024, quant=quant)
    m = orc.Llama(cfg, orc.BF16)
    m.init_random(0x5EED)
    m.decode_timed(1, KV_LEN, 1)  # warm-up (page-in, thread pool)
    sec, _ = m.decode_timed(1, KV_LEN + 1, steps)
    threads = orc.num_threads()
    m.close()
    return {"value": steps / sec, "unit": "tokens/s", "cores": threads, "kind": "port",
            "sample": f"{steps} greedy decode steps of the full model at KV length {KV_LEN} (zero-filled cache), scalar C++ oracle with OpenMP over output rows",
            "ms_per_step": 1e3 * sec / steps}


def decode_config(args, workload, world, tp_on):
    """The `config` object of a decode line: the same for this repository's arm and for the reference arm (the driver compares them)."""
    if tp_on:
        par = (f"tp{world}: ONE model, column/row-split blocks, all-reduce fused into the kernels over NVLink peer memory; roofline per GPU shard"
               if args.tp_collective == "fused" else
               f"tp{world}: ONE model, column/row-split blocks, COMPARATOR: two ncclAllReduce calls per block between the per-op kernels (CUDA graph); roofline per GPU shard")
    else:
        par = f"{world} replica(s), one sequence stream per GPU"
    return {"workload": workload, "kv_len": KV_LEN, "batch": args.batch, "parallelism": par,
            "l2": "the weights streamed per step (0.85 GB for the smallest workload) exceed the 126 MB L2: no flush needed between steps"}


def run_reference(args, shape, quant, workload):
    rank, world, _ = dist_env()
    if rank != 0:
        return
    tp_on = world > 1 and not args.replicas
    host_threads_for_reference()
    from oracle import orc

    try:
        orc.set_num_threads(len(os.sched_getaffinity(0)))
    except AttributeError:
        orc.set_num_threads(os.cpu_count() or 1)
    cfg = orc.make_cfg(**shape, max_seq_len=1024, quant=quant)
    m = orc.Llama(cfg, orc.BF16)
    m.init_random(0x5EED)
    budget_s = 150.0
    t_probe, _ = m.decode_timed(1, KV_LEN, 1)
    warm = max(0, min(args.warmup, int(10.0 / max(t_probe, 1e-3))))
    if warm:
        m.decode_timed(1, KV_LEN, warm)
    steps = max(1, min(args.steps, int(budget_s / max(t_probe, 1e-3))))
    sec, _ = m.decode_timed(1, KV_LEN, steps)
    v = steps / sec
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": "tokens/s", "n_gpus": args.gpus, "steps": steps, "warmup": warm,
        "ms_per_step": 1e3 * sec / steps, "higher_is_better": True, "scaling": "strong" if tp_on else "weak", "vs_baseline": None,
        "dtype": "bf16" if not quant else "bf16 activations, int4 weights (bf16 dequant), fp32 accumulate",
        "data": "synthetic ids, random-init weights (counter-hash seed 0x5EED)",
        "config": decode_config(args, workload, world, tp_on),
        "cpu_baseline": {"value": v, "unit": "tokens/s", "cores": orc.num_threads(), "kind": "port",
                         "sample": f"{steps} decode steps at KV length {KV_LEN}; the reference is Metal-only, so this is the scalar C++ oracle port on host cores"},
        "e2e": {"value": v, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def measure_gemv_family(capi, dev, shape, hbm_peak):
    """Isolated CUDA-event timing of every GEMV shape of one decode step (weights larger than L2 in aggregate:
    each shape is run over a ring of distinct weight buffers so that no launch re-reads L2-resident data)."""
    import ctypes as C

    try:
        import torch
    except Exception:
        torch = None
    D, F, V = shape["dim"], shape["ffn_dim"], shape["vocab"]
    QKV = (shape["n_heads"] + 2 * shape["n_kv_heads"]) * shape["head_dim"]
    QO = shape["n_heads"] * shape["head_dim"]
    L = shape["n_layers"]
    shapes = [("wqkv", QKV, D, L), ("wo", D, QO, L), ("w13", 2 * F, D, L), ("w2", D, F, L), ("head", V, D, 1)]
    out = []
    tot_b = tot_t = 0.0
    for name, N, K, per_step in shapes:
        wbytes = N * K * 2
        ring = max(2, min(16, int(400e6 // wbytes) + 1))  # > 126 MB L2 in aggregate
        ws = [dev.alloc(wbytes) for _ in range(ring)]
        for w in ws:
            capi.check(capi.lib().mc_memset(dev.h, w.h, 0, 0x3c, wbytes))
        x = dev.upload(np.full(K, 0x3c00, np.uint16))
        y = dev.alloc(N * 2)
        reps = max(ring * 3, 24)
        for i in range(ring):
            capi.linear_bf16(dev, y, x, ws[i % ring], 1, N, K)
        dev.synchronize()
        t0 = time.perf_counter()
        if torch is not None:
            # events on the engine's own stream via an external stream handle
            st = torch.cuda.ExternalStream(dev.stream())
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            for i in range(reps):
                capi.linear_bf16(dev, y, x, ws[i % ring], 1, N, K)
            e1.record(st)
            dev.synchronize()
            ms = e0.elapsed_time(e1) / reps
        else:
            for i in range(reps):
                capi.linear_bf16(dev, y, x, ws[i % ring], 1, N, K)
            dev.synchronize()
            ms = 1e3 * (time.perf_counter() - t0) / reps
        alg = wbytes + K * 2 + N * 2
        gbs = alg / (ms * 1e-3) / 1e9
        out.append({"kernel": f"gemv_bf16 {name} [{N}x{K}]", "us": ms * 1e3, "GBps": gbs, "frac": gbs / hbm_peak, "launches_per_step": per_step})
        tot_b += alg * per_step
        tot_t += ms * 1e-3 * per_step
        for w in ws:
            w.release()
    return out, tot_b / tot_t / 1e9


def init_dist(local: int):
    """One process per GPU over NCCL.  The communicator is created (first collective) with stdout pointed at stderr: NCCL prints its
    version banner on stdout, and this program's stdout carries exactly one JSON line."""
    import torch
    import torch.distributed as dist

    torch.cuda.set_device(local)
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    try:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        dist.barrier()
        torch.cuda.synchronize()
    finally:
        sys.stdout.flush()
        os.dup2(saved, 1)
        os.close(saved)
    return dist


def tensor_peak():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return float(d["bf16_tflops_sustained"]), "measured (MEASURED_PEAKS.json bf16_tflops_sustained: the GEMMs run back to back inside a long step)"
    return 1373.0, "fallback (B200_PROFILING.md sustained)"


def prefill_flops(shape: dict, S: int):
    """Algorithmic flops of one prompt of S positions: 2 flop per (weight, position) in the blocks, causal attention
    (half of the S x S score / value products), the vocabulary projection of the LAST position only (nn/llama.h:128-133)."""
    D, F, hd, H, KV, L = shape["dim"], shape["ffn_dim"], shape["head_dim"], shape["n_heads"], shape["n_kv_heads"], shape["n_layers"]
    per_layer_w = D * (H + 2 * KV) * hd + H * hd * D + 3 * D * F
    linear = 2.0 * per_layer_w * S * L
    attn = 2.0 * (S * (S + 1) / 2) * hd * H * 2 * L
    head = 2.0 * shape["vocab"] * D
    return linear, attn, head


def run_prefill(args, shape_name, shape, quant=0):
    """Prompt throughput (BASELINE.json metric "prefill tok/s @2048"): a step = one prompt of --prompt positions through the
    tensor-core prompt path (tcgen05 GEMMs + causal attention), logits of the last position produced."""
    rank, world, local = dist_env()
    S = args.prompt
    fmt_name = "QLoRA int4 (resident bf16 image + adaptor kernels)" if quant else "bf16"
    workload = f"llama-{shape_name} {fmt_name} prefill of {S} positions (BASELINE.json configs[3] prompt phase), 1 sequence per GPU"
    if args.impl == "reference":
        if rank != 0:
            return
        host_threads_for_reference()
        from oracle import orc

        n = 64
        m = orc.Llama(orc.make_cfg(**shape, max_seq_len=max(128, n), quant=quant), orc.BF16)
        m.init_random(0x5EED)
        ids = np.random.default_rng(1).integers(0, shape["vocab"], size=n).tolist()
        t0 = time.perf_counter()
        m.forward(ids, 0)
        sec = time.perf_counter() - t0
        v = n / sec
        print(json.dumps({"impl": "reference", "metric": "prefill_tokens_per_s", "value": v, "unit": "tokens/s", "n_gpus": args.gpus, "steps": 1, "warmup": 0,
                          "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
                          "data": "synthetic ids, random-init weights (counter-hash seed 0x5EED)", "config": {"workload": workload, "prompt": S},
                          "cpu_baseline": {"value": v, "unit": "tokens/s", "cores": orc.num_threads(), "kind": "port",
                                           "sample": f"one {n}-position prompt through the full model (scalar C++ oracle port, OpenMP over output rows)"},
                          "e2e": {"value": v, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}), flush=True)
        return
    import torch

    dist = init_dist(local) if world > 1 else None
    from metalchat_b200 import capi

    dev = capi.Device(local)
    tp_on = world > 1 and not args.replicas  # ONE model sharded over the GPUs (every rank feeds the same prompt) instead of one replica per GPU
    pf_flags = (capi.LLAMA_NO_TC_PREFILL if args.per_op else 0) | (capi.LLAMA_W4_PACKED if quant else 0)
    if tp_on:
        from metalchat_b200 import tp

        m = tp.create(dev, **shape, max_seq_len=S, quant=quant, n_seqs=1, flags=pf_flags)
    else:
        m = capi.Llama(dev, capi.llama_config(**shape, max_seq_len=S, quant=quant, n_seqs=1, flags=pf_flags))
    m.init_random(0x5EED)
    m.finalize()
    rng = np.random.default_rng(0x5EED + (0 if tp_on else rank))
    prompts = [rng.integers(0, shape["vocab"], size=S, dtype=np.int32) for _ in range(4)]
    steps = min(args.steps, 32)

    def barrier():
        dev.synchronize()
        if dist is not None:
            torch.cuda.synchronize()
            dist.barrier()

    for i in range(args.warmup):
        m.prefill(prompts[i % 4], 0, 0)
    barrier()
    st = torch.cuda.ExternalStream(dev.stream())
    l0 = dev.launches()
    with ClockSampler(local) as clocks:
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for i in range(steps):
            m.prefill(prompts[i % 4], 0, 0)
        e1.record(st)
        dev.synchronize()
        ms = e0.elapsed_time(e1)
        launches = dev.launches() - l0
        barrier()
        # end to end: host ids in, logits of the last position back on the host, every step
        t0 = time.perf_counter()
        for i in range(steps):
            m.prefill(prompts[i % 4], 0, 0)
            lg = m.logits(0)
        e2e_ms = (time.perf_counter() - t0) * 1e3
    if dist is not None:
        t = torch.tensor([ms, e2e_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_ms = float(t[0]), float(t[1])
    if rank != 0:
        dist.destroy_process_group()
        return
    lin, att, head = prefill_flops(shape, S)
    if tp_on:
        lin, att, head = lin / world, att / world, head / world  # the roofline is per GPU shard
    peak, peak_src = tensor_peak()
    tf = (lin + att + head) / (ms / steps * 1e-3) / 1e12
    roof = {"bound": "tensor", "achieved": tf, "peak": peak, "unit": "TFLOP/s", "frac": tf / peak, "traffic": None, "peak_source": peak_src,
            "kernel": "whole prompt step (gemm_tc_kernel = tcgen05 GEMMs carry %.1f%% of the flops; causal attention on mma.sync)" % (100 * lin / (lin + att + head)),
            "algorithmic_flops_per_step": lin + att + head}
    if not args.per_op and not quant and not tp_on:
        # the dominant kernel alone: the four GEMM shapes of a block, CUDA events on the engine stream (mc_gemm_bf16)
        D, F = shape["dim"], shape["ffn_dim"]
        QKV = (shape["n_heads"] + 2 * shape["n_kv_heads"]) * shape["head_dim"]
        QO = shape["n_heads"] * shape["head_dim"]
        per = []
        tot_f = tot_t = 0.0
        for name, N, K, mode in [("wqkv", QKV, D, 0), ("wo", D, QO, 2), ("w13", 2 * F, D, 3), ("w2", D, F, 2)]:
            x, w = dev.alloc(S * K * 2), dev.alloc(N * K * 2)
            ncols = N // 2 if mode == 3 else N
            y, r = dev.alloc(S * ncols * 2), dev.alloc(S * N * 2) if mode == 2 else None
            for b, n in ((x, S * K * 2), (w, N * K * 2)) + (((r, S * N * 2),) if r is not None else ()):
                capi.check(capi.lib().mc_memset(dev.h, b.h, 0, 0x3c, n))
            capi.gemm_bf16(dev, y, x, w, S, N, K, mode=mode, res=r, iters=3)
            g_ms = capi.gemm_bf16(dev, y, x, w, S, N, K, mode=mode, res=r, iters=20) / 20
            fl = 2.0 * S * N * K
            per.append({"kernel": f"gemm_tc_kernel {name} [{S}x{N}x{K}]", "us": g_ms * 1e3, "TFLOPs": fl / (g_ms * 1e-3) / 1e12, "frac": fl / (g_ms * 1e-3) / 1e12 / peak})
            tot_f += fl
            tot_t += g_ms * 1e-3
            for b in (x, w, y, r):
                if b is not None:
                    b.release()
        roof["gemm_tc_kernel"] = {"achieved": tot_f / tot_t / 1e12, "frac": tot_f / tot_t / 1e12 / peak, "per_shape": per,
                                  "note": "the four GEMMs of one block timed alone (20 back-to-back launches each, operands L2-resident between launches)"}
    tokens = steps * S * (1 if tp_on else world)
    line = {
        "metric": "prefill_tokens_per_s", "value": tokens / (ms * 1e-3), "unit": "tokens/s", "n_gpus": world, "steps": steps, "warmup": args.warmup,
        "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "strong" if tp_on else "weak", "vs_baseline": None, "dtype": "bf16" if not quant else "bf16 activations, int4 weights (bf16 dequant), fp32 accumulate",
        "data": "synthetic ids, random-init weights (counter-hash seed 0x5EED)",
        "config": {"workload": workload, "prompt": S,
                   "parallelism": (f"tp{world}: ONE model, column/row-split GEMMs, fp32 partial sums all-reduced by tc::tp_allreduce_rows over NVLink peer memory; roofline per GPU shard"
                                   if tp_on else f"{world} replica(s)"),
                   "path": "4-row GEMV prompt path" if args.per_op else "tcgen05 GEMM prompt path",
                   "l2": "weights of one pass (%.0f MB) > 126 MB L2; four distinct prompts cycled" % (m.weight_bytes()[0] / 1e6)},
        "e2e": {"value": tokens / (e2e_ms * 1e-3), "unit": "tokens/s", "h2d_bytes_per_step": 4 * S, "d2h_bytes_per_step": 2 * shape["vocab"],
                "ms_per_step": e2e_ms / steps},
        "gpu_launches": int(launches), "launches_per_step": launches // steps, "clocks": clocks.summary(), "roofline": roof,
    }
    if not args.no_cpu_baseline:
        from oracle import orc

        n = 32
        o = orc.Llama(orc.make_cfg(**shape, max_seq_len=128, quant=quant), orc.BF16)
        o.init_random(0x5EED)
        t0 = time.perf_counter()
        o.forward(prompts[0][:n].tolist(), 0)
        sec = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": n / sec, "unit": "tokens/s", "cores": orc.num_threads(), "kind": "port",
                                "sample": f"one {n}-position prompt through the full model (scalar C++ oracle port, OpenMP over output rows)"}
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def workload_name(shape_name, fmt, batch, quant):
    idx = {"1b": 1 if quant == 0 else 2, "8b": 3, "70b": 4}.get(shape_name, 1)
    return f"llama-{'3.2' if shape_name == '1b' else '3.1'}-{shape_name} {fmt} decode, batch {batch}, {KV_LEN}-token KV cache (BASELINE.json configs[{idx}])"


def measure_decode(capi, torch, dist, dev, shape_name, fmt, batch, steps, warmup, rank, world, local, tp_on=False, per_op=False, e2e=True,
                   clocks_index=None, workload_key=None, collective="fused"):
    """One decode workload on this rank's GPU (or this rank's shard when tp_on): builds the model, fills a KV_LEN-token cache per
    sequence with the engine's own prefill, warms up, then times `steps` decode steps (CUDA events on the engine's stream inside
    mc_llama_decode_loop; max over ranks) and the same number of steps through the per-token call with host buffers."""
    shape = SHAPES[shape_name]
    quant = 0 if fmt == "bf16" else 1
    flags = (capi.LLAMA_W4_PACKED if quant else 0) | (capi.LLAMA_NO_STREAM if per_op else 0)
    steps = max(1, min(steps, 1024 - KV_LEN - warmup - 8))
    if tp_on:
        from metalchat_b200 import tp

        m = tp.create(dev, collective=collective, **shape, max_seq_len=1024, quant=quant, n_seqs=batch, flags=flags)
    else:
        m = capi.Llama(dev, capi.llama_config(**shape, max_seq_len=1024, quant=quant, n_seqs=batch, flags=flags))
    m.init_random(0x5EED)
    m.finalize()
    streamed, resident = m.weight_bytes()
    B = batch
    rng = np.random.default_rng(0x5EED + (0 if tp_on else rank))  # tensor parallel: every rank feeds the same ids
    for sq in range(B):
        m.prefill(rng.integers(0, shape["vocab"], size=KV_LEN, dtype=np.int32), 0, sq)
    if tp_on:
        # the vocabulary is sharded: one greedy step on the device gives every rank the global argmax
        first = m.decode(np.zeros(B, np.int32), np.full(B, KV_LEN, np.int32)).tolist()
    else:
        first = [int(np.argmax((m.logits(sq).astype(np.uint32) << 16).view(np.float32))) for sq in range(B)]
    pos0 = [KV_LEN] * B

    def barrier():
        dev.synchronize()
        if dist is not None:
            torch.cuda.synchronize()
            dist.barrier()

    m.decode_loop(first, pos0, warmup)  # warm-up (instantiates the CUDA graph)
    barrier()
    l0 = dev.launches()
    sampler = ClockSampler(local if clocks_index is None else clocks_index)
    with sampler as clocks:
        barrier()
        toks, ms = m.decode_loop(first, [KV_LEN + warmup] * B, steps)
        barrier()
        e2e_ms = None
        if e2e:
            ids = np.array(first, np.int32)
            t0 = time.perf_counter()
            for i in range(steps):
                ids = m.decode(ids, np.full(B, KV_LEN + warmup + i, np.int32))
            dev.synchronize()
            e2e_ms = (time.perf_counter() - t0) * 1e3
    launches = dev.launches() - l0
    if dist is not None:
        t = torch.tensor([ms, e2e_ms or 0.0], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_max = float(t[0]), float(t[1])
        e2e_ms = e2e_max if e2e else None
    hbm_peak, peak_src = peaks()
    tokens = steps * B * (1 if tp_on else world)
    kv_bytes = 2 * shape["n_layers"] * (shape["n_kv_heads"] // (world if tp_on else 1)) * shape["head_dim"] * 2 * (KV_LEN + warmup + steps // 2) * B
    step_bytes = streamed + kv_bytes
    step_gbs = step_bytes / (ms * 1e-3 / steps) / 1e9
    streaming = m.launches_per_step() == 1
    kname = ("decode_stream_kernel (persistent: the whole decode step is one kernel; `steps` steps per launch)" if streaming
             else "whole decode step (per-op kernels under one CUDA graph; the GEMV / GEMM family streams >97% of the bytes)")
    # algorithmic bytes of one step (weights + norms + adaptors streamed once, KV read once) / device time of one step, both of the
    # timed region above (CUDA events on the engine's stream around the launches)
    roof = {"bound": "hbm", "achieved": step_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": step_gbs / hbm_peak, "traffic": None,
            "kernel": kname, "peak_source": peak_src, "algorithmic_bytes_per_step": step_bytes}
    ncu = ROOT / "profiles" / "r02_ncu_stream_summary.json"
    if not ncu.exists():
        ncu = ROOT / "profiles" / "r01_ncu_stream_summary.json"
    if streaming and ncu.exists():
        try:
            t = json.loads(ncu.read_text()).get((workload_key or f"{shape_name}-{fmt}") if B == 1 else "", {})
            if t:
                roof["traffic"] = t["dram_bytes_read"] + t["dram_bytes_write"]
                roof["traffic_source"] = t.get("source", "ncu --set full, one launch of one step") + f" ({ncu.name})"
        except Exception:
            pass
    res = {
        "workload": workload_name(shape_name, fmt, B, quant), "shape": shape, "quant": quant, "batch": B, "steps": steps, "ms": ms, "tokens": tokens,
        "value": tokens / (ms * 1e-3), "ms_per_step": ms / steps, "e2e_value": (tokens / (e2e_ms * 1e-3)) if e2e_ms else None,
        "e2e_ms_per_step": (e2e_ms / steps) if e2e_ms else None, "launches": int(launches), "launches_per_step": m.launches_per_step(),
        "streaming": streaming, "streamed": streamed, "resident": resident, "roofline": roof, "clocks": clocks.summary(),
        "dtype": "bf16" if quant == 0 else "bf16 activations, int4 weights (bf16 dequant), fp32 accumulate",
    }
    m.close()
    return res


def measure_prefill_brief(capi, torch, dev, shape_name, quant, S, steps, warmup):
    """Prompt throughput of one sequence on this GPU for the `also` list (the full line: --workload 1b-bf16-prefill)."""
    shape = SHAPES[shape_name]
    m = capi.Llama(dev, capi.llama_config(**shape, max_seq_len=S, quant=quant, n_seqs=1, flags=capi.LLAMA_W4_PACKED if quant else 0))
    m.init_random(0x5EED)
    m.finalize()
    rng = np.random.default_rng(0x5EED)
    prompts = [rng.integers(0, shape["vocab"], size=S, dtype=np.int32) for _ in range(4)]
    for i in range(warmup):
        m.prefill(prompts[i % 4], 0, 0)
    dev.synchronize()
    st = torch.cuda.ExternalStream(dev.stream())
    l0 = dev.launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for i in range(steps):
        m.prefill(prompts[i % 4], 0, 0)
    e1.record(st)
    dev.synchronize()
    ms = e0.elapsed_time(e1)
    launches = dev.launches() - l0
    t0 = time.perf_counter()
    for i in range(steps):
        m.prefill(prompts[i % 4], 0, 0)
        m.logits(0)
    e2e_ms = (time.perf_counter() - t0) * 1e3
    lin, att, head = prefill_flops(shape, S)
    peak, peak_src = tensor_peak()
    tf = (lin + att + head) / (ms / steps * 1e-3) / 1e12
    m.close()
    return {"workload": f"llama-{shape_name} {'QLoRA int4' if quant else 'bf16'} prefill of {S} positions (BASELINE.json configs[3] prompt phase, 1B shape), 1 sequence",
            "metric": "prefill_tokens_per_s", "value": steps * S / (ms * 1e-3), "unit": "tokens/s", "steps": steps, "ms_per_step": ms / steps,
            "e2e": {"value": steps * S / (e2e_ms * 1e-3), "unit": "tokens/s", "h2d_bytes_per_step": 4 * S, "d2h_bytes_per_step": 2 * shape["vocab"]},
            "gpu_launches": int(launches),
            "roofline": {"bound": "tensor", "achieved": tf, "peak": peak, "unit": "TFLOP/s", "frac": tf / peak, "traffic": None, "peak_source": peak_src,
                         "kernel": "whole prompt step (tcgen05 GEMMs + causal attention)", "algorithmic_flops_per_step": lin + att + head}}


def brief(res):
    """Compact `also` entry of a decode measurement."""
    r = res["roofline"]
    return {"workload": res["workload"], "metric": METRIC, "value": res["value"], "unit": "tokens/s", "steps": res["steps"], "ms_per_step": res["ms_per_step"],
            "dtype": res["dtype"], "path": "streaming persistent kernel" if res["streaming"] else "per-op kernels + CUDA graph",
            "e2e": {"value": res["e2e_value"], "unit": "tokens/s", "h2d_bytes_per_step": 8 * res["batch"], "d2h_bytes_per_step": 4 * res["batch"]},
            "gpu_launches": res["launches"], "launches_per_step": res["launches_per_step"],
            "roofline": {k: r[k] for k in ("bound", "achieved", "peak", "unit", "frac", "traffic", "algorithmic_bytes_per_step")},
            "clocks": res["clocks"]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=256)
    ap.add_argument("--warmup", type=int, default=16)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--workload", default=None, help="default: 1b-bf16 on one GPU (configs[1]), 8b-bf16 tensor-parallel on several (configs[3])")
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-roofline", action="store_true", help="skip the isolated per-shape GEMV timing of the per-op path")
    ap.add_argument("--roofline-gemv", action="store_true", help="add the isolated per-shape timing of the per-op GEMV kernel")
    ap.add_argument("--per-op", action="store_true", help="per-op kernels under a CUDA graph instead of the streaming persistent kernel")
    ap.add_argument("--prompt", type=int, default=2048, help="prompt length of the *-prefill workloads")
    ap.add_argument("--tp", action="store_true", help="N > 1: ONE model sharded tensor-parallel over the N GPUs (the default for N > 1)")
    ap.add_argument("--tp-collective", choices=["fused", "nccl"], default="fused",
                    help="tensor parallel: the all-reduces of a block fused into the kernels over peer memory (default) or, as a comparator, ncclAllReduce calls between the per-op kernels")
    ap.add_argument("--replicas", action="store_true", help="N > 1: N independent replicas of the single-GPU workload (weak scaling) instead of one sharded model")
    ap.add_argument("--no-also", action="store_true", help="N = 1 default line: skip the additional workloads of the `also` list")
    args = ap.parse_args()
    args.warmup = max(3, args.warmup)
    rank, world, local = dist_env()
    tp_on = world > 1 and not args.replicas
    default_line = args.workload is None and args.batch == 1 and not args.per_op
    if args.workload is None:
        args.workload = "8b-bf16" if tp_on else "1b-bf16"
    if args.workload.endswith("-prefill"):
        shape_name, fmt = args.workload.split("-")[:2]
        run_prefill(args, shape_name, SHAPES[shape_name], quant=0 if fmt == "bf16" else 1)
        return
    shape_name, fmt = args.workload.split("-")
    shape = SHAPES[shape_name]
    quant = 0 if fmt == "bf16" else 1
    workload = workload_name(shape_name, fmt, args.batch, quant)

    if args.impl == "reference":
        run_reference(args, shape, quant, workload)
        return

    import torch

    dist = init_dist(local) if world > 1 else None

    from metalchat_b200 import capi

    dev = capi.Device(local)
    base = None
    if tp_on and default_line and rank == 0:
        # strong-scaling base: the same workload on ONE GPU (this rank's), measured in this run before the sharded model
        try:
            b = measure_decode(capi, torch, None, dev, shape_name, fmt, args.batch, min(args.steps, 64), args.warmup, 0, 1, local, e2e=False)
            base = {"n_gpus": 1, "value": b["value"], "unit": "tokens/s", "ms_per_step": b["ms_per_step"], "roofline_frac": b["roofline"]["frac"],
                    "how": "same workload on one GPU (rank 0's), measured in this run before the sharded model"}
        except Exception as e:  # e.g. the model does not fit one GPU
            base = {"n_gpus": 1, "unavailable": str(e)[:200]}
    res = measure_decode(capi, torch, dist, dev, shape_name, fmt, args.batch, args.steps, args.warmup, rank, world, local, tp_on=tp_on, per_op=args.per_op,
                         collective=args.tp_collective if tp_on else "fused")
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    roof = res["roofline"]
    if args.roofline_gemv and quant == 0:
        hbm_peak, _ = peaks()
        per, fam = measure_gemv_family(capi, dev, shape, hbm_peak)
        roof.update({"per_op_gemv_family": {"achieved": fam, "frac": fam / hbm_peak, "per_shape": per,
                                            "note": "isolated CUDA-event timing of the per-op GEMV kernel, weights cycled through a >L2 ring"}})
    B = args.batch
    line = {
        "metric": METRIC, "value": res["value"], "unit": "tokens/s", "n_gpus": world, "steps": res["steps"], "warmup": args.warmup,
        "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": "strong" if tp_on else "weak", "vs_baseline": None,
        "dtype": res["dtype"],
        "data": "synthetic ids, random-init weights (counter-hash seed 0x5EED)",
        "config": decode_config(args, workload, world, tp_on),
        "path": "streaming persistent kernel" if res["streaming"] else "per-op kernels + CUDA graph",
        "weights_streamed_per_step_mb": res["streamed"] / 1e6,
        "e2e": {"value": res["e2e_value"], "unit": "tokens/s", "h2d_bytes_per_step": 8 * B, "d2h_bytes_per_step": 4 * B,
                "ms_per_step": res["e2e_ms_per_step"]},
        "gpu_launches": res["launches"], "launches_per_step": res["launches_per_step"],
        "clocks": res["clocks"], "roofline": roof,
    }
    if base is not None:
        line["strong_scaling_base"] = base
    if default_line and world == 1 and not args.no_also:
        # the other single-GPU workloads of BASELINE.json's metric, measured in the same run (bounded: ~64 steps each)
        also = []
        k = min(args.steps, 64)
        for sn, f, b in (("1b", "w4", 1), ("1b", "bf16", 32), ("8b", "bf16", 1)):
            try:
                also.append(brief(measure_decode(capi, torch, None, dev, sn, f, b, k, args.warmup, 0, 1, local)))
            except Exception as e:
                also.append({"workload": workload_name(sn, f, b, 0 if f == "bf16" else 1), "unavailable": str(e)[:300]})
        try:
            also.append(measure_prefill_brief(capi, torch, dev, "1b", 0, 2048, 8, 3))
        except Exception as e:
            also.append({"workload": "llama-1b bf16 prefill of 2048 positions", "unavailable": str(e)[:300]})
        line["also"] = also
    if not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(shape, quant, 24 if shape_name == "1b" else 4)
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
