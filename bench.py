#!/usr/bin/env python
"""bench.py — PQ decode-attention throughput on Llama-3.1-8B shapes (BASELINE.json configs[1]).

A "step" = one decoded token: the PQ decode-attention hot path run for all 32 layers over a batch of sequences whose
32K-token KV caches are resident as PQ codes (M=64, C=256: the reference's "4-bit") plus a 128-token fp16 window.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--bs B] [--ctx T]

Prints ONE JSON line (see the task contract).  `value` = tokens/s with everything resident in HBM (CUDA-graph replay of
the 32 launches, CUDA events); `e2e` = tokens/s through the reference-facing cache API (DynamicPQCache.decoding) with
HOST inputs/outputs copied inside the timed region; `roofline` = algorithmic bytes / measured kernel time against the
measured HBM peak; `cpu_baseline` = the C oracle (OpenMP, all host cores) on a bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

LAYERS, NH, NH_K, D, M, C, LT = 32, 32, 8, 128, 64, 256, 128


def algorithmic_bytes(bs, nk, r, nh=NH, nh_k=NH_K, d=D, m=M, c=C):
    """codes once per KV head + fp16 window + both codebooks + q/o."""
    return 2 * bs * nh_k * nk * m + 2 * bs * nh_k * r * d * 2 + 2 * m * c * (d // m) * 2 + 2 * bs * nh * d * 2


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        j = json.load(open(p))
        return float(j["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = max(mx, float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ CPU arms


def cpu_sample_inputs(nk, bs=1):
    import numpy as np
    from oracle import pq_oracle as O
    return O.make_inputs(bs=bs, nh=NH, nh_k=NH_K, nk=nk, d=D, M=M, C=C, Lt=LT, seed=42)


def cpu_baseline_port(nk, r):
    """C oracle (oracle/pq_oracle.c, OpenMP) on ONE layer, bs=1; scaled to 32 layers."""
    from oracle import c_oracle as CO
    inp = cpu_sample_inputs(nk)
    f32 = {k: (v.astype("float32") if v.dtype != "uint8" else v) for k, v in inp.items()}
    CO.pq_decode_attn(f32["q"], f32["kc"], f32["vc"], f32["kcent"], f32["vcent"], f32["kres"], f32["vres"], r)
    t0, n = time.perf_counter(), 0
    while n < 3 or time.perf_counter() - t0 < 10.0:
        CO.pq_decode_attn(f32["q"], f32["kc"], f32["vc"], f32["kcent"], f32["vcent"], f32["kres"], f32["vres"], r)
        n += 1
    per_layer = (time.perf_counter() - t0) / n
    return {"value": 1.0 / (per_layer * LAYERS), "unit": "tokens/s", "cores": CO.num_threads(), "kind": "port",
            "sample": f"1 of {LAYERS} layers, bs 1, ctx {nk + r}: {n} calls of oracle/pq_oracle.c, {per_layer*1e3:.2f} ms each, scaled x{LAYERS} layers"}


def reference_arm(args):
    """The reference's own CPU path for this op, restated with the same torch calls it makes: sa_decode_4d-style
    gather of both code caches (pq_utils.py:501-540) + concat with the fp16 window + F.scaled_dot_product_attention
    without mask (pq_utils.py:360-368), fp32 on all host threads.  One step = one layer of the 32-layer token on a
    bounded batch (bs 1); tokens/s scaled to 32 layers."""
    import numpy as np
    import torch
    torch.set_num_threads(os.cpu_count() or 1)
    nk, r = args.ctx - LT, LT
    inp = cpu_sample_inputs(nk)
    q = torch.from_numpy(inp["q"]).float()
    kc, vc = torch.from_numpy(inp["kc"]).long(), torch.from_numpy(inp["vc"]).long()
    kcent, vcent = torch.from_numpy(inp["kcent"]).float(), torch.from_numpy(inp["vcent"]).float()
    kres, vres = torch.from_numpy(inp["kres"]).float(), torch.from_numpy(inp["vres"]).float()
    ar = torch.arange(M)

    def layer():
        K = torch.cat([kcent[ar, kc].reshape(1, NH_K, nk, D), kres[:, :, :r]], 2).repeat_interleave(NH // NH_K, 1)
        V = torch.cat([vcent[ar, vc].reshape(1, NH_K, nk, D), vres[:, :, :r]], 2).repeat_interleave(NH // NH_K, 1)
        return torch.nn.functional.scaled_dot_product_attention(q, K, V)

    for _ in range(args.warmup):
        layer()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        layer()
    per_layer = (time.perf_counter() - t0) / args.steps
    val = 1.0 / (per_layer * LAYERS)
    line = {"metric": "pq_decode_attn_tokens_per_s", "value": val, "unit": "tokens/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": per_layer * LAYERS * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "impl": "reference",
            "config": {"workload": f"llama31-8b-shapes_pq4bit_decode_attn_ctx{args.ctx}", "bs": 1, "layers": LAYERS},
            "cpu_baseline": {"value": val, "unit": "tokens/s", "cores": torch.get_num_threads(), "kind": "port",
                             "sample": f"each step = 1 of {LAYERS} layers at bs 1 (torch CPU gather + SDPA fp32), scaled x{LAYERS}"},
            "e2e": {"value": val, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------ GPU arm


def build_layers(torch, bs, nk, device, n_layers):
    g = torch.Generator(device=device)
    g.manual_seed(42)
    kcent = torch.randn(M, C, D // M, device=device, generator=g).half()
    vcent = torch.randn(M, C, D // M, device=device, generator=g).half()
    layers = []
    for _ in range(n_layers):
        layers.append(dict(
            kc=torch.randint(0, C, (bs, NH_K, nk, M), dtype=torch.uint8, device=device, generator=g),
            vc=torch.randint(0, C, (bs, NH_K, nk, M), dtype=torch.uint8, device=device, generator=g),
            kres=torch.randn(bs, NH_K, LT, D, device=device, generator=g).half(),
            vres=torch.randn(bs, NH_K, LT, D, device=device, generator=g).half(),
            q=torch.randn(bs, NH, 1, D, device=device, generator=g).half(),
            out=torch.empty(bs, NH, 1, D, device=device, dtype=torch.float16)))
    return kcent, vcent, layers


def time_resident(torch, ops, bs, nk, r, steps, warmup, device, impl=0, k_out=0):
    """Graph-replayed 32-layer step, inputs resident.  Returns (seconds for `steps` steps, launches).
    k_out > 0: with a K-side outlier store of that many (dim, delta) records per coded token (extension)."""
    kcent, vcent, layers = build_layers(torch, bs, nk, device, LAYERS)
    ws = ops.attn_workspace(device, bs, NH, NH_K, D, ops.default_splits(bs, NH_K, nk))
    for L_ in layers:
        L_["ko"] = (torch.randint(0, D, (bs, NH_K, nk, k_out), dtype=torch.uint8, device=device),
                    torch.randn(bs, NH_K, nk, k_out, device=device).half()) if k_out else None

    def step():
        for L_ in layers:
            ops.pq_decode_attn(L_["q"], L_["kc"], L_["vc"], kcent, vcent, L_["kres"], L_["vres"], r, out=L_["out"], workspace=ws, impl=impl,
                               k_outliers=L_["ko"])

    s = torch.cuda.Stream(device=device)
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        step()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        step()
    for _ in range(warmup):
        graph.replay()
    return graph, layers, (kcent, vcent)


def encode_rate(torch, ops, device):
    """PQ encode throughput on Llama-3.1-8B prefill shapes: K of one layer = 8 kv-heads x 32768 tokens, for the encoder AUTO
    picks (the exact candidate-grid encoder for d/M = 2) and for the tcgen05 distance encoder."""
    from million_b200 import _lib as L
    n = 32768
    X = torch.randn(1, NH_K, n, D, device=device).half()
    cent = torch.randn(M, C, D // M, device=device).half().float().contiguous()
    codes = torch.empty(1, NH_K, n, M, dtype=torch.uint8, device=device)
    out = {}
    for name, impl in (("auto_grid", L.IMPL_AUTO), ("tcgen05", L.IMPL_FAST)):
        for _ in range(3):
            ops.pq_encode_into(X, cent, codes, impl=impl)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(10):
            ops.pq_encode_into(X, cent, codes, impl=impl)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        vec_per_s = NH_K * n / (ms * 1e-3)
        out[name] = {"M_head_vectors_per_s": vec_per_s / 1e6, "Mtok_per_s_32_layers_K_and_V": vec_per_s / (2 * NH_K * LAYERS) / 1e6,
                     "brute_force_equivalent_TFLOPs": vec_per_s * 2 * D * C / 1e12, "ms_per_layer_K_32k": ms,
                     "hbm_GBps_input_plus_codes": vec_per_s * (2 * D + M) / 1e9}
    return out


def ncu_traffic(bs):
    """dram bytes per launch from the committed ncu capture of this kernel (profiles/r01_traffic.json), or None."""
    p = os.path.join(ROOT, "profiles", "r01_traffic.json")
    if os.path.exists(p):
        j = json.load(open(p)).get(f"attn_fast_bs{bs}")
        if j:
            return j["dram_bytes_read"] + j["dram_bytes_write"]
    return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--bs", type=int, default=8)
    ap.add_argument("--ctx", type=int, default=32768)
    ap.add_argument("--kernel", type=int, default=0, help="0 auto, 1 generic, 2 fast")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        if rank == 0:
            reference_arm(args)
        return

    import torch
    import torch.distributed as dist
    from million_b200 import _lib, ops
    from million_b200.pq_utils import DynamicPQCache, Singleton
    _lib.lib()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a GPU (there is no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL prints its version banner on stdout when the first communicator is created; rank 0 must print ONE JSON line,
        # so file descriptor 1 points at stderr until the communicator exists
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=device)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)

    bs, nk, r = args.bs, args.ctx - LT, LT
    steps, warmup = args.steps, max(args.warmup, 3)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- resident-inputs number (value) + roofline
    graph, layers, cents = time_resident(torch, ops, bs, nk, r, steps, warmup, device, impl=args.kernel)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    with ClockSampler(local) as clk:
        e0.record()
        for _ in range(steps):
            graph.replay()
        e1.record()
        barrier()
    secs = e0.elapsed_time(e1) / 1e3
    if world > 1:
        t = torch.tensor([secs], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        secs = float(t.item())
    launches = LAYERS * steps
    per_launch = secs / launches
    alg = algorithmic_bytes(bs, nk, r)
    peak, peak_src = measured_peaks()
    achieved = alg / per_launch / 1e9
    value = world * bs * steps / secs

    # ---- batch-1 latency shape of the same config (reported alongside, not the headline)
    extra = {}
    if rank == 0 and bs != 1:
        g1, l1, c1 = time_resident(torch, ops, 1, nk, r, steps, warmup, device, impl=args.kernel)
        torch.cuda.synchronize(); e0.record()
        for _ in range(steps):
            g1.replay()
        e1.record(); torch.cuda.synchronize()
        s1 = e0.elapsed_time(e1) / 1e3
        a1 = algorithmic_bytes(1, nk, r) / (s1 / launches) / 1e9
        extra["bs1"] = {"tokens_per_s": steps / s1, "us_per_layer": s1 / launches * 1e6, "achieved_GBps": a1, "frac": a1 / peak}
        del g1, l1, c1
        # the same step with a K-side outlier store (2 records per coded token and KV head: +6 bytes on 128)
        g2, l2, c2 = time_resident(torch, ops, bs, nk, r, steps, warmup, device, impl=args.kernel, k_out=2)
        torch.cuda.synchronize(); e0.record()
        for _ in range(steps):
            g2.replay()
        e1.record(); torch.cuda.synchronize()
        s2 = e0.elapsed_time(e1) / 1e3
        alg2 = alg + 3 * 2 * bs * NH_K * nk
        extra["k_outliers_2"] = {"tokens_per_s": bs * steps / s2, "us_per_layer": s2 / launches * 1e6,
                                 "achieved_GBps": alg2 / (s2 / launches) / 1e9, "frac": alg2 / (s2 / launches) / 1e9 / peak}
        del g2, l2, c2
        try:
            extra["encode"] = encode_rate(torch, ops, device)
        except Exception as e:
            extra["encode"] = {"error": repr(e)}
    barrier()

    # ---- end-to-end through the reference-facing API with HOST buffers
    del graph
    Singleton.clear_instance()
    cache = DynamicPQCache(bs=bs, nh=NH, num_key_value_heads=NH_K, M=M, layer_num=LAYERS, d=D, scalar_t=torch.float16, device=device)
    cache.set_cent(*cents)
    for li, L_ in enumerate(layers):              # adopt the resident code caches as the prefilled state
        cache._k[li].buf, cache._k[li].cap, cache._k[li].len = L_["kc"], nk, nk
        cache._v[li].buf, cache._v[li].cap, cache._v[li].len = L_["vc"], nk, nk
        cache.seen_tokens[li] = nk
    e2e_steps = min(steps, LT - 1)
    hq = torch.randn(LAYERS, bs, NH, 1, D).half().pin_memory()
    hk = torch.randn(LAYERS, bs, NH_K, 1, D).half().pin_memory()
    hv = torch.randn(LAYERS, bs, NH_K, 1, D).half().pin_memory()
    ho = torch.empty(LAYERS, bs, NH, 1, D, dtype=torch.float16).pin_memory()
    dq, dk, dv = (torch.empty_like(x, device=device) for x in (hq, hk, hv))
    do = torch.empty(LAYERS, bs, NH, 1, D, dtype=torch.float16, device=device)

    def e2e_step():
        dq.copy_(hq, non_blocking=True); dk.copy_(hk, non_blocking=True); dv.copy_(hv, non_blocking=True)
        for li in range(LAYERS):
            do[li] = cache.decoding(dq[li], dk[li], dv[li], li)
        ho.copy_(do, non_blocking=True)

    for _ in range(3):
        e2e_step()
    barrier()
    e0.record()
    for _ in range(e2e_steps - 3):
        e2e_step()
    e1.record()
    barrier()
    e2e_secs = e0.elapsed_time(e1) / 1e3
    if world > 1:
        t = torch.tensor([e2e_secs], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_secs = float(t.item())
    e2e_val = world * bs * (e2e_steps - 3) / e2e_secs
    h2d = hq.numel() * 2 + hk.numel() * 2 + hv.numel() * 2
    d2h = ho.numel() * 2

    if rank == 0:
        line = {
            "metric": "pq_decode_attn_tokens_per_s", "value": value, "unit": "tokens/s", "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": secs / steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f16",
            "data": "synthetic",
            "config": {"workload": f"llama31-8b-shapes_pq4bit_decode_attn_ctx{args.ctx}", "bs_per_gpu": bs, "layers": LAYERS, "nh": NH, "nh_k": NH_K,
                       "d": D, "M": M, "C": C, "window": r, "sharding": "independent sequences per rank (no collective)" if world > 1 else "single GPU",
                       "l2": f"inputs larger than L2: {LAYERS} layer caches, {LAYERS * alg / 1e9:.2f} GB touched per step", "kernel": args.kernel},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": ncu_traffic(bs),
                         "peak_source": peak_src, "algorithmic_bytes_per_launch": alg, "us_per_launch": per_launch * 1e6},
            "e2e": {"value": e2e_val, "unit": "tokens/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "api": "DynamicPQCache.decoding x 32 layers, pinned host q/k/v in, host out"},
            "gpu_launches": launches, "clocks": clk.summary(), "extra": extra,
        }
        if world == 1 and not args.no_cpu_baseline:
            try:
                line["cpu_baseline"] = cpu_baseline_port(nk, r)
            except Exception as e:  # the oracle is test infrastructure; its absence must not hide the GPU number
                line["cpu_baseline"] = {"error": repr(e)}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
