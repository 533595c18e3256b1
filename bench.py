#!/usr/bin/env python
"""Benchmark of the PointNet++ set-abstraction path (BASELINE.json metric: SSG forward clouds/sec at N=1024).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B] [--precision bf16|fp32]

One "step" = one forward of pointnet2_cls_ssg (configs[0]'s architecture and input recipe: seeded uniform
clouds, per-cloud normalised, [B,3,1024]) over one batch of B clouds per GPU.  `value` times K steps with the
inputs resident in HBM (CUDA events, barrier + synchronize on both sides, max over ranks); `e2e` repeats them
through the public module call with pinned HOST inputs and a host read of the result inside the timed region.
`--impl reference` times the CPU port of the reference path (oracle/pn2_torch_port.py, all host threads) on a
bounded sample of the same workload.  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "pointnet-like-pose-estimation_b200")
for p in (PKG, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

N_POINTS = 1024
MODEL = "pointnet2_cls_ssg"
CPU_BATCH = 24          # configs[0]: the reference's own CPU-runnable case
N_ROTATING = 4

# dense algorithmic FLOPs per cloud of the shared MLPs, unpadded Cin (SURVEY.md 8a / DESIGN.md)
SA_SPECS = {"pointnet2_cls_ssg": [(512, 32, [3, 64, 64, 128]), (128, 64, [131, 128, 128, 256]), (1, 128, [259, 256, 512, 1024])]}


def mlp_flops(S, K, dims):
    return 2.0 * S * K * sum(a * b for a, b in zip(dims[:-1], dims[1:]))


def model_flops_per_cloud(name):
    return sum(mlp_flops(*s) for s in SA_SPECS[name]) + 2.0 * (1024 * 512 + 512 * 256 + 256 * 7)


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return dict(hbm=d["hbm_gbs"], bf16_burst=d["bf16_tflops"], bf16_sustained=d["bf16_tflops_sustained"], source="measured")
    return dict(hbm=6650.0, bf16_burst=1590.0, bf16_sustained=1400.0, source="fallback")


def make_inputs(B, seed):
    import cloudgen
    return cloudgen.channels_first(cloudgen.uniform_cloud(B, N_POINTS, seed))      # [B,3,N] fp32, CPU


def build_model(sa_impl=None):
    import pn2_models
    torch.manual_seed(0)
    m = pn2_models.get_model(MODEL, 7, sa_impl=sa_impl)
    pn2_models.randomize_bn_(m)
    return m.eval()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed regions run."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            pass

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            out = self.p.communicate(timeout=5)[0]
        except subprocess.TimeoutExpired:
            self.p.kill()
            out = self.p.communicate()[0]
        sm, mx, pw, reasons = [], [], [], set()
        for line in out.splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        busy = [s for s, p in zip(sm, pw) if p >= 0.5 * max(pw)] or sm      # samples taken under load
        return {"sm_mhz": statistics.median(busy), "sm_max_mhz": max(mx), "power_w_max": max(pw), "samples": len(sm),
                "reasons": sorted(reasons)}


def cpu_port_throughput(budget_s, threads):
    """clouds/s of the CPU port of the reference path at configs[0]'s batch (B=24), bounded to ~budget_s."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pn2_torch_port as port          # bench.py's CPU legs are the one place outside tests/ that may use oracle/
    torch.set_num_threads(threads)
    model = build_model(sa_impl=port)
    x = make_inputs(CPU_BATCH, 1001)
    times = []
    with torch.no_grad():
        torch.manual_seed(42)
        model(x)                            # warm-up
        t_end = time.perf_counter() + budget_s
        while len(times) < 3 or (time.perf_counter() < t_end and len(times) < 10):
            torch.manual_seed(42)
            t0 = time.perf_counter()
            model(x)
            times.append(time.perf_counter() - t0)
    return CPU_BATCH / statistics.median(times), len(times)


def workload_config(B, world, precision):
    """The `config` object of the JSON line -- the same workload for both arms (the reference arm adds its bounded sample)."""
    return {"workload": f"{MODEL} forward, N={N_POINTS} xyz (configs[0] architecture and input recipe), B={B} clouds per GPU per step",
            "batch_per_gpu": B, "global_batch": B * world, "parallelism": f"batch-sharded x{world}, no collective",
            "l2": f"{N_ROTATING} rotating input batches; per-step working set (l1 features {B * 512 * 128 * 4 / 1e6:.0f} MB) exceeds the 126 MB L2",
            "precision_mode": precision}


def run_reference(args, rank):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pn2_torch_port as port
    torch.set_num_threads(threads)
    model = build_model(sa_impl=port)
    x = make_inputs(CPU_BATCH, 1001)
    with torch.no_grad():
        for _ in range(max(1, min(args.warmup, 2))):
            torch.manual_seed(42)
            model(x)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            torch.manual_seed(42)
            model(x)
        dt = time.perf_counter() - t0
    value = CPU_BATCH * args.steps / dt
    sample = f"{MODEL} forward on {CPU_BATCH} clouds (configs[0] batch) per step, {args.steps} steps"
    print(json.dumps({
        "impl": "reference", "metric": "PointNet++ SSG fwd clouds/sec (N=1024)", "value": value, "unit": "clouds/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": dict(workload_config(args.batch, 1, "fp32"), sample=sample + "; CPU port of the reference path (ATen/MKL/oneDNN), "
                       f"{threads} host threads, no GPU"),
        "cpu_baseline": {"value": value, "unit": "clouds/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "clouds/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


def summarise_trace(trace, steps, pk):
    """Per-kernel device time from the CUDA events recorded around every C-ABI launch of the timed region."""
    per = {}
    for name, meta, s, e in trace:
        key = name
        if name == "sa_fused":
            key = f"sa_fused[S={meta['S']},K={meta['K']},{'x'.join(map(str, meta['dims']))}]"
        elif name in ("fps", "ball_query"):
            key = f"{name}[N={meta['N']},S={meta['S']}" + (f",K={meta['K']}]" if name == "ball_query" else "]")
        d = per.setdefault(key, {"ms": 0.0, "launches": 0, "name": name, "meta": meta})
        d["ms"] += s.elapsed_time(e)
        d["launches"] += 1
    total = sum(d["ms"] for d in per.values()) or 1.0
    table = {k: {"ms_per_launch": d["ms"] / d["launches"], "launches_per_step": d["launches"] / steps, "share": d["ms"] / total}
             for k, d in per.items()}
    top_key = max(per, key=lambda k: per[k]["ms"])
    top = per[top_key]
    m = top["meta"]
    dur_s = 1e-3 * top["ms"] / top["launches"]
    if top["name"] == "sa_fused":
        flops = m["B"] * mlp_flops(m["S"], m["K"], list(m["dims"]))
        peak = pk["bf16_sustained"] if m["prec"] == 1 else None
        roof = {"kernel": top_key, "bound": "tensor", "achieved": flops / dur_s / 1e12, "peak": peak, "unit": "TFLOP/s",
                "frac": (flops / dur_s / 1e12 / peak) if peak else None, "traffic": None,
                "peak_source": f"{pk['source']} bf16 sustained (kernel timed inside a long step)", "algorithmic_flops_per_launch": flops}
    else:
        if top["name"] == "fps":
            nbytes = m["B"] * (4 * m["N"] * m["C"] + 4 * m["S"] * m["C"] + 8)
        else:
            nbytes = m["B"] * (4 * m["N"] * m["C"] + 4 * m["S"] * m["C"] + (8 if m.get("i64") else 4) * m["S"] * m.get("K", 1))
        roof = {"kernel": top_key, "bound": "hbm", "achieved": nbytes / dur_s / 1e9, "peak": pk["hbm"], "unit": "GB/s",
                "frac": nbytes / dur_s / 1e9 / pk["hbm"], "traffic": None, "peak_source": f"{pk['source']} HBM copy",
                "algorithmic_bytes_per_launch": nbytes}
    traffic_file = os.path.join(ROOT, "profiles", "dram_traffic.json")
    if os.path.exists(traffic_file):
        with open(traffic_file) as f:
            t = json.load(f).get(top_key)
        if t and t.get("batch") == m["B"]:
            roof["traffic"] = t["dram_bytes_per_launch"]
    return table, roof


def device_time_us(fn, reps=30, warm=5, flush=None):
    """Median device time of one call of `fn` in us: CUDA events around every call, NO host synchronisation between the calls
    (the host enqueues ahead of the GPU, so the bracket holds the kernel and not the Python / ctypes launch latency -- with a
    synchronize per call a 100 us kernel reads 40 us longer), an L2 flush (`flush`: a buffer larger than the 126 MB L2,
    rewritten outside the brackets) before every call."""
    evs = []
    for i in range(warm + reps):
        if flush is not None:
            flush.add_(1)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        evs.append((s, e))
    torch.cuda.synchronize()
    return statistics.median(1e3 * s.elapsed_time(e) for s, e in evs[warm:])


def micro_cfg3(dev, be):
    """configs[2]: FPS + ball query at B=16, N=8192 -> 2048 centroids, r=0.2, nsample=32 (median of 30 launches, us;
    device time, L2 flushed before every launch)."""
    import cloudgen
    import pointnet2_utils as pu
    x = cloudgen.uniform_cloud(16, 8192, 1003).to(dev)
    start = torch.randint(0, 8192, (16,), generator=torch.Generator().manual_seed(42), dtype=torch.long).to(dev)
    r2 = pu._radius_sq(0.2)
    flush = torch.zeros(64 * 1024 * 1024, dtype=torch.int32, device=dev)          # 256 MB
    _, _, new_pts = be.fps(x, start, 2048, want_idx64=True, want_new_points=True)
    res = {"fps_us": device_time_us(lambda: be.fps(x, start, 2048, want_idx64=True, want_new_points=True), flush=flush),
           "ball_query_us": device_time_us(lambda: be.ball_query(new_pts, x, r2, 32, want_idx64=True), flush=flush)}
    del flush
    return res


def run_ours(args, rank, local_rank, world):
    import pointnet2_utils as pu
    dev = torch.device(f"cuda:{local_rank}")
    torch.cuda.set_device(dev)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        dist.init_process_group("nccl", device_id=dev)
    pk = peaks()
    pu.set_precision(args.precision)
    be = pu._backend()
    model = build_model().to(dev)
    B, K, W = args.batch, args.steps, args.warmup
    host = [make_inputs(B, 1001 + 97 * rank + i).pin_memory() for i in range(N_ROTATING)]
    resident = [h.to(dev) for h in host]
    gen = torch.Generator().manual_seed(42 + rank)
    starts = [[torch.randint(0, n, (B,), generator=gen, dtype=torch.long).to(dev) for n in (1024, 512)] for _ in range(N_ROTATING)]
    out_host = torch.empty(B, 7, dtype=torch.float32).pin_memory()
    pred_host = torch.empty(B, dtype=torch.int64).pin_memory()

    class Starts:
        """FPS start indices are part of the resident input set of a step (drawn once, seeded, uploaded)."""
        def __init__(self):
            self.step, self.call = 0, 0

        def __call__(self, B_, N_):
            t = starts[self.step % N_ROTATING][self.call % 2]
            self.call += 1
            return t

    src = Starts()

    def step_resident(i):
        src.step, src.call = i, 0
        return model(resident[i % N_ROTATING])

    # e2e pipeline: the H2D copy of step i+1 runs on a copy stream underneath the forward of step i (two device
    # staging buffers); the result of every step is read back to pinned host memory on the compute stream
    copy_stream = torch.cuda.Stream(device=dev)
    stage = [torch.empty_like(resident[0]) for _ in range(2)]
    staged = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]

    def e2e_upload(i):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[i % 2])              # the forward that read this buffer two steps ago is done
            stage[i % 2].copy_(host[i % N_ROTATING], non_blocking=True)
            staged[i % 2].record(copy_stream)

    def step_e2e(i, last):
        src.step, src.call = i, 0
        if not last:
            e2e_upload(i + 1)
        torch.cuda.current_stream().wait_event(staged[i % 2])
        logp, _, pred = model(stage[i % 2])
        consumed[i % 2].record()
        out_host.copy_(logp, non_blocking=True)
        pred_host.copy_(pred, non_blocking=True)

    def run_e2e(n):
        for ev in consumed:
            ev.record()
        e2e_upload(0)
        for i in range(n):
            step_e2e(i, i == n - 1)

    def barrier():
        if dist is not None:
            dist.barrier(device_ids=[local_rank])
        torch.cuda.synchronize(dev)

    def reduce_max(ms):
        if dist is None:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    clocks = ClockSampler(local_rank) if rank == 0 else None
    with torch.no_grad(), pu.fps_start_source(src):
        for i in range(max(W, 3)):
            step_resident(i)
        # ---- value: K steps, inputs resident in HBM ---------------------------------------------------------
        barrier()
        be.trace = []
        l0 = be.launches
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for i in range(K):
            step_resident(i)
        ev1.record()
        barrier()
        launches = be.launches - l0
        trace, be.trace = be.trace, None
        ms = reduce_max(ev0.elapsed_time(ev1))
        # ---- e2e: host buffers in, host result out, copies inside the timed region --------------------------------
        run_e2e(2)
        barrier()
        ev0.record()
        run_e2e(K)                       # K uploads, K forwards, K read-backs inside the timed region
        ev1.record()
        barrier()
        ms_e2e = reduce_max(ev0.elapsed_time(ev1))
        # ---- configs[0] literal batch (B=24), latency-bound; and the fp32 mode ---------------------------------
        extra = {}
        if rank == 0 and not args.quick:
            xb = resident[0][:CPU_BATCH].contiguous()
            sb = [s[:CPU_BATCH].contiguous() for s in starts[0]]

            def small():
                it = iter(sb)
                with pu.fps_start_source(lambda B_, N_: next(it)):
                    model(xb)
            for _ in range(5):
                small()
            torch.cuda.synchronize(dev)
            ev0.record()
            for _ in range(20):
                small()
            ev1.record()
            torch.cuda.synchronize(dev)
            extra["b24_clouds_per_s"] = CPU_BATCH * 20 / (1e-3 * ev0.elapsed_time(ev1))
            # the same 24-cloud forward replayed as one CUDA graph (pn2_graph.GraphedForward: launch latency removed,
            # FPS start indices still drawn on the CPU generator per forward and uploaded by the graph's memcpy nodes)
            import pn2_graph
            g = pn2_graph.GraphedForward(model, xb)
            for _ in range(5):
                g(xb)
            torch.cuda.synchronize(dev)
            ev0.record()
            for _ in range(20):
                g(xb)
            ev1.record()
            torch.cuda.synchronize(dev)
            extra["b24_graph_clouds_per_s"] = CPU_BATCH * 20 / (1e-3 * ev0.elapsed_time(ev1))
            other = "fp32" if args.precision == "bf16" else "bf16"
            pu.set_precision(other)
            with pu.fps_start_source(src):
                for i in range(3):
                    step_resident(i)
                torch.cuda.synchronize(dev)
                n_other = max(3, K // 4)
                ev0.record()
                for i in range(n_other):
                    step_resident(i)
                ev1.record()
                torch.cuda.synchronize(dev)
            extra[f"{other}_mode_clouds_per_s"] = B * n_other / (1e-3 * ev0.elapsed_time(ev1))
            pu.set_precision(args.precision)
            extra.update(micro_cfg3(dev, be))
    clk = clocks.stop() if clocks else None
    if rank != 0:
        if dist is not None:
            dist.barrier(device_ids=[local_rank])
            dist.destroy_process_group()
        return
    kernels, roof = summarise_trace(trace, K, pk)
    cpu = None
    if world == 1 and not args.no_cpu_baseline and not args.quick:
        threads = os.cpu_count() or 1
        v, n = cpu_port_throughput(15.0, threads)
        cpu = {"value": v, "unit": "clouds/s", "cores": threads, "kind": "port",
               "sample": f"{MODEL} forward on {CPU_BATCH} clouds (configs[0] batch), median of {n} forwards, torch CPU port of the reference path"}
    value = B * world * K / (1e-3 * ms)
    flops = model_flops_per_cloud(MODEL)
    line = {
        "metric": "PointNet++ SSG fwd clouds/sec (N=1024)", "value": value, "unit": "clouds/s", "n_gpus": world, "steps": K, "warmup": max(W, 3),
        "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
        "config": workload_config(B, world, args.precision),
        "e2e": {"value": B * world * K / (1e-3 * ms_e2e), "unit": "clouds/s", "h2d_bytes_per_step": B * 3 * N_POINTS * 4,
                "d2h_bytes_per_step": B * 7 * 4 + B * 8},
        "gpu_launches": launches,
        "clocks": clk,
        "roofline": roof,
        "model_tflops": value * flops / 1e12,
        "model_tensor_frac": value / world * flops / 1e12 / pk["bf16_sustained"] if args.precision == "bf16" else None,
        "kernels": kernels,
        "cpu_baseline": cpu,
    }
    line.update(extra)
    print(json.dumps(line))
    if dist is not None:
        dist.barrier(device_ids=[local_rank])
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--batch", type=int, default=2048, help="clouds per GPU per step")
    ap.add_argument("--precision", choices=["bf16", "fp32"], default="bf16")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--quick", action="store_true", help="timed regions only (for runs under a profiler)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU path (use --impl reference for the CPU baseline)")
    run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
