#!/usr/bin/env python
"""Benchmark of the PointNet++ set-abstraction path (BASELINE.json: "PointNet++ SSG fwd clouds/sec (N=1024) @1/2/4/8 B200;
FPS+ball-query us").

    python bench.py [--config 0..4] [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B] [--precision bf16|fp32]

`--config` selects one of BASELINE.json's configs (default 0 = the headline):
  0  pointnet2_cls_ssg forward, N = 1024            (clouds/s; B = 2048 clouds per GPU per step, weak scaling)
  1  pointnet2_cls_msg forward, N = 1024            (clouds/s; B = 512 per GPU, weak scaling; three-radius MSG grouping)
  2  FPS + ball query, B = 16, N = 8192 -> 2048, r = 0.2, K = 32     (us per call; N GPUs = N replicas)
  3  full pose pipeline (6 PointNet++ heads + the v1 width head), B = 1024 split over the GPUs     (clouds/s, strong scaling)
  4  rotation_msg training step, N = 2048, 32 clouds per GPU, flat-bucket NCCL all-reduce          (clouds/s, weak scaling)

One "step" = one pass of the path over one batch.  `value` times K steps with the inputs resident in HBM (CUDA events, barrier +
synchronize on both sides, max over ranks); `e2e` repeats them through the public module call with pinned HOST inputs and a host
read of the result inside the timed region.  Both go through the product's default FPS start draw (torch.randint on the CPU
generator + upload, pointnet2_utils.py:59), nothing is pre-uploaded.  Forward configs report the headline precision mode AND a
full record (value, e2e, roofline) of the other one.  `--impl reference` times the CPU port of the reference path
(oracle/pn2_torch_port.py, all host threads) on a bounded sample of the same workload.  Prints ONE JSON line on rank 0.
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
for p in (PKG, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tools")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

N_ROTATING = 4
FORWARD = {
    0: dict(model="pointnet2_cls_ssg", points=1024, batch=2048, cpu_batch=24, seed=1001,
            metric="PointNet++ SSG fwd clouds/sec (N=1024)"),
    1: dict(model="pointnet2_cls_msg", points=1024, batch=512, cpu_batch=32, seed=1002,
            metric="PointNet++ MSG fwd clouds/sec (N=1024)"),
}
CFG2 = dict(B=16, N=8192, S=2048, r=0.2, K=32, seed=1003)


# ---------------------------------------------------------------------------------------------------------------------
# shared helpers
# ---------------------------------------------------------------------------------------------------------------------
def mlp_flops(S, K, dims):
    return 2.0 * S * K * sum(a * b for a, b in zip(dims[:-1], dims[1:]))


def model_flops_per_cloud(model, n_points):
    """Dense algorithmic FLOPs per cloud: every set-abstraction MLP with unpadded Cin and dense K + the FC head."""
    total, n = 0.0, n_points
    for i in range(1, getattr(model, "n_sa", 0) + 1):
        sa = getattr(model, f"sa{i}")
        if hasattr(sa, "conv_blocks"):
            for K, convs in zip(sa.sample_number_list, sa.conv_blocks):
                total += mlp_flops(sa.point_number, K, [convs[0].in_channels] + [c.out_channels for c in convs])
            n = sa.point_number
        else:
            dims = [sa.mlp_convs[0].in_channels] + [c.out_channels for c in sa.mlp_convs]
            S, K = (1, n) if sa.group_all else (sa.point_number, sa.sample_number)
            total += mlp_flops(S, K, dims)
            n = S
    for name in ("fc1", "fc2", "fc3"):
        fc = getattr(model, name, None)
        if fc is not None:
            total += 2.0 * fc.in_features * fc.out_features
    return total


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return dict(hbm=d["hbm_gbs"], bf16_burst=d["bf16_tflops"], bf16_sustained=d["bf16_tflops_sustained"], source="measured")
    return dict(hbm=6650.0, bf16_burst=1590.0, bf16_sustained=1400.0, source="fallback (B200_PROFILING.md)")


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


class Ranks:
    """One process per GPU: device, NCCL group (world > 1), barrier and max-over-ranks."""

    def __init__(self):
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.dist = None
        self.dev = None

    def init_cuda(self):
        self.dev = torch.device(f"cuda:{self.local}")
        torch.cuda.set_device(self.dev)
        if self.world > 1:
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=self.dev)
            self.dist = dist

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier(device_ids=[self.local])
        torch.cuda.synchronize(self.dev)

    def reduce_max(self, v):
        if self.dist is None:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def finish(self):
        if self.dist is not None:
            self.dist.barrier(device_ids=[self.local])
            self.dist.destroy_process_group()


def cpu_threads():
    return os.cpu_count() or 1


def port():
    """The CPU port of the reference path (bench.py's CPU legs are the one place outside tests/ that may use oracle/)."""
    p = os.path.join(ROOT, "oracle")
    if p not in sys.path:
        sys.path.insert(0, p)
    import pn2_torch_port
    return pn2_torch_port


def make_inputs(B, n_points, seed):
    import cloudgen
    return cloudgen.channels_first(cloudgen.uniform_cloud(B, n_points, seed))      # [B,3,N] fp32, CPU


def build_model(name, sa_impl=None):
    import pn2_models
    torch.manual_seed(0)
    m = pn2_models.get_model(name, 7, sa_impl=sa_impl)
    pn2_models.randomize_bn_(m)
    return m.eval()


def workload_config(cfg_id, B, world):
    """The `config` object of the JSON line -- IDENTICAL for both arms.  The reference arm times bounded samples of this workload
    (`reference_arm_sample`: the reference's own CPU-sized batch; clouds/s is a per-cloud rate); the product arm reports its
    throughput at that same batch next to the headline (`same_batch_as_reference_sample`)."""
    f = FORWARD[cfg_id]
    return {"workload": f"{f['model']} forward, N={f['points']} xyz (configs[{cfg_id}] architecture and input recipe), B={B} clouds per GPU per step",
            "batch_per_gpu": B, "global_batch": B * world, "parallelism": f"batch-sharded x{world}, no collective",
            "l2": f"{N_ROTATING} rotating input batches; the per-step working set (level-1 features {B * 512 * 4 * (128 if cfg_id == 0 else 320) / 1e6:.0f} MB) exceeds the 126 MB L2",
            "fps_start": "drawn per layer per step by the product (torch.randint on the CPU generator + upload), inside the timed region",
            "reference_arm_sample": f"{f['cpu_batch']}-cloud forwards (configs[{cfg_id}]'s own batch) of the same model and input recipe"}


def pick_peak(pk, region_s, clk):
    """Burst peak for a short, un-capped timed region; sustained when the region is long or a power cap was sampled."""
    capped = bool(clk) and "sw_power_cap" in (clk.get("reasons") or [])
    if region_s < 1.0 and not capped:
        return pk["bf16_burst"], f"{pk['source']} bf16 burst (timed region {region_s:.2f} s, no power cap sampled)"
    return pk["bf16_sustained"], f"{pk['source']} bf16 sustained (timed region {region_s:.2f} s{', sw_power_cap sampled' if capped else ''})"


def summarise_trace(trace, steps):
    """Per-kernel device time from the CUDA events recorded around every C-ABI launch of the timed region."""
    per = {}
    for name, meta, s, e in trace:
        key = name
        if name == "sa_fused":
            key = f"sa_fused[S={meta['S']},K={meta['K']},{'x'.join(map(str, meta.get('algo_dims') or meta['dims']))}]"
        elif name in ("fps", "ball_query"):
            key = f"{name}[N={meta['N']},S={meta['S']}" + (f",K={meta['K']}]" if name == "ball_query" else "]")
        elif name == "ball_query_multi":
            key = f"ball_query_multi[N={meta['N']},S={meta['S']},K={'/'.join(map(str, meta['K']))}]"
        elif name == "rows_gemm":
            key = f"rows_gemm[{meta['M']}x{meta['K']}x{meta['N']}]"
        d = per.setdefault(key, {"ms": 0.0, "launches": 0, "name": name, "meta": meta})
        d["ms"] += s.elapsed_time(e)
        d["launches"] += 1
    total = sum(d["ms"] for d in per.values()) or 1.0
    table = {k: {"ms_per_launch": d["ms"] / d["launches"], "launches_per_step": d["launches"] / steps, "share": d["ms"] / total}
             for k, d in per.items()}
    top_key = max(per, key=lambda k: per[k]["ms"])
    return table, top_key, per[top_key]


def roofline_of(top_key, top, prec, pk, region_s, clk, fp32_tc=False):
    m = top["meta"]
    dur_s = 1e-3 * top["ms"] / top["launches"]
    if top["name"] == "sa_fused":
        flops = m["B"] * mlp_flops(m["S"], m["K"], list(m.get("algo_dims") or m["dims"]))
        ach = flops / dur_s / 1e12
        roof = {"kernel": top_key, "bound": "tensor", "achieved": ach, "unit": "TFLOP/s", "traffic": None,
                "algorithmic_flops_per_launch": flops}
        if prec == "bf16":
            peak, src = pick_peak(pk, region_s, clk)
            roof.update(peak=peak, frac=ach / peak, peak_source=src,
                        frac_vs_burst=ach / pk["bf16_burst"], frac_vs_sustained=ach / pk["bf16_sustained"])
        elif fp32_tc:
            peak, src = pick_peak(pk, region_s, clk)
            roof.update(peak=peak / 4, frac=ach / (peak / 4), frac_vs_burst=ach / (pk["bf16_burst"] / 4), frac_vs_sustained=ach / (pk["bf16_sustained"] / 4),
                        peak_source=src + " / 4: fp32 mode runs every algorithmic MMA as four bf16 tensor-core passes on split (hi, lo) operands "
                                          "(pn2_sa_tcs.cu), so its ceiling is a quarter of the bf16 peak")
        else:
            sm = (clk or {}).get("sm_mhz") or 1965.0
            peak = 148 * 128 * 2 * sm * 1e6 / 1e12
            roof.update(bound="fp32", peak=peak, frac=ach / peak, peak_source=f"FP32 FMA pipe: 148 SMs x 128 lanes x 2 x {sm:.0f} MHz")
    elif top["name"] not in ("fps", "ball_query", "ball_query_multi"):
        # not a kernel of the path's roofline table (under a profiler the serialised launches can make a helper the longest span)
        return {"kernel": top_key, "bound": None, "achieved": None, "peak": None, "unit": None, "frac": None, "traffic": None}
    else:
        if top["name"] == "fps":
            nbytes = m["B"] * (4 * m["N"] * m["C"] + 4 * m["S"] * m["C"] + 8)
        else:
            ks = m.get("K", 1)
            ks = sum(ks) if isinstance(ks, (list, tuple)) else ks
            nbytes = m["B"] * (4 * m["N"] * m["C"] + 4 * m["S"] * m["C"] + (8 if m.get("i64") else 4) * m["S"] * ks)
        roof = {"kernel": top_key, "bound": "hbm", "achieved": nbytes / dur_s / 1e9, "peak": pk["hbm"], "unit": "GB/s",
                "frac": nbytes / dur_s / 1e9 / pk["hbm"], "traffic": None, "peak_source": f"{pk['source']} HBM copy",
                "algorithmic_bytes_per_launch": nbytes}
    traffic_file = os.path.join(ROOT, "profiles", "dram_traffic.json")
    if os.path.exists(traffic_file):
        with open(traffic_file) as f:
            t = json.load(f).get(top_key if prec == "bf16" else f"{top_key}|{prec}")
        if t and t.get("batch") == m["B"]:
            roof["traffic"] = t["dram_bytes_per_launch"]
    return roof


def device_time_us(fn, reps=30, warm=5, flush=None):
    """Median device time of one call of `fn` in us: CUDA events around every call, NO host synchronisation between the calls
    (the host enqueues ahead of the GPU, so the bracket holds the kernel and not the Python / ctypes launch latency), an L2
    flush (`flush`: a buffer larger than the 126 MB L2, rewritten outside the brackets) before every call."""
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


def cfg2_inputs():
    import cloudgen
    c = CFG2
    x = cloudgen.uniform_cloud(c["B"], c["N"], c["seed"])
    start = torch.randint(0, c["N"], (c["B"],), generator=torch.Generator().manual_seed(42), dtype=torch.long)
    return x, start


def micro_cfg2(dev, be):
    """configs[2]: FPS + ball query at B=16, N=8192 -> 2048 centroids, r=0.2, nsample=32 (median of 30 launches, us;
    device time, L2 flushed before every launch)."""
    import pointnet2_utils as pu
    c = CFG2
    x, start = cfg2_inputs()
    x, start = x.to(dev), start.to(dev)
    r2 = pu._radius_sq(c["r"])
    flush = torch.zeros(64 * 1024 * 1024, dtype=torch.int32, device=dev)          # 256 MB
    _, _, new_pts = be.fps(x, start, c["S"], want_idx64=True, want_new_points=True)
    res = {"fps_us": device_time_us(lambda: be.fps(x, start, c["S"], want_idx64=True, want_new_points=True), flush=flush),
           "ball_query_us": device_time_us(lambda: be.ball_query(new_pts, x, r2, c["K"], want_idx64=True), flush=flush)}
    del flush
    return res


def cpu_cfg2(threads):
    """The same two calls on the CPU port (one run each: they take seconds)."""
    P = port()
    torch.set_num_threads(threads)
    c = CFG2
    x, start = cfg2_inputs()
    t0 = time.perf_counter()
    idx = P.fps(x, c["S"], start)
    t1 = time.perf_counter()
    cen = P.take_rows(x, idx)
    t2 = time.perf_counter()
    P.ball_query(c["r"], c["K"], x, cen)
    t3 = time.perf_counter()
    return {"cpu_fps_us": 1e6 * (t1 - t0), "cpu_ball_query_us": 1e6 * (t3 - t2), "cpu_cores": threads,
            "cpu_kind": "port (torch CPU restatement of pointnet2_utils.py:47-90), one run"}


def cpu_forward_throughput(cfg_id, budget_s, threads):
    """clouds/s of the CPU port of the reference path at the config's own batch, bounded to ~budget_s."""
    f = FORWARD[cfg_id]
    torch.set_num_threads(threads)
    model = build_model(f["model"], sa_impl=port())
    x = make_inputs(f["cpu_batch"], f["points"], f["seed"])
    times = []
    with torch.no_grad():
        torch.manual_seed(42)
        model(x)                            # warm-up
        t_end = time.perf_counter() + budget_s
        while len(times) < 2 or (time.perf_counter() < t_end and len(times) < 10):
            torch.manual_seed(42)
            t0 = time.perf_counter()
            model(x)
            times.append(time.perf_counter() - t0)
    return f["cpu_batch"] / statistics.median(times), len(times)


# ---------------------------------------------------------------------------------------------------------------------
# configs[0] / configs[1]: whole-model forward
# ---------------------------------------------------------------------------------------------------------------------
def run_forward(cfg_id, args, R):
    import pointnet2_utils as pu
    f = FORWARD[cfg_id]
    dev = R.dev
    pk = peaks()
    be = pu._backend()
    model = build_model(f["model"]).to(dev)
    B, K, W = args.batch or f["batch"], args.steps, max(args.warmup, 3)
    npts = f["points"]
    host = [make_inputs(B, npts, f["seed"] + 97 * R.rank + i).pin_memory() for i in range(N_ROTATING)]
    resident = [h.to(dev) for h in host]
    out_host = torch.empty(B, 7, dtype=torch.float32).pin_memory()
    pred_host = torch.empty(B, dtype=torch.int64).pin_memory()

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

    def measure(prec):
        """value (resident inputs) + e2e (host buffers) + per-kernel trace of one precision mode; every rank runs it."""
        pu.set_precision(prec)
        with torch.no_grad():
            for i in range(W):
                model(resident[i % N_ROTATING])
            R.barrier()
            be.trace = []
            l0 = be.launches
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
            for i in range(K):
                model(resident[i % N_ROTATING])
            ev1.record()
            R.barrier()
            launches = be.launches - l0
            trace, be.trace = be.trace, None
            ms = R.reduce_max(ev0.elapsed_time(ev1))
            run_e2e(2)
            R.barrier()
            ev0.record()
            run_e2e(K)                       # K uploads, K forwards, K read-backs inside the timed region
            ev1.record()
            R.barrier()
            ms_e2e = R.reduce_max(ev0.elapsed_time(ev1))
        return dict(ms=ms, ms_e2e=ms_e2e, launches=launches, trace=trace)

    clocks = ClockSampler(R.local) if R.rank == 0 else None
    other = "fp32" if args.precision == "bf16" else "bf16"
    main = measure(args.precision)
    second = None if args.quick else measure(other)
    pu.set_precision(args.precision)
    extra = {}
    if R.rank == 0 and not args.quick:
        # ---- the reference arm's batch on the GPU (same_batch figure), eager and as one CUDA graph ------------------------
        import pn2_graph
        cb = f["cpu_batch"]
        xb = resident[0][:cb].contiguous()
        with torch.no_grad():
            for _ in range(5):
                model(xb)
            torch.cuda.synchronize(dev)
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
            for _ in range(20):
                model(xb)
            ev1.record()
            torch.cuda.synchronize(dev)
        eager = cb * 20 / (1e-3 * ev0.elapsed_time(ev1))
        g = pn2_graph.GraphedForward(model, xb)
        for _ in range(5):
            g(xb)
        torch.cuda.synchronize(dev)
        ev0.record()
        for _ in range(20):
            g(xb)
        ev1.record()
        torch.cuda.synchronize(dev)
        extra["same_batch_as_reference_sample"] = {"batch": cb, "precision_mode": args.precision, "eager_clouds_per_s": eager,
                                                   "cuda_graph_clouds_per_s": cb * 20 / (1e-3 * ev0.elapsed_time(ev1))}
        if cfg_id == 0:
            extra.update(micro_cfg2(dev, be))
    clk = clocks.stop() if clocks else None
    if R.rank != 0:
        R.finish()
        return
    flops = model_flops_per_cloud(model, npts)

    def record(m, prec):
        kernels, top_key, top = summarise_trace(m["trace"], K)
        value = B * R.world * K / (1e-3 * m["ms"])
        rec = {"value": value, "ms_per_step": m["ms"] / K,
               "e2e": {"value": B * R.world * K / (1e-3 * m["ms_e2e"]), "unit": "clouds/s", "h2d_bytes_per_step": B * 3 * npts * 4,
                       "d2h_bytes_per_step": B * 7 * 4 + B * 8},
               "gpu_launches": m["launches"],
               "roofline": roofline_of(top_key, top, prec, pk, 1e-3 * m["ms"], clk, prec == "fp32" and fp32_on_tensor_cores()),
               "model_tflops": value * flops / 1e12, "kernels": kernels}
        if prec == "bf16":
            peak, _ = pick_peak(pk, 1e-3 * m["ms"], clk)
            rec["model_tensor_frac"] = value / R.world * flops / 1e12 / peak
        return rec

    def fp32_on_tensor_cores():
        return os.environ.get("PN2_F32_FFMA") is None

    cpu = None
    if R.world == 1 and not args.no_cpu_baseline and not args.quick:
        threads = cpu_threads()
        v, n = cpu_forward_throughput(cfg_id, 15.0, threads)
        cpu = {"value": v, "unit": "clouds/s", "cores": threads, "kind": "port",
               "sample": f"{f['model']} forward on {f['cpu_batch']} clouds (configs[{cfg_id}] batch), median of {n} forwards, torch CPU port of the reference path"}
        if cfg_id == 0:
            extra.update(cpu_cfg2(threads))
    rec = record(main, args.precision)
    line = {"metric": f["metric"], "value": rec["value"], "unit": "clouds/s", "n_gpus": R.world, "steps": K, "warmup": W,
            "ms_per_step": rec["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
            "config": workload_config(cfg_id, B, R.world), "precision_mode": args.precision,
            "headline": f"{args.precision} mode (north_star sanctions bf16 at <= 2e-2; the reference arm computes in fp32 -- the "
                        f"like-for-like record is `{other}_mode` below)" if args.precision == "bf16" else "fp32 mode (the reference's own precision)",
            "e2e": rec["e2e"], "gpu_launches": rec["gpu_launches"], "clocks": clk, "roofline": rec["roofline"],
            "model_tflops": rec["model_tflops"], "model_tensor_frac": rec.get("model_tensor_frac"), "kernels": rec["kernels"],
            "cpu_baseline": cpu}
    if second is not None:
        line[f"{other}_mode"] = record(second, other)
    line.update(extra)
    print(json.dumps(line))
    R.finish()


def run_forward_reference(cfg_id, args, R):
    if R.rank != 0:
        return
    f = FORWARD[cfg_id]
    threads = cpu_threads()
    torch.set_num_threads(threads)
    model = build_model(f["model"], sa_impl=port())
    x = make_inputs(f["cpu_batch"], f["points"], f["seed"])
    with torch.no_grad():
        for _ in range(max(1, min(args.warmup, 2))):
            torch.manual_seed(42)
            model(x)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            torch.manual_seed(42)
            model(x)
        dt = time.perf_counter() - t0
    value = f["cpu_batch"] * args.steps / dt
    sample = (f"{f['model']} forward on {f['cpu_batch']} clouds (configs[{cfg_id}] batch) per step, {args.steps} steps; CPU port of the "
              f"reference path (ATen/MKL/oneDNN, fp32), {threads} host threads, no GPU")
    print(json.dumps({
        "impl": "reference", "metric": f["metric"], "value": value, "unit": "clouds/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(cfg_id, args.batch or f["batch"], R.world),
        "cpu_baseline": {"value": value, "unit": "clouds/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "clouds/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


# ---------------------------------------------------------------------------------------------------------------------
# configs[2]: FPS + ball-query microbenchmark
# ---------------------------------------------------------------------------------------------------------------------
def cfg2_config(world):
    c = CFG2
    return {"workload": f"configs[2]: farthest_point_sample + query_ball_point, B={c['B']}, N={c['N']} -> {c['S']} centres, r={c['r']}, "
                        f"K={c['K']}, contiguous [B,N,3] uniform clouds",
            "parallelism": f"{world} independent replica(s); the 16-cloud problem does not shard", "l2": "256 MB buffer rewritten before every launch",
            "reference_arm_sample": "the same two calls, whole problem, on the CPU port"}


def run_cfg2(args, R):
    import pointnet2_utils as pu
    dev, c = R.dev, CFG2
    pk = peaks()
    be = pu._backend()
    clocks = ClockSampler(R.local) if R.rank == 0 else None
    m = micro_cfg2(dev, be)
    # e2e through the public functions: pinned host cloud in, int64 indices out on the host
    x, _ = cfg2_inputs()
    xh = x.pin_memory()
    o1 = torch.empty(c["B"], c["S"], dtype=torch.int64).pin_memory()
    o2 = torch.empty(c["B"], c["S"], c["K"], dtype=torch.int64).pin_memory()

    def e2e_step():
        xd = xh.to(dev, non_blocking=True)
        f = pu.farthest_point_sample(xd, c["S"])
        q = pu.query_ball_point(c["r"], c["K"], xd, pu.index_points(xd, f))
        o1.copy_(f, non_blocking=True)
        o2.copy_(q, non_blocking=True)

    for _ in range(max(3, args.warmup)):
        e2e_step()
    R.barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        e2e_step()
    ev1.record()
    R.barrier()
    e2e_us = R.reduce_max(1e3 * ev0.elapsed_time(ev1) / args.steps)
    fps_us, bq_us = R.reduce_max(m["fps_us"]), R.reduce_max(m["ball_query_us"])
    clk = clocks.stop() if clocks else None
    if R.rank != 0:
        R.finish()
        return
    bytes_fps = c["B"] * (4 * c["N"] * 3 + 4 * c["S"] * 3 + 8 * c["S"])
    bytes_bq = c["B"] * (4 * c["N"] * 3 + 4 * c["S"] * 3 + 8 * c["S"] * c["K"])
    pairs = c["B"] * c["S"] * c["N"]
    cpu = None
    extra = {}
    if R.world == 1 and not args.no_cpu_baseline:
        extra = cpu_cfg2(cpu_threads())
        cpu = {"value": extra["cpu_fps_us"] + extra["cpu_ball_query_us"], "unit": "us", "cores": extra["cpu_cores"], "kind": "port",
               "sample": "the whole configs[2] problem, FPS + ball query once each on the torch CPU port"}
    line = {"metric": "FPS + ball-query us (B=16, N=8192 -> 2048, r=0.2, K=32)", "value": fps_us + bq_us, "unit": "us", "n_gpus": R.world,
            "steps": 30, "warmup": 5, "ms_per_step": 1e-3 * (fps_us + bq_us), "higher_is_better": False, "scaling": "replicas", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": cfg2_config(R.world),
            "e2e": {"value": e2e_us, "unit": "us", "h2d_bytes_per_step": c["B"] * c["N"] * 12, "d2h_bytes_per_step": o1.numel() * 8 + o2.numel() * 8,
                    "what": "public farthest_point_sample -> index_points -> query_ball_point, pinned host cloud in, int64 indices out"},
            "gpu_launches": 2, "clocks": clk, "fps_us": fps_us, "ball_query_us": bq_us,
            "roofline": {"kernel": "ball_query", "bound": "hbm", "achieved": bytes_bq / (bq_us * 1e-6) / 1e9, "peak": pk["hbm"], "unit": "GB/s",
                         "frac": bytes_bq / (bq_us * 1e-6) / 1e9 / pk["hbm"], "traffic": None, "peak_source": f"{pk['source']} HBM copy",
                         "algorithmic_bytes_per_launch": bytes_bq,
                         "note": "nominally HBM-bound, practically FP32-issue bound: see pair_evals_per_s (SURVEY.md 8d)"},
            "roofline_fps": {"bound": "latency", "achieved_gbs": bytes_fps / (fps_us * 1e-6) / 1e9, "distance_updates_per_s": pairs / (fps_us * 1e-6),
                             "iterations": c["S"], "ns_per_iteration": 1e3 * fps_us / c["S"]},
            "pair_evals_per_s": pairs / (bq_us * 1e-6), "cpu_baseline": cpu}
    line.update(extra)
    print(json.dumps(line))
    R.finish()


def run_cfg2_reference(args, R):
    if R.rank != 0:
        return
    threads = cpu_threads()
    n = max(1, min(args.steps, 3))
    runs = [cpu_cfg2(threads) for _ in range(n)]
    fps = statistics.median(r["cpu_fps_us"] for r in runs)
    bq = statistics.median(r["cpu_ball_query_us"] for r in runs)
    sample = f"the whole configs[2] problem, median of {n} run(s) of FPS + ball query on the torch CPU port, {threads} host threads"
    print(json.dumps({"impl": "reference", "metric": "FPS + ball-query us (B=16, N=8192 -> 2048, r=0.2, K=32)", "value": fps + bq, "unit": "us",
                      "n_gpus": args.gpus, "steps": n, "warmup": 0, "ms_per_step": 1e-3 * (fps + bq), "higher_is_better": False,
                      "scaling": "replicas", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg2_config(R.world),
                      "cpu_fps_us": fps, "cpu_ball_query_us": bq,
                      "cpu_baseline": {"value": fps + bq, "unit": "us", "cores": threads, "kind": "port", "sample": sample},
                      "e2e": {"value": fps + bq, "unit": "us", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}))


# ---------------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, default=0, choices=[0, 1, 2, 3, 4], help="BASELINE.json configs[i]")
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--batch", type=int, default=0, help="clouds per GPU per step (configs 0, 1, 4) / global batch (config 3); 0 = the config's default")
    ap.add_argument("--precision", choices=["bf16", "fp32"], default="bf16")
    ap.add_argument("--graph", type=int, default=2, choices=[0, 1, 2],
                    help="config 3: 0 = eager heads, 1 = one CUDA graph per head back to back, 2 = the graphs on one stream per head")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--quick", action="store_true", help="timed regions of the headline mode only (for runs under a profiler)")
    args = ap.parse_args()
    R = Ranks()
    if args.impl == "reference":
        if args.config in (0, 1):
            run_forward_reference(args.config, args, R)
        elif args.config == 2:
            run_cfg2_reference(args, R)
        elif args.config == 3:
            import bench_pipeline
            bench_pipeline.run_reference(args, R)
        else:
            import bench_train
            bench_train.run_reference(args, R)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU path (use --impl reference for the CPU baseline)")
    R.init_cuda()
    if args.config in (0, 1):
        run_forward(args.config, args, R)
    elif args.config == 2:
        run_cfg2(args, R)
    elif args.config == 3:
        import bench_pipeline
        bench_pipeline.run(args, R)
    else:
        import bench_train
        bench_train.run(args, R)


if __name__ == "__main__":
    main()
