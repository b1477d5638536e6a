#!/usr/bin/env python
"""Benchmark of the AdaFocus offline-inference hot path on N B200s (BASELINE.json configs 2-5).

  python bench.py --gpus N --steps K --warmup W                 # cfg3 (headline): our CUDA path, 64 clips / GPU / step
  python bench.py --impl reference --steps K --warmup W         # the reference's own CPU implementation, host cores
  python bench.py --workload cfg2|cfg4|cfg5 ...                 # fL only / global batch 256 / Something-Something shape

A "step" is one pass of the whole path -- fG over B*T_g frames, policy rollout, crop, fL over B*T_f patches, classifier
-- over one batch of B synthetic clips per GPU.  `value` = clips/s with the clips already resident in HBM; `e2e` = the
same through the host-buffer API (pinned HOST tensors, H2D of the clips and D2H of the logits inside the timed region);
`value_api` = device-resident clips through the reference-facing call `model(input=..., scan=..., one_step=True)`;
`torch_gpu_baseline` = the UNMODIFIED reference classes (baseline/_ref, stock PyTorch / cuDNN) on the same B200.
One JSON line on stdout (rank 0).
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

UNIT = "clips/s"
METRICS = {
    "cfg3": "clips/sec (16-frame, 128^2 patch)",
    "cfg4": "clips/sec (16-frame, 128^2 patch)",
    "cfg2": "clips/sec (16-frame, 128^2 patch), fL only",
    "cfg5": "videos/sec (Sth-Sth shape, 8+12 frames, 144^2 patch, ResNet-101 fL)",
    "cfg1": "clips/sec (8-frame 224^2 clip, MobileNet-v2 fG only)",
}
WORKLOADS = {
    "cfg3": ("cfg3: full AdaFocus (ACT tree) MobileNet-V2 fG + ResNet-50 fL + GRU policy/classifier, T=16, 224^2 "
             "frames, P=128, 49 actions, 200 classes, synthetic ActivityNet-shape clips"),
    "cfg4": ("cfg4: cfg3 at a GLOBAL batch of 256 clips sharded over the ranks (32 clips/GPU at 8 GPUs), one NCCL "
             "all-gather of the (B,200) logits"),
    "cfg2": ("cfg2: ResNet-50 fL trunk + global average pool over pre-cropped 16 x 3 x 128^2 synthetic patches per "
             "clip (fL kernels only)"),
    "cfg5": ("cfg5: Something-Something shape (STH tree), TSM-MobileNet-V2 fG over 8 frames + continuous policy + "
             "TSM-ResNet-101 fL over 12 patches of 144^2, 174 classes"),
    "cfg1": ("cfg1: single 8-frame 224^2 synthetic clip, TSM-MobileNet-V2 fG only (STH tree GFV.glance(), the call "
             "evaluate.py makes at STH/evaluate.py:195), 174 classes"),
}
FL_GFLOP_PER_PATCH = 2.669      # ResNet-50 trunk @128^2, 2*MAC
R101_GFLOP_PER_PATCH_144 = 6.583
TOTAL_GFLOP_PER_CLIP = 53.04


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg3", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=None, help="clips per GPU per step (default 64; cfg4: 256 / gpus)")
    ap.add_argument("--global-batch", type=int, default=256, help="cfg4 only")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-torch-gpu-baseline", action="store_true")
    ap.add_argument("--focuser-frames", type=int, default=12, help="cfg5 only: T_f in {8, 12, 16}")
    return ap.parse_args()


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return d.get("bf16_tflops_sustained", 1389.0), d.get("hbm_gbs", 6551.0), "measured (MEASURED_PEAKS.json)"
    return 1400.0, 6650.0, "fallback (B200_PROFILING.md)"


def committed_traffic(workload):
    """DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) of the dominant stage from the committed ncu launch
    list of THIS build: profiles/traffic.json records the hash of the CUDA sources it was captured from; a stale
    record yields None instead of a number that no longer describes the kernels."""
    from adafocus_b200 import build as b
    path = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        with open(path) as f:
            rec = json.load(f)
    except (OSError, ValueError):
        return None, "no profiles/traffic.json"
    ent = rec.get(workload)
    if not ent:
        return None, f"no {workload} entry in profiles/traffic.json"
    if ent.get("src_hash") != b.source_hash():
        return None, f"stale: profiles/traffic.json was captured from another build ({ent.get('launch_list')})"
    return ent, ent.get("launch_list")


# ------------------------------------------------------------------------------------------------ reference (CPU)
def _act_args(workload, batch):
    from adafocus_b200 import synth
    return synth.act_args(batch_size=batch)


def reference_cpu(workload, steps, warmup, clips_per_step=None, budget_s=150.0, focuser_frames=12):
    """Times the reference's CPU implementation of the path on the host cores: the UNMODIFIED reference classes when
    the reference is on this machine (baseline/_ref or /root/reference; kind "reference"), else the oracle port
    (kind "port").  Threads: the faster of {nproc, nproc/2}.  Returns (value, info)."""
    import torch
    from adafocus_b200 import synth
    from oracle import reference_loader as rl
    from oracle import reference_runner as rr
    cores, cpu_model = rr.host_info()
    cands = sorted({max(1, min(cores, 64)), max(1, min(cores, 64) // 2)}, reverse=True)
    tree = "STH" if workload in ("cfg5", "cfg1") else "ACT"
    kind = "reference" if rl.available(tree) else "port"
    with torch.no_grad(), rl.force_cpu():      # the reference hard-codes .cuda(): keep the CPU arm on the CPU
        if workload == "cfg1":
            args = synth.sth_args()
            if kind == "reference":
                model, _ = rr.build_sth(args)
                what = "reference STH GFV.glance (STH/models/gfv_net.py:101-112, evaluate.py:188-196)"

                def make(n):
                    gi = synth.synth_clips(n, args.num_segments_glancer, 224, synth.SEED + 1)
                    return lambda: model.glance(torch.nn.functional.interpolate(gi, (args.glance_size, args.glance_size)))
            else:
                from adafocus_b200.models_sth.gfv_net import GFV
                from oracle import adafocus_oracle as orc
                m = GFV(args)
                synth.strip_fc_sth(m)
                ck = synth.synth_checkpoint_sth(m)
                del m
                what = "oracle.mobilenet_v2_features_flat (TSM) + classifier"

                def make(n):
                    gi = synth.synth_clips(n, args.num_segments_glancer, 224, synth.SEED + 1)
                    return lambda: orc.mobilenet_v2_features_flat(gi.view(n * 8, 3, 224, 224), ck["glancer"], "net.features.",
                                                                  (8, args.shift_div))
        elif workload == "cfg5":
            args = synth.sth_args(base_model="resnet101", num_segments_focuser=focuser_frames)
            if kind == "reference":
                model, _ = rr.build_sth(args)
                what = ("reference STH GFV.glance + action_stage2 (STH/evaluate.py:188-201, incl. the random-patch "
                        "baseline pass)")

                def make(n):
                    gi = synth.synth_clips(n, args.num_segments_glancer, 224, synth.SEED + 1)
                    fi = synth.synth_clips(n, args.num_segments_focuser, 224, synth.SEED + 2)
                    return lambda: rr.sth_forward(model, gi, fi, args, with_baseline=True)
            else:
                from adafocus_b200.models_sth.gfv_net import GFV
                from oracle import adafocus_oracle as orc
                m = GFV(args)
                synth.strip_fc_sth(m)
                ck = synth.synth_checkpoint_sth(m)
                del m
                what = "oracle.sth_forward (pred only)"

                def make(n):
                    gi = synth.synth_clips(n, args.num_segments_glancer, 224, synth.SEED + 1)
                    fi = synth.synth_clips(n, args.num_segments_focuser, 224, synth.SEED + 2)
                    return lambda: orc.sth_forward(gi, fi, ck, args.patch_size, args.num_segments_glancer,
                                                   args.num_segments_focuser, args.video_div, args.shift_div,
                                                   layers=(3, 4, 23, 3))
        elif workload == "cfg2":
            args = synth.act_args()
            if kind == "reference":
                model, _ = rr.build_act(args)
                net = model.focuser.net
                what = "reference ResNet.get_featmap(pooled=True) (ACT/models/resnet.py:211-225) over 16 patches / clip"

                def make(n):
                    g = torch.Generator().manual_seed(synth.SEED)
                    x = torch.randn(n * 16, 3, 128, 128, generator=g)
                    return lambda: net.get_featmap(x, pooled=True)
            else:
                from adafocus_b200.models.gfv_net import GFV
                from oracle import adafocus_oracle as orc
                m = GFV(args)
                ck = synth.synth_checkpoint_act(m)
                del m
                what = "oracle.resnet_trunk"

                def make(n):
                    g = torch.Generator().manual_seed(synth.SEED)
                    x = torch.randn(n * 16, 3, 128, 128, generator=g)
                    return lambda: orc.resnet_trunk(x, ck["focuser"], "net.")
        else:
            args = synth.act_args()
            if kind == "reference":
                model, _ = rr.build_act(args)
                what = "reference ACT GFV.forward(one_step=True) (ACT/main_dist.py:332,368)"

                def make(n):
                    x = synth.synth_clips(n, args.num_segments, args.input_size)
                    return lambda: rr.act_forward(model, x, args.glance_size)
            else:
                from adafocus_b200.models.gfv_net import GFV
                from oracle import adafocus_oracle as orc
                m = GFV(args)
                ck = synth.synth_checkpoint_act(m)
                del m
                what = "oracle.act_forward"

                def make(n):
                    x = synth.synth_clips(n, args.num_segments, args.input_size)
                    return lambda: orc.act_forward(x, x, ck, args.patch_size, args.action_dim)
        one = make(1)
        t0 = time.perf_counter()
        threads, seen = rr.pick_threads(one, cands)
        sweep_s = time.perf_counter() - t0
        warm1 = seen[threads]
        if clips_per_step is None:
            clips_per_step = int(max(1, min(8, budget_s / max(1e-3, warm1 * (steps + warmup)))))
        fn = make(clips_per_step) if clips_per_step != 1 else one
        ts = rr._time_cpu(fn, steps, warmup)
    total = sum(ts)
    value = clips_per_step * steps / total
    med = sorted(ts)[len(ts) // 2]
    info = {"value": value, "unit": UNIT, "cores": threads, "kind": kind,
            "sample": f"{steps} timed steps x {clips_per_step} clip(s) of the {workload} workload through {what}, fp32, "
                      f"torch {torch.__version__} CPU ops, {threads} threads (sweep "
                      f"{ {k: round(v, 3) for k, v in seen.items()} } s/clip) of {cores} host cores [{cpu_model}]; "
                      f"median {clips_per_step / med:.2f} clips/s, best {clips_per_step / min(ts):.2f}; cold start + "
                      f"thread sweep {sweep_s:.1f}s excluded",
            "ms_per_step": 1e3 * total / steps, "clips_per_step": clips_per_step}
    return value, info


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    value, info = reference_cpu(a.workload, a.steps, max(1, a.warmup), focuser_frames=a.focuser_frames)
    line = {
        "impl": "reference", "metric": METRICS[a.workload], "value": value, "unit": UNIT, "n_gpus": a.gpus,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": info["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOADS[a.workload], "clips_per_step": info["clips_per_step"],
                   "note": "the reference's own CPU implementation of the path on the host cores; a bounded sample "
                           "(clips per step) of the same per-clip workload"},
        "cpu_baseline": {k: info[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ reference (GPU)
def torch_gpu_baseline(workload, batch, dev, focuser_frames=12):
    """The UNMODIFIED reference classes through stock PyTorch/cuDNN on this B200 (BASELINE.md
    section 3 'the real bar'): (i) as shipped -- fp32 NCHW, cudnn.benchmark=True (ACT/main_dist.py:190), TF32 convs
    allowed (torch default); (ii) strict fp32; (iii) fp16 autocast + channels_last.  CUDA-event timed, same batch."""
    import torch
    from adafocus_b200 import synth
    from oracle import reference_loader as rl
    from oracle import reference_runner as rr
    tree = "STH" if workload in ("cfg5", "cfg1") else "ACT"
    if not rl.available(tree):
        return {"unavailable": "reference sources are not on this machine (baseline/_ref missing)"}
    out = {"batch": batch, "source": rl.tree_path(tree).replace(ROOT + "/", ""), "torch": torch.__version__,
           "cudnn": torch.backends.cudnn.version()}
    prev = (torch.backends.cudnn.benchmark, torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    try:
        torch.backends.cudnn.benchmark = True
        with torch.no_grad():
            if workload == "cfg1":
                args = synth.sth_args(batch_size=batch)
                model, _ = rr.build_sth(args, dev)
                gi = synth.synth_clips(batch, args.num_segments_glancer, 224, synth.SEED + 1).to(dev)
                fn = lambda: model.glance(gi)                                                # noqa: E731
                fn2 = None
                out["api"] = "GFV.glance (TSM-MobileNet-V2 features + per-frame logits)"
            elif workload == "cfg5":
                args = synth.sth_args(base_model="resnet101", num_segments_focuser=focuser_frames, batch_size=batch)
                model, _ = rr.build_sth(args, dev)
                gi = synth.synth_clips(batch, args.num_segments_glancer, 224, synth.SEED + 1).to(dev)
                fi = synth.synth_clips(batch, args.num_segments_focuser, 224, synth.SEED + 2).to(dev)
                fn = lambda: rr.sth_forward(model, gi, fi, args, with_baseline=False)       # noqa: E731
                fn2 = lambda: rr.sth_forward(model, gi, fi, args, with_baseline=True)       # noqa: E731
                out["api"] = "glance + action_stage3 (pred only); *_with_baseline = action_stage2 as evaluate.py drives it"
            elif workload == "cfg2":
                args = synth.act_args(batch_size=batch)
                model, _ = rr.build_act(args, dev)
                net = model.focuser.net
                x = torch.randn(batch * 16, 3, 128, 128, device=dev,
                                generator=torch.Generator(device=dev).manual_seed(synth.SEED))
                fn = lambda: net.get_featmap(x, pooled=True)                                # noqa: E731
                fn2 = None
                out["api"] = "ResNet.get_featmap(pooled=True) over B*16 patches in one batch"
            else:
                args = synth.act_args(batch_size=batch)
                model, _ = rr.build_act(args, dev)
                x = synth.synth_clips(min(batch, 8), args.num_segments, args.input_size).to(dev)
                x = x.repeat((batch + x.shape[0] - 1) // x.shape[0], 1, 1, 1)[:batch].contiguous()
                fn = lambda: rr.act_forward(model, x, args.glance_size, gpu=dev.index)      # noqa: E731
                fn2 = None
                out["api"] = "GFV.forward(input, scan, one_step=True) (ACT/main_dist.py:332,368)"

            def timed(f, steps=5, warmup=3):
                ms = rr.time_gpu(f, steps, warmup, dev)
                return {"value": batch / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms}

            torch.backends.cudnn.allow_tf32 = True
            torch.backends.cuda.matmul.allow_tf32 = False       # torch default: TF32 for cuDNN convs only
            out["fp32_tf32_as_shipped"] = timed(fn)
            if fn2 is not None:
                out["fp32_tf32_as_shipped_with_baseline"] = timed(fn2)
            torch.backends.cudnn.allow_tf32 = False
            out["fp32_strict"] = timed(fn)
            torch.backends.cudnn.allow_tf32 = True

            def amp(f):
                def g():
                    with torch.autocast("cuda", dtype=torch.float16):
                        return f()
                return g
            try:
                model.to(memory_format=torch.channels_last)
                out["fp16_autocast_channels_last"] = timed(amp(fn))
                if fn2 is not None:
                    out["fp16_autocast_channels_last_with_baseline"] = timed(amp(fn2))
            except RuntimeError as exc:
                # the STH tree .view()s activations (STH/ops/temporal_shift.py:31): channels_last tensors cannot be viewed
                out["fp16_autocast_channels_last"] = {"failed": str(exc).splitlines()[0][:160]}
                model.to(memory_format=torch.contiguous_format)
                out["fp16_autocast"] = timed(amp(fn))
                if fn2 is not None:
                    out["fp16_autocast_with_baseline"] = timed(amp(fn2))
        out["best"] = max(v["value"] for v in out.values() if isinstance(v, dict) and "value" in v)
    except Exception as exc:                       # the baseline is a report, never the product path
        out["failed"] = f"{type(exc).__name__}: {exc}"
    finally:
        torch.backends.cudnn.benchmark, torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = prev
        try:
            del model
        except NameError:
            pass
        rl.unload()
        torch.cuda.empty_cache()
    return out


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """nvidia-smi sampler (the profiling recipe's clocks line); keeps only the samples taken inside the timed region."""
    FIELDS = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.proc = None
        self.index = index
        self.t0 = self.t1 = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", "25"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def region_begin(self):
        self.t0 = time.time()

    def region_end(self):
        self.t1 = time.time()

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.1)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        import datetime
        sm_in, sm_all, mx, reasons, power = [], [], None, set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in out.strip().splitlines():
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 8:
                continue
            try:
                ts = datetime.datetime.strptime(parts[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                clk = float(parts[1])
                mx = float(parts[2])
            except ValueError:
                continue
            sm_all.append(clk)
            inside = self.t0 is not None and self.t0 - 0.03 <= ts <= self.t1 + 0.03
            if inside:
                sm_in.append(clk)
                try:
                    power.append(float(parts[3]))
                except ValueError:
                    pass
                for nm, v in zip(names, parts[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
        use = sorted(sm_in) if sm_in else sorted(sm_all)
        med = use[len(use) // 2] if use else None
        return {"sm_mhz": med, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm_in),
                "power_w_max": max(power) if power else None}


# ------------------------------------------------------------------------------------------------ ours: shared pieces
class Harness:
    """Process-group setup, barrier, max-over-ranks device timing -- the contract's timed region."""

    def __init__(self, a):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py --impl ours needs a CUDA device: adafocus_b200 has no CPU path")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        from adafocus_b200 import numa
        self.numa = numa.bind_to_gpu(self.local)          # host threads + pinned allocations next to this GPU
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=self.dev)
        self.a = a
        self.steps, self.warmup = a.steps, max(3, a.warmup)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, ms):
        t = self.torch.tensor([ms], device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def timed(self, step, sampler=None, steps=None, warmup=None):
        """W untimed steps, then exactly K steps between barrier + synchronize; CUDA events; max over ranks."""
        torch = self.torch
        steps = steps or self.steps
        for _ in range(warmup or self.warmup):
            step()
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.barrier()
        if sampler:
            sampler.region_begin()
        e0.record()
        for _ in range(steps):
            step()
        e1.record()
        self.barrier()
        if sampler:
            sampler.region_end()
        return self.max_over_ranks(e0.elapsed_time(e1))

    def finish(self):
        if self.world > 1:
            self.dist.destroy_process_group()


def h2d_ceiling(h, nbytes, reps=6):
    """Bare pinned-host -> device cudaMemcpyAsync rate of this rank while ALL ranks copy at once (the ceiling of the
    fp32 e2e path).  Returns GB/s per rank, max-over-ranks time."""
    torch = h.torch
    host = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    dst = torch.empty(nbytes, dtype=torch.uint8, device=h.dev)
    for _ in range(2):
        dst.copy_(host, non_blocking=True)
    h.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        dst.copy_(host, non_blocking=True)
    e1.record()
    h.barrier()
    ms = h.max_over_ranks(e0.elapsed_time(e1)) / reps
    del host, dst
    return nbytes / (ms * 1e-3) / 1e9


# ------------------------------------------------------------------------------------------------ cfg3 / cfg4
def run_act(a):
    h = Harness(a)
    torch, dist, dev, world, rank = h.torch, h.dist, h.dev, h.world, h.rank
    from adafocus_b200 import synth
    from adafocus_b200.models.gfv_net import GFV
    from adafocus_b200.pipeline import StreamingEvaluator

    if a.workload == "cfg4":
        if a.global_batch % world:
            raise SystemExit("--global-batch must be divisible by the number of ranks")
        b = a.batch or a.global_batch // world
        scaling = "strong"
    else:
        b = a.batch or 64
        scaling = "weak"
    args = synth.act_args(batch_size=b)
    t_pack = time.perf_counter()
    model = GFV(args)
    synth.load_checkpoint_act(model, synth.synth_checkpoint_act(model))
    model = model.to(dev)
    model.eval()
    t, s, c = args.num_segments, args.input_size, args.num_classes

    # synthetic clips, generated on the device (seeded), already resident in HBM for the device-timed loop
    gen = torch.Generator(device=dev).manual_seed(synth.SEED + rank)
    inp, _ = model.input_buffers(b, dev)
    coarse = torch.randn(b * t, 3, 7, 7, generator=gen, device=dev)
    inp.copy_((torch.nn.functional.interpolate(coarse, size=(s, s)) +
               0.5 * torch.randn(b * t, 3, s, s, generator=gen, device=dev)).view(b, 3 * t, s, s))
    plan = model.fused_plan(b, t, s, s, args.glance_size, dev, True)
    plan.run()
    torch.cuda.synchronize()
    first_forward_s = time.perf_counter() - t_pack      # model construction + weight packing + plan recording + 1 run
    gathered = torch.empty(world * b, c, device=dev) if world > 1 else None

    def step():
        plan.run()
        last = plan.logits.view(b, t, -1)[:, -1, :c]
        if world > 1:
            dist.all_gather_into_tensor(gathered, last.contiguous())     # one NCCL all-gather of per-clip logits
        return last

    sampler = ClockSampler(h.local) if rank == 0 else None
    if sampler:
        sampler.start()
    ms_total = h.timed(step, sampler)
    clocks = sampler.stop() if sampler else None
    stage_acc = plan.stage_ms()            # stages of the last timed step, from CUDA-event marks inside the plan
    ms_step = ms_total / a.steps
    value = world * b * a.steps / (ms_total * 1e-3)

    # ---- the reference-facing call on device-resident clips: model(input=..., scan=..., one_step=True)
    def api_step():
        logits, last = model(input=inp, scan=inp, training=False, backbone_pred=False, one_step=True, gpu=h.local)
        if world > 1:
            dist.all_gather_into_tensor(gathered, last.contiguous())
        return last
    api_ms = h.timed(api_step) / a.steps
    value_api = {"value": world * b / (api_ms * 1e-3), "unit": UNIT, "ms_per_step": api_ms,
                 "api": "adafocus_b200.models.gfv_net.GFV.forward(input=, scan=, one_step=True) -> (logits, last_out)"}

    # ---- end to end from pinned host buffers (H2D + compute + D2H inside the timed region)
    e2e = e2e_u8 = e2e_raw = ceiling = None

    def time_e2e(ev, host):
        nb = a.steps
        ev.run([host[i % 2] for i in range(h.warmup)], collect=False)
        h.barrier()
        t0 = time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ev.run([host[i % 2] for i in range(nb)], collect=False)
        if world > 1:
            dist.all_gather_into_tensor(gathered, ev.plans[(nb - 1) % 2].logits.view(b, t, -1)[:, -1, :c].contiguous())
        e1.record()
        h.barrier()
        wall = time.perf_counter() - t0
        e2e_ms = h.max_over_ranks(max(e0.elapsed_time(e1), 0.0))
        return {"value": world * b * nb / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": ev.h2d_bytes * world,
                "d2h_bytes_per_step": ev.d2h_bytes * world, "ms_per_step": e2e_ms / nb, "wall_s": wall}

    if not a.no_e2e:
        # (1) the reference's input contract: pinned fp32 (B,3T,H,W) batches, as its DataLoader produces them
        ev = StreamingEvaluator(model, b, dev, slots=2)
        host = [torch.empty(b, 3 * t, s, s, dtype=torch.float32).pin_memory() for _ in range(2)]
        for hbuf in host:
            hbuf.copy_(inp.cpu())
        e2e = time_e2e(ev, host)
        e2e["api"] = ("adafocus_b200.pipeline.StreamingEvaluator.run (2 input slots, chunked H2D overlapped with "
                      "compute)")
        gbs = h2d_ceiling(h, host[0].numel() * 4)
        ceiling = {"h2d_gbs_per_gpu_all_ranks_copying": gbs,
                   "clips_per_s_if_copy_bound": world * gbs * 1e9 / (3 * t * s * s * 4),
                   "numa": h.numa}
        del host
        # (2) same clips shipped as stacked uint8 frames (B,H,W,3T), ToTorchFormatTensor + GroupNormalize on the device
        ev8 = StreamingEvaluator(model, b, dev, slots=2, input_format="u8_hwc")
        gen8 = torch.Generator().manual_seed(1007)
        host8 = [torch.randint(0, 256, (b, s, s, 3 * t), dtype=torch.uint8, generator=gen8).pin_memory()
                 for _ in range(2)]
        e2e_u8 = time_e2e(ev8, host8)
        e2e_u8["api"] = ("StreamingEvaluator(input_format='u8_hwc'): uint8 frames over PCIe, af_frames_u8_to_f32 "
                         "(bit-identical to the reference's transform chain) in front of the plan")
        # (3) decoded frames at their stored size (340x256, the reference's extracted-JPEG height, ACT/ops/video_jpg.py):
        #     GroupScale + GroupCenterCrop + Stack + ToTorchFormatTensor + GroupNormalize all on the device (f-2)
        evr = StreamingEvaluator(model, b, dev, slots=2, input_format="u8_frames", frame_hw=(256, 340))
        hostr = [torch.randint(0, 256, (b * t, 256, 340, 3), dtype=torch.uint8, generator=gen8).pin_memory()
                 for _ in range(2)]
        e2e_raw = time_e2e(evr, hostr)
        e2e_raw["api"] = ("StreamingEvaluator(input_format='u8_frames'): decoded 340x256 uint8 frames over PCIe, "
                          "af_resize_crop_u8 (Pillow-exact GroupScale + GroupCenterCrop) + af_frames_u8_to_f32 on the device")
        del host8, hostr, ev, ev8, evr

    if rank != 0:
        h.finish()
        return

    # ---- crop / gather roofline (get_patch as one HBM-bound kernel, ACT/models/utils.py:37-51), measured live
    from adafocus_b200.engine import get_engine
    eng = get_engine(dev)
    frames = inp.view(b * t, 3, s, s)
    acts = torch.rand(b * t, 2, device=dev, generator=gen)
    patches = torch.empty(b * t, 3, args.patch_size, args.patch_size, device=dev)
    for _ in range(3):
        eng.crop(frames, action=acts, patch=args.patch_size, out=patches)
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 20
    c0.record()
    for _ in range(reps):
        eng.crop(frames, action=acts, patch=args.patch_size, out=patches)
    c1.record()
    torch.cuda.synchronize()
    crop_ms = c0.elapsed_time(c1) / reps
    crop_bytes = b * t * 2 * 3 * args.patch_size * args.patch_size * 4
    del patches

    tf_peak, hbm_peak, peak_src = measured_peaks()
    fl_ms = stage_acc["fL"]
    fl_tflops = b * t * FL_GFLOP_PER_PATCH * 1e9 / (fl_ms * 1e-3) / 1e12
    tr, tr_src = committed_traffic("cfg3")
    roofline = {
        "bound": "tensor", "achieved": fl_tflops, "peak": tf_peak, "unit": "TFLOP/s", "frac": fl_tflops / tf_peak,
        # dram__bytes_read.sum + dram__bytes_write.sum over the fL launches of one replay, from the committed ncu launch
        # list of THIS build (hash-checked), scaled linearly with the number of patches; None when stale
        "traffic": (tr["fl_dram_bytes"] * (b * t) / tr["patches"]) if tr else None,
        "traffic_source": tr_src,
        "kernel": "conv_gemm_kernel / bottleneck kernels (tcgen05 implicit GEMM), all fL launches of one step",
        "how": f"{b * t} patches x {FL_GFLOP_PER_PATCH} GFLOP (ResNet-50 trunk @128^2) / fL stage time {fl_ms:.3f} ms "
               f"measured by CUDA-event marks inside the timed plan replay (includes stem staging, maxpool, avgpool); "
               f"peak = sustained bf16/fp16 dense, {peak_src}",
    }
    line = {
        "metric": METRICS[a.workload], "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps,
        "warmup": h.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
        "dtype": "f16", "data": "synthetic",
        "config": {"workload": WORKLOADS[a.workload], "clips_per_gpu_per_step": b, "global_clips_per_step": b * world,
                   "parallelism": f"dp{world} (clips sharded, weights replicated, one all-gather of (B,200) logits)",
                   "l2": f"per-step input {b * 3 * t * s * s * 4 / 2**20:.0f} MiB/GPU exceeds the 126 MB L2; no flush"},
        "clocks": clocks, "e2e": e2e, "e2e_u8_frames": e2e_u8, "e2e_decoded_frames": e2e_raw, "e2e_h2d_ceiling": ceiling,
        "value_api": value_api,
        "gpu_launches": plan.plan.num_launches * a.steps,
        "roofline": roofline, "stages_ms": stage_acc, "first_forward_s": first_forward_s,
        "crop_roofline": {"bound": "hbm", "achieved": crop_bytes / (crop_ms * 1e-3) / 1e9, "peak": hbm_peak,
                          "unit": "GB/s", "frac": crop_bytes / (crop_ms * 1e-3) / 1e9 / hbm_peak,
                          "kernel": "crop_nchw_f32_vec4_kernel (get_patch, fp32 -> fp32)",
                          "how": f"{b * t} patches x 393216 B (read + write) / {crop_ms * 1e3:.1f} us (CUDA events, "
                                 f"{reps} launches, source frames {b * 3 * t * s * s * 4 / 2**20:.0f} MiB > L2)"},
        "model_tflops": value * TOTAL_GFLOP_PER_CLIP / 1e3 / world,
    }
    if world == 1 and not a.no_torch_gpu_baseline:
        del plan
        model._plans.clear()
        torch.cuda.empty_cache()
        line["torch_gpu_baseline"] = torch_gpu_baseline("cfg3", b, dev)
    if world == 1 and not a.no_cpu_baseline:
        try:
            _, info = reference_cpu("cfg3", 3, 1, budget_s=25.0)
            line["cpu_baseline"] = {k: info[k] for k in ("value", "unit", "cores", "kind", "sample")}
        except Exception as exc:   # the baseline is a report, never the product path
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": f"failed: {exc}"}
    print(json.dumps(line), flush=True)
    h.finish()


# ------------------------------------------------------------------------------------------------ cfg2 (fL only)
def run_cfg2(a):
    h = Harness(a)
    torch, dev, world, rank = h.torch, h.dev, h.world, h.rank
    from adafocus_b200 import synth
    from adafocus_b200.engine import get_engine
    from adafocus_b200.models.gfv_net import GFV
    b = a.batch or 64
    n, p = b * 16, 128
    args = synth.act_args(batch_size=b)
    model = GFV(args)
    synth.load_checkpoint_act(model, synth.synth_checkpoint_act(model))
    model = model.to(dev)
    model.eval()
    eng = get_engine(dev)
    runner = model.focuser.net.runner()
    gen = torch.Generator(device=dev).manual_seed(synth.SEED + rank)
    slots = []
    for _ in range(2):
        patches = torch.randn(n, 3, p, p, device=dev, generator=gen)
        feat = torch.zeros(n, 2048, dtype=torch.float16, device=dev)
        eng.begin_plan()
        try:
            m0 = eng.mark()
            runner.run_pooled_chunked(eng, patches, feat, 2048, None)
            m1 = eng.mark()
        finally:
            plan = eng.end_plan()
        slots.append((patches, feat, plan, (m0, m1)))
    patches, feat, plan, marks = slots[0]
    stream = lambda: torch.cuda.current_stream(dev).cuda_stream      # noqa: E731
    sampler = ClockSampler(h.local) if rank == 0 else None
    if sampler:
        sampler.start()
    ms_total = h.timed(lambda: plan.run(stream()), sampler)
    clocks = sampler.stop() if sampler else None
    ms_step = ms_total / a.steps
    fl_ms = plan.elapsed_ms(*marks)
    value = world * b / (ms_step * 1e-3)

    # e2e: pinned fp32 patches -> H2D -> fL -> D2H of the pooled features (fp16)
    e2e = None
    if not a.no_e2e:
        host = [torch.randn(n, 3, p, p).pin_memory() for _ in range(2)]
        out_host = [torch.empty(n, 2048, dtype=torch.float16).pin_memory() for _ in range(2)]
        copy_s, comp_s = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        copied = [torch.cuda.Event() for _ in range(2)]
        consumed = [torch.cuda.Event() for _ in range(2)]

        def run_stream(nb):
            for sl in range(2):
                consumed[sl].record(comp_s)
            for i in range(nb):
                sl = i % 2
                pt, ft, pl, _ = slots[sl]
                with torch.cuda.stream(copy_s):
                    copy_s.wait_event(consumed[sl])
                    pt.copy_(host[sl], non_blocking=True)
                    copied[sl].record(copy_s)
                with torch.cuda.stream(comp_s):
                    comp_s.wait_event(copied[sl])
                    pl.run(comp_s.cuda_stream)
                    consumed[sl].record(comp_s)
                    out_host[sl].copy_(ft, non_blocking=True)
            comp_s.synchronize()
        run_stream(h.warmup)
        h.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run_stream(a.steps)
        e1.record()
        h.barrier()
        ems = h.max_over_ranks(e0.elapsed_time(e1)) / a.steps
        e2e = {"value": world * b / (ems * 1e-3), "unit": UNIT, "h2d_bytes_per_step": world * n * 3 * p * p * 4,
               "d2h_bytes_per_step": world * n * 2048 * 2, "ms_per_step": ems,
               "api": "pinned fp32 patches -> recorded fL plan (ResNetRunner.run_pooled_chunked) -> pinned fp16 features"}
        del host
    if rank != 0:
        h.finish()
        return
    tf_peak, _, peak_src = measured_peaks()
    tfl = n * FL_GFLOP_PER_PATCH * 1e9 / (fl_ms * 1e-3) / 1e12
    tr, tr_src = committed_traffic("cfg3")
    line = {
        "metric": METRICS["cfg2"], "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": h.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f16",
        "data": "synthetic",
        "config": {"workload": WORKLOADS["cfg2"], "clips_per_gpu_per_step": b, "patches_per_gpu_per_step": n,
                   "l2": f"per-step input {n * 3 * p * p * 4 / 2**20:.0f} MiB/GPU exceeds the 126 MB L2; no flush"},
        "clocks": clocks, "e2e": e2e, "gpu_launches": plan.num_launches * a.steps,
        "roofline": {"bound": "tensor", "achieved": tfl, "peak": tf_peak, "unit": "TFLOP/s", "frac": tfl / tf_peak,
                     "traffic": (tr["fl_dram_bytes"] * n / tr["patches"]) if tr else None, "traffic_source": tr_src,
                     "kernel": "all fL launches of one step",
                     "how": f"{n} patches x {FL_GFLOP_PER_PATCH} GFLOP / {fl_ms:.3f} ms (marks inside the replay); "
                            f"peak = sustained bf16/fp16 dense, {peak_src}"},
    }
    if world == 1 and not a.no_torch_gpu_baseline:
        line["torch_gpu_baseline"] = torch_gpu_baseline("cfg2", b, dev)
    if world == 1 and not a.no_cpu_baseline:
        try:
            _, info = reference_cpu("cfg2", 3, 1, budget_s=25.0)
            line["cpu_baseline"] = {k: info[k] for k in ("value", "unit", "cores", "kind", "sample")}
        except Exception as exc:
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": f"failed: {exc}"}
    print(json.dumps(line), flush=True)
    h.finish()


# ------------------------------------------------------------------------------------------------ cfg5 (STH)
def run_cfg5(a):
    """BASELINE config 5: STH tree, T_g=8, T_f in {8,12,16}, P=144, ResNet-101 fL, C=174; videos sharded over the
    ranks, ONE all-gather of the (B,174) predictions.  `value` is the fused forward_eval plan (pred only);
    `value_with_baseline_pass` drives glance() + action_stage2() like STH/evaluate.py:188-201 (policy patches AND the
    random-patch baseline pass, i.e. fL twice)."""
    h = Harness(a)
    torch, dist, dev, world, rank = h.torch, h.dist, h.dev, h.world, h.rank
    from adafocus_b200 import synth
    from adafocus_b200.models_sth.gfv_net import GFV
    b = a.batch or 32
    tf_ = a.focuser_frames
    args = synth.sth_args(base_model="resnet101", batch_size=b, num_segments_focuser=tf_)
    model = GFV(args).to(dev)
    synth.strip_fc_sth(model)
    synth.load_checkpoint_sth(model, synth.synth_checkpoint_sth(model))
    model.focuser.policy.policy.to(dev).eval()
    model.focuser.policy.policy_old.to(dev).eval()
    model.eval()
    c = args.num_classes
    plan = model.eval_plan(args, b, dev)
    gen = torch.Generator(device=dev).manual_seed(synth.SEED + rank)
    plan.glancer_images.copy_(torch.randn(plan.glancer_images.shape, generator=gen, device=dev))
    plan.focuser_images.copy_(torch.randn(plan.focuser_images.shape, generator=gen, device=dev))
    gathered = torch.empty(world * b, c, device=dev) if world > 1 else None

    def step():
        plan.run()
        if world > 1:
            dist.all_gather_into_tensor(gathered, plan.pred.contiguous())

    sampler = ClockSampler(h.local) if rank == 0 else None
    if sampler:
        sampler.start()
    ms_total = h.timed(step, sampler)
    clocks = sampler.stop() if sampler else None
    ms = ms_total / a.steps
    stages = plan.stage_ms()
    value = world * b / (ms * 1e-3)

    gi, fi = plan.glancer_images, plan.focuser_images.view(b, tf_, 3, 224, 224)

    def step_eval_py():
        fmap, glogit = model.glance(gi)
        lp, pred = None, None
        for k in range(args.video_div):
            pred, _base, lp = model.action_stage2(fi, fmap, glogit, k, args, prev_local_patch=lp, training=False)
        if world > 1:
            dist.all_gather_into_tensor(gathered, pred.contiguous())
    ms2 = h.timed(step_eval_py, steps=max(3, a.steps // 4), warmup=3) / max(3, a.steps // 4)

    e2e = None
    if not a.no_e2e:
        hg = [torch.randn(plan.glancer_images.shape).pin_memory() for _ in range(2)]
        hf = [torch.randn(plan.focuser_images.shape).pin_memory() for _ in range(2)]
        plans = [plan, model.eval_plan(args, b, dev, slot=1)]
        out_host = [torch.empty(b, c).pin_memory() for _ in range(2)]
        copy_s, comp_s = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        copied = [torch.cuda.Event() for _ in range(2)]
        consumed = [torch.cuda.Event() for _ in range(2)]

        def run_stream(nb):
            for sl in range(2):
                consumed[sl].record(comp_s)
            for i in range(nb):
                sl = i % 2
                with torch.cuda.stream(copy_s):
                    copy_s.wait_event(consumed[sl])
                    plans[sl].glancer_images.copy_(hg[sl], non_blocking=True)
                    plans[sl].focuser_images.copy_(hf[sl], non_blocking=True)
                    copied[sl].record(copy_s)
                with torch.cuda.stream(comp_s):
                    comp_s.wait_event(copied[sl])
                    plans[sl].plan.run(comp_s.cuda_stream)
                    consumed[sl].record(comp_s)
                    out_host[sl].copy_(plans[sl].pred, non_blocking=True)
            comp_s.synchronize()
        run_stream(h.warmup)
        h.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run_stream(a.steps)
        if world > 1:
            dist.all_gather_into_tensor(gathered, plans[(a.steps - 1) % 2].pred.contiguous())
        e1.record()
        h.barrier()
        ems = h.max_over_ranks(e0.elapsed_time(e1)) / a.steps
        e2e = {"value": world * b / (ems * 1e-3), "unit": "videos/s",
               "h2d_bytes_per_step": world * (hg[0].numel() + hf[0].numel()) * 4, "d2h_bytes_per_step": world * b * c * 4,
               "ms_per_step": ems, "api": "pinned fp32 glancer + focuser frames -> GFV.eval_plan replay -> pinned pred"}
        del hg, hf
    if rank != 0:
        h.finish()
        return
    tf_peak, _, peak_src = measured_peaks()
    gfl = {8: 6.583, 12: 6.583, 16: 6.583}[tf_] if tf_ in (8, 12, 16) else R101_GFLOP_PER_PATCH_144
    tfl = b * tf_ * gfl * 1e9 / (stages["fL"] * 1e-3) / 1e12
    tr, tr_src = committed_traffic("cfg5")
    line = {
        "metric": METRICS["cfg5"], "value": value, "unit": "videos/s", "n_gpus": world, "steps": a.steps,
        "warmup": h.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f16", "data": "synthetic",
        "config": {"workload": WORKLOADS["cfg5"], "videos_per_gpu_per_step": b, "global_videos_per_step": b * world,
                   "glancer_frames": 8, "focuser_frames": tf_, "patch": 144,
                   "parallelism": f"dp{world} (videos sharded, weights replicated, one all-gather of (B,174) preds)"},
        "clocks": clocks, "stages_ms": stages, "gpu_launches": plan.plan.num_launches * a.steps, "e2e": e2e,
        "value_with_baseline_pass": {"value": world * b / (ms2 * 1e-3), "unit": "videos/s", "ms_per_step": ms2,
                                     "api": "GFV.glance + GFV.action_stage2 (evaluate.py's loop body; fL runs twice)"},
        "model_tflops": value / world * (8 * 0.599 + tf_ * gfl) / 1e3,
        "roofline": {"bound": "tensor", "achieved": tfl, "peak": tf_peak, "unit": "TFLOP/s", "frac": tfl / tf_peak,
                     "traffic": (tr["fl_dram_bytes"] * (b * tf_) / tr["patches"]) if tr else None,
                     "traffic_source": tr_src, "kernel": "all fL launches of one step (TSM-ResNet-101 @144^2)",
                     "how": f"{b * tf_} patches x {gfl} GFLOP / fL stage {stages['fL']:.3f} ms; peak = sustained "
                            f"bf16/fp16 dense, {peak_src}"},
    }
    if world == 1 and not a.no_torch_gpu_baseline:
        del plan
        model._plans.clear()
        torch.cuda.empty_cache()
        line["torch_gpu_baseline"] = torch_gpu_baseline("cfg5", b, dev, tf_)
    if world == 1 and not a.no_cpu_baseline:
        try:
            _, info = reference_cpu("cfg5", 3, 1, budget_s=25.0, focuser_frames=tf_)
            line["cpu_baseline"] = {k: info[k] for k in ("value", "unit", "cores", "kind", "sample")}
            line["cpu_baseline"]["unit"] = "videos/s"
        except Exception as exc:
            line["cpu_baseline"] = {"value": None, "unit": "videos/s", "cores": 0, "kind": "port",
                                    "sample": f"failed: {exc}"}
    print(json.dumps(line), flush=True)
    h.finish()


def run_cfg1(a):
    """BASELINE config 1: one 8-frame 224^2 clip, fG only -- the reference's own CPU-runnable plumbing case; here the
    same call (STH GFV.glance) on the GPU at batch 1 (latency) with the reference's CPU and GPU paths beside it."""
    h = Harness(a)
    torch, dev = h.torch, h.dev
    from adafocus_b200 import synth
    from adafocus_b200.models_sth.gfv_net import GFV
    b = a.batch or 1
    args = synth.sth_args(batch_size=b)
    model = GFV(args).to(dev)
    synth.strip_fc_sth(model)
    synth.load_checkpoint_sth(model, synth.synth_checkpoint_sth(model))
    model.eval()
    gi = synth.synth_clips(b, args.num_segments_glancer, 224, synth.SEED + 1).to(dev)
    host = gi.cpu().pin_memory()
    out_host = torch.empty(b, args.num_segments_glancer, args.num_classes).pin_memory()
    launches0 = None
    from adafocus_b200.engine import get_engine
    eng = get_engine(dev)

    def step():
        return model.glance(gi)
    step()
    launches0 = eng.launch_count
    step()
    per_step = eng.launch_count - launches0
    sampler = ClockSampler(h.local)
    sampler.start()
    ms = h.timed(step, sampler) / a.steps
    clocks = sampler.stop()

    def step_e2e():
        gi.copy_(host, non_blocking=True)
        fmap, logit = model.glance(gi)
        out_host.copy_(logit, non_blocking=True)
    ems = h.timed(step_e2e) / a.steps
    line = {
        "metric": METRICS["cfg1"], "value": b / (ms * 1e-3), "unit": UNIT, "n_gpus": 1, "steps": a.steps,
        "warmup": h.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f16", "data": "synthetic",
        "config": {"workload": WORKLOADS["cfg1"], "clips_per_step": b, "l2": "latency case: the working set fits the L2"},
        "clocks": clocks, "gpu_launches": per_step * a.steps,
        "e2e": {"value": b / (ems * 1e-3), "unit": UNIT, "h2d_bytes_per_step": host.numel() * 4,
                "d2h_bytes_per_step": out_host.numel() * 4, "ms_per_step": ems,
                "api": "pinned clip -> GFV.glance (adafocus_b200.models_sth) -> pinned per-frame logits"},
    }
    if not a.no_torch_gpu_baseline:
        line["torch_gpu_baseline"] = torch_gpu_baseline("cfg1", b, dev)
    if not a.no_cpu_baseline:
        try:
            _, info = reference_cpu("cfg1", 5, 2, clips_per_step=b, budget_s=25.0)
            line["cpu_baseline"] = {k: info[k] for k in ("value", "unit", "cores", "kind", "sample")}
        except Exception as exc:
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": f"failed: {exc}"}
    print(json.dumps(line), flush=True)
    h.finish()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    elif a.workload == "cfg1":
        run_cfg1(a)
    elif a.workload == "cfg5":
        run_cfg5(a)
    elif a.workload == "cfg2":
        run_cfg2(a)
    else:
        run_act(a)
