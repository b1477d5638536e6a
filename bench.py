#!/usr/bin/env python
"""Benchmark of the AdaFocus offline-inference hot path (BASELINE.json config 3) on N B200s.

  python bench.py --gpus N --steps K --warmup W            # our CUDA path (one process per GPU under torchrun)
  python bench.py --impl reference --steps K --warmup W    # the reference algorithm on the host CPU (oracle port)

A "step" is one pass of the whole path -- fG over B*16 frames, 16-step policy rollout, crop, fL over B*16 patches, GRU
classifier -- over one batch of B synthetic clips per GPU (16 x 3 x 224 x 224 fp32 each).  `value` = clips/s with the
clips already resident in HBM; `e2e` = the same through adafocus_b200.pipeline.StreamingEvaluator with pinned HOST
buffers (H2D of the clips and D2H of the logits inside the timed region).  One JSON line on stdout (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "clips/sec (16-frame, 128^2 patch)"
UNIT = "clips/s"
WORKLOAD = ("cfg3: full AdaFocus (ACT tree) MobileNet-V2 fG + ResNet-50 fL + GRU policy/classifier, T=16, 224^2 "
            "frames, P=128, 49 actions, 200 classes, synthetic ActivityNet-shape clips")
FL_GFLOP_PER_PATCH = 2.669      # ResNet-50 trunk @128^2, 2*MAC
TOTAL_GFLOP_PER_CLIP = 53.04


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=64, help="clips per GPU per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--workload", default="cfg3", choices=["cfg3", "cfg5"],
                    help="cfg3 = the headline ACT configuration (default); cfg5 = Something-Something shape, "
                         "TSM-MobileNet-V2 (8 frames) + TSM-ResNet-101 (12 frames, 144^2 patches), extra line only")
    return ap.parse_args()


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return d.get("bf16_tflops_sustained", 1389.0), d.get("hbm_gbs", 6551.0), "measured (MEASURED_PEAKS.json)"
    return 1400.0, 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------ CPU reference
def cpu_reference_clips_per_sec(steps, warmup, clips_per_step=None, budget_s=None):
    """Times the oracle port of the reference's CPU path (oracle/adafocus_oracle.act_forward: GFV.forward(one_step=
    True) of ACT/models/gfv_net.py:95-133 in fp32) on the host cores.  Returns (clips/s, info dict)."""
    import torch
    from adafocus_b200 import synth
    from adafocus_b200.models.gfv_net import GFV
    from oracle import adafocus_oracle as orc
    cores = os.cpu_count() or 1
    threads = max(1, min(cores, 64))
    torch.set_num_threads(threads)
    args = synth.act_args()
    model = GFV(args)
    ck = synth.synth_checkpoint_act(model)
    del model
    x1 = synth.synth_clips(1, args.num_segments, args.input_size)
    t0 = time.perf_counter()
    orc.act_forward(x1, x1, ck, args.patch_size, args.action_dim)       # cold start (oneDNN primitive creation)
    cold = time.perf_counter() - t0
    t0 = time.perf_counter()
    orc.act_forward(x1, x1, ck, args.patch_size, args.action_dim)
    warm1 = time.perf_counter() - t0
    if clips_per_step is None:
        budget = budget_s or 150.0
        clips_per_step = int(max(1, min(8, budget / max(1e-3, warm1 * (steps + warmup)))))
    x = synth.synth_clips(clips_per_step, args.num_segments, args.input_size)
    for _ in range(warmup):
        orc.act_forward(x, x, ck, args.patch_size, args.action_dim)
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        orc.act_forward(x, x, ck, args.patch_size, args.action_dim)
        times.append(time.perf_counter() - t0)
    total = sum(times)
    value = clips_per_step * steps / total
    info = {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"{steps} timed steps x {clips_per_step} clip(s) of the cfg3 workload through "
                      f"oracle.act_forward (fp32, torch CPU ops, {threads} threads of {cores} host cores); "
                      f"cold first call {cold:.1f}s excluded",
            "ms_per_step": 1e3 * total / steps, "clips_per_step": clips_per_step}
    return value, info


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    value, info = cpu_reference_clips_per_sec(a.steps, max(1, a.warmup))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": info["ms_per_step"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "clips_per_step": info["clips_per_step"],
                   "note": "reference algorithm on host CPU cores; /root/reference is absent on the GPU box, so the "
                           "oracle port (pinned to the reference by tests/golden) is what runs"},
        "cpu_baseline": {k: info[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


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


# ------------------------------------------------------------------------------------------------ ours
def run_ours(a):
    import torch
    import torch.distributed as dist
    from adafocus_b200 import synth
    from adafocus_b200.models.gfv_net import GFV
    from adafocus_b200.pipeline import StreamingEvaluator

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a CUDA device: adafocus_b200 has no CPU path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    args = synth.act_args(batch_size=a.batch)
    model = GFV(args)
    synth.load_checkpoint_act(model, synth.synth_checkpoint_act(model))
    model = model.to(dev)
    model.eval()
    b, t, s = a.batch, args.num_segments, args.input_size
    c = args.num_classes

    # synthetic clips, generated on the device (seeded), already resident in HBM for the device-timed loop
    gen = torch.Generator(device=dev).manual_seed(synth.SEED + rank)
    inp, _ = model.input_buffers(b, dev)
    coarse = torch.randn(b * t, 3, 7, 7, generator=gen, device=dev)
    inp.copy_((torch.nn.functional.interpolate(coarse, size=(s, s)) +
               0.5 * torch.randn(b * t, 3, s, s, generator=gen, device=dev)).view(b, 3 * t, s, s))
    plan = model.fused_plan(b, t, s, s, args.glance_size, dev, True)
    gathered = torch.empty(world * b, c, device=dev) if world > 1 else None

    def step():
        plan.run()
        last = plan.logits.view(b, t, -1)[:, -1, :c]
        if world > 1:
            dist.all_gather_into_tensor(gathered, last.contiguous())     # one NCCL all-gather of per-clip logits
        return last

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    for _ in range(max(3, a.warmup)):
        step()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stage_acc = {}
    barrier()
    if sampler:
        sampler.region_begin()
    ev0.record()
    for _ in range(a.steps):
        step()
    ev1.record()
    barrier()
    if sampler:
        sampler.region_end()
    clocks = sampler.stop() if sampler else None
    ms_total = ev0.elapsed_time(ev1)
    stage_acc = plan.stage_ms()            # stages of the last timed step, from CUDA-event marks inside the plan
    tmax = torch.tensor([ms_total], device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms_total = float(tmax.item())
    ms_step = ms_total / a.steps
    value = world * b * a.steps / (ms_total * 1e-3)

    # ---- end to end from pinned host buffers (H2D + compute + D2H inside the timed region)
    e2e = None
    e2e_u8 = None

    def time_e2e(ev, host):
        nb = a.steps
        ev.run([host[i % 2] for i in range(max(3, a.warmup))], collect=False)
        barrier()
        t0 = time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        outs = ev.run([host[i % 2] for i in range(nb)], collect=False)
        if world > 1:
            dist.all_gather_into_tensor(gathered, ev.plans[(nb - 1) % 2].logits.view(b, t, -1)[:, -1, :c].contiguous())
        e1.record()
        barrier()
        wall = time.perf_counter() - t0
        e2e_ms = max(e0.elapsed_time(e1), 0.0)
        tm = torch.tensor([e2e_ms], device=dev)
        if world > 1:
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        e2e_ms = float(tm.item())
        del outs
        return {"value": world * b * nb / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": ev.h2d_bytes * world,
                "d2h_bytes_per_step": ev.d2h_bytes * world, "ms_per_step": e2e_ms / nb, "wall_s": wall}

    if not a.no_e2e:
        # (1) the reference's input contract: pinned fp32 (B,3T,H,W) batches, as its DataLoader produces them
        ev = StreamingEvaluator(model, b, dev, slots=2)
        host = [torch.empty(b, 3 * t, s, s, dtype=torch.float32).pin_memory() for _ in range(2)]
        for hbuf in host:
            hbuf.copy_(inp.cpu())
        e2e = time_e2e(ev, host)
        e2e["api"] = "adafocus_b200.pipeline.StreamingEvaluator.run (2 input slots, H2D overlapped with compute)"
        del host
        # (2) same clips shipped as stacked uint8 frames (B,H,W,3T), ToTorchFormatTensor + GroupNormalize on the device
        ev8 = StreamingEvaluator(model, b, dev, slots=2, input_format="u8_hwc")
        gen8 = torch.Generator().manual_seed(1007)
        host8 = [torch.randint(0, 256, (b, s, s, 3 * t), dtype=torch.uint8, generator=gen8).pin_memory()
                 for _ in range(2)]
        e2e_u8 = time_e2e(ev8, host8)
        e2e_u8["api"] = ("StreamingEvaluator(input_format='u8_hwc'): uint8 frames over PCIe, af_frames_u8_to_f32 "
                         "(bit-identical to the reference's transform chain) in front of the plan")
        del host8

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
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

    tf_peak, hbm_peak, peak_src = measured_peaks()
    fl_ms = stage_acc["fL"]
    fl_tflops = b * t * FL_GFLOP_PER_PATCH * 1e9 / (fl_ms * 1e-3) / 1e12
    roofline = {
        "bound": "tensor", "achieved": fl_tflops, "peak": tf_peak, "unit": "TFLOP/s", "frac": fl_tflops / tf_peak,
        # dram__bytes_read.sum + dram__bytes_write.sum over the 56 fL launches of one step at 64 clips/GPU
        # (profiles/r1_v15_launches_dram_b64.csv, launches 74-129 of a replay); scaled linearly with the number of patches
        "traffic": 16.71e9 * (b * t) / 1024.0,
        "kernel": "conv_gemm_kernel (tcgen05 implicit-GEMM), all fL launches of one step",
        "how": f"{b * t} patches x {FL_GFLOP_PER_PATCH} GFLOP (ResNet-50 trunk @128^2) / fL stage time {fl_ms:.3f} ms "
               f"measured by CUDA-event marks inside the timed plan replay (includes stem staging, maxpool, avgpool); "
               f"peak = sustained bf16/fp16 dense, {peak_src}",
    }
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": max(3, a.warmup),
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f16",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "clips_per_gpu_per_step": b, "global_clips_per_step": b * world,
                   "parallelism": f"dp{world} (clips sharded, weights replicated, one all-gather of (B,200) logits)",
                   "l2": f"per-step input {b * 3 * t * s * s * 4 / 2**20:.0f} MiB/GPU exceeds the 126 MB L2; no flush"},
        "clocks": clocks, "e2e": e2e, "e2e_u8_frames": e2e_u8, "gpu_launches": plan.plan.num_launches * a.steps,
        "roofline": roofline, "stages_ms": stage_acc,
        "crop_roofline": {"bound": "hbm", "achieved": crop_bytes / (crop_ms * 1e-3) / 1e9, "peak": hbm_peak,
                          "unit": "GB/s", "frac": crop_bytes / (crop_ms * 1e-3) / 1e9 / hbm_peak,
                          "kernel": "crop_nchw_f32_vec4_kernel (get_patch, fp32 -> fp32)",
                          "how": f"{b * t} patches x 393216 B (read + write) / {crop_ms * 1e3:.1f} us (CUDA events, "
                                 f"{reps} launches, source frames {b * 3 * t * s * s * 4 / 2**20:.0f} MiB > L2)"},
        "model_tflops": value * TOTAL_GFLOP_PER_CLIP / 1e3 / world,
    }
    if world == 1 and not a.no_cpu_baseline:
        try:
            _, info = cpu_reference_clips_per_sec(3, 1, budget_s=25.0)
            line["cpu_baseline"] = {k: info[k] for k in ("value", "unit", "cores", "kind", "sample")}
        except Exception as exc:   # the baseline is a report, never the product path
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": f"failed: {exc}"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_cfg5(a):
    """Secondary workload (BASELINE config 5 shape on one GPU): STH tree, T_g=8, T_f=12, P=144, ResNet-101 fL, C=174,
    fused forward_eval plan, device-resident inputs.  Not the headline metric; prints its own JSON line."""
    import torch
    from adafocus_b200 import synth
    from adafocus_b200.models_sth.gfv_net import GFV
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    args = synth.sth_args(base_model="resnet101", batch_size=a.batch)
    model = GFV(args).to(dev)
    synth.strip_fc_sth(model)
    synth.load_checkpoint_sth(model, synth.synth_checkpoint_sth(model))
    model.focuser.policy.policy.to(dev).eval()
    model.focuser.policy.policy_old.to(dev).eval()
    model.eval()
    b = a.batch
    plan = model.eval_plan(args, b, dev)
    gen = torch.Generator(device=dev).manual_seed(synth.SEED)
    plan.glancer_images.copy_(torch.randn(plan.glancer_images.shape, generator=gen, device=dev))
    plan.focuser_images.copy_(torch.randn(plan.focuser_images.shape, generator=gen, device=dev))
    for _ in range(max(3, a.warmup)):
        plan.run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        plan.run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    stages = plan.stage_ms()
    gflop_clip = 8 * 0.599 + 12 * 6.583
    print(json.dumps({
        "metric": "videos/sec (Sth-Sth shape, 8+12 frames, 144^2 patch, ResNet-101 fL)", "value": b / (ms * 1e-3),
        "unit": "videos/s", "n_gpus": 1, "steps": a.steps, "warmup": max(3, a.warmup), "ms_per_step": ms,
        "higher_is_better": True, "dtype": "f16", "data": "synthetic",
        "config": {"workload": "cfg5 shape on 1 GPU: STH tree forward_eval (pred only)", "videos_per_step": b},
        "stages_ms": stages, "gpu_launches": plan.plan.num_launches * a.steps,
        "model_tflops": b / (ms * 1e-3) * gflop_clip / 1e3,
        "roofline": {"bound": "tensor", "achieved": b * 12 * 6.583e9 / (stages["fL"] * 1e-3) / 1e12,
                     "peak": measured_peaks()[0], "unit": "TFLOP/s",
                     "frac": b * 12 * 6.583e9 / (stages["fL"] * 1e-3) / 1e12 / measured_peaks()[0], "traffic": None},
    }), flush=True)


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    elif a.workload == "cfg5":
        run_cfg5(a)
    else:
        run_ours(a)
