"""TEST / MEASUREMENT INFRASTRUCTURE ONLY -- builds and times the UNMODIFIED reference classes (via
oracle/reference_loader) on the seeded synthetic checkpoint of adafocus_b200.synth.

Used by bench.py (`--impl reference`, `cpu_baseline`, `torch_gpu_baseline`) and by the GPU tests that measure the
reference's own precision band.  The product never imports this."""
import os
import time

import torch

from adafocus_b200 import synth

from . import reference_loader as rl


def build_act(args, device="cpu"):
    """Reference ACT GFV (ACT/models/gfv_net.py:13) with the synthetic checkpoint loaded through the reference's own
    sequence (ACT/main_dist.py:100-110), eval mode."""
    ref = rl.import_act()
    torch.manual_seed(0)
    model = ref.GFV(args)
    ck = synth.synth_checkpoint_act(model, synth.SEED)
    synth.load_checkpoint_act(model, ck)
    model.eval()
    if torch.device(device).type == "cuda":
        model.cuda(device)
    return model, ck


def build_sth(args, device="cpu"):
    """Reference STH GFV driven like STH/evaluate.py (:83 fc strip, :141-146 loads, :175-177 eval of both policies)."""
    ref = rl.import_sth()
    torch.manual_seed(0)
    model = ref.GFV(args)
    synth.strip_fc_sth(model)
    ck = synth.synth_checkpoint_sth(model, synth.SEED)
    synth.load_checkpoint_sth(model, ck)
    model.eval()
    model.focuser.policy.policy.eval()
    model.focuser.policy.policy_old.eval()
    if torch.device(device).type == "cuda":
        model.cuda(device)
        model.focuser.policy.policy.cuda(device)
        model.focuser.policy.policy_old.cuda(device)
    return model, ck


def act_forward(model, x, glance_size, gpu=None):
    """The stage-3 call of ACT/main_dist.py:332,368."""
    scan = torch.nn.functional.interpolate(x, (glance_size, glance_size))
    return model(input=x, scan=scan, training=False, backbone_pred=False, one_step=True, gpu=gpu)


def sth_forward(model, gi, fi, args, with_baseline=True):
    """The per-batch body of STH/evaluate.py:188-201 (action_stage2, incl. the random-patch baseline pass) or the
    pred-only action_stage3 variant."""
    b = gi.shape[0]
    gin = torch.nn.functional.interpolate(gi, (args.glance_size, args.glance_size))
    fimg = fi.view(b, args.num_segments_focuser, 3, 224, 224)
    fmap, glogit = model.glance(gin)
    lp, pred = None, None
    for step in range(args.video_div):
        if with_baseline:
            pred, _base, lp = model.action_stage2(fimg, fmap, glogit, step, args, prev_local_patch=lp, training=False)
        else:
            pred, lp = model.action_stage3(fimg, fmap, glogit, step, args, prev_local_patch=lp)
    return pred


def _time_cpu(fn, steps, warmup):
    for _ in range(warmup):
        fn()
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    return ts


def pick_threads(fn, candidates):
    """n in {nproc, nproc/2}; keep the faster (one warm call each after a shared cold call)."""
    fn()                                  # cold start (oneDNN primitive creation), excluded
    best, best_t, seen = None, None, {}
    for n in candidates:
        torch.set_num_threads(n)
        fn()
        t0 = time.perf_counter()
        fn()
        seen[n] = time.perf_counter() - t0
        if best_t is None or seen[n] < best_t:
            best, best_t = n, seen[n]
    torch.set_num_threads(best)
    return best, seen


def time_gpu(fn, steps, warmup, device):
    """CUDA-event timing of `fn` (one step) on the current stream; returns ms per step."""
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize(device)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize(device)
    return e0.elapsed_time(e1) / steps


def host_info():
    cores = os.cpu_count() or 1
    model = ""
    try:
        with open("/proc/cpuinfo") as f:
            for ln in f:
                if ln.startswith("model name"):
                    model = ln.split(":", 1)[1].strip()
                    break
    except OSError:
        pass
    return cores, model
