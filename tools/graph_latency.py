"""Latency of one fused forward replayed natively (af_plan_run) vs from a captured CUDA graph, per batch size."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from adafocus_b200 import synth
from adafocus_b200.models.gfv_net import GFV


def timeit(fn, iters=50):
    for _ in range(5):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    dev = torch.device("cuda", 0)
    args = synth.act_args()
    model = GFV(args)
    synth.load_checkpoint_act(model, synth.synth_checkpoint_act(model))
    model = model.to(dev)
    model.eval()
    for b in (1, 2, 8, 64):
        x = synth.synth_clips(b, args.num_segments, args.input_size).to(dev)
        plan = model.fused_plan(b, args.num_segments, args.input_size, args.input_size, model.glance_size, dev, True)
        plan.input.copy_(x)
        plan.run()
        torch.cuda.synchronize()
        ref = plan.logits.clone()
        t_eager = timeit(plan.run)
        plan.capture_graph()
        plan.logits.zero_()
        plan.run()
        torch.cuda.synchronize()
        same = torch.equal(ref, plan.logits)
        t_graph = timeit(plan.run)
        print(f"batch {b}: native replay {t_eager:.3f} ms, CUDA graph {t_graph:.3f} ms, identical logits: {same}", flush=True)


if __name__ == "__main__":
    main()
