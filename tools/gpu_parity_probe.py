"""Prints per-stage error statistics of the CUDA path against the oracle (CPU fp32) on the GPU box."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from adafocus_b200 import synth
from adafocus_b200.models.gfv_net import GFV
from oracle import adafocus_oracle as orc


def stats(name, got, ref):
    got, ref = got.float().cpu(), ref.float().cpu()
    err = (got - ref).abs()
    print(f"  {name:10s} max_abs_err {err.max():.4e} mean_abs_err {err.mean():.4e} ref_absmax {ref.abs().max():.4f} "
          f"ref_rms {ref.pow(2).mean().sqrt():.4f} rel_rms {(err.pow(2).mean().sqrt() / ref.pow(2).mean().sqrt()):.3e}",
          flush=True)


def main():
    torch.set_num_threads(os.cpu_count() or 8)
    b = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    over = {}
    args = synth.act_args(**over)
    dev = torch.device("cuda", 0)
    model = GFV(args)
    ck = synth.synth_checkpoint_act(model)
    synth.load_checkpoint_act(model, ck)
    model = model.to(dev)
    model.eval()
    x = synth.synth_clips(b, args.num_segments, args.input_size)
    t0 = time.time()
    ref = orc.act_forward(x, x, ck, args.patch_size, args.action_dim)
    print(f"oracle {time.time() - t0:.1f}s for {b} clips", flush=True)
    xd = x.to(dev)
    with torch.no_grad():
        t0 = time.time()
        logits, last = model(input=xd, scan=xd, training=False, backbone_pred=False, one_step=True, gpu=0)
        torch.cuda.synchronize()
        print(f"first forward (pack + plan build + run) {time.time() - t0:.2f}s, plan launches "
              f"{model.last_plan.plan.num_launches}, workspace {model.last_plan.plan.workspace_bytes / 2**20:.0f} MiB")
        plan = model.last_plan
        t = args.num_segments
        # stage: fG
        fmap, gvec = model.glance(xd)
        stats("fG map", fmap, ref["fmap"])
        stats("fG vec", gvec, ref["gvec"])
        # actions
        a = plan.action_idx.view(b, t).cpu().long()
        agree = (a == ref["actions"]).float().mean().item()
        print(f"  actions agree {agree:.4f}  ours {a.tolist()}\n                      ref  {ref['actions'].tolist()}")
        yx = plan.yx.view(b, t, 2).cpu().numpy()
        print("  coords equal where actions equal:",
              bool(np.array_equal(yx[(a == ref['actions']).numpy()], ref['coords'][(a == ref['actions']).numpy()])))
        # stage: fL on the ORACLE's patches (staged parity, independent of the policy)
        patches = ref["patches"].reshape(b * t, 3, args.patch_size, args.patch_size).to(dev)
        lf = model.focuser.net.get_featmap(patches, pooled=True).view(b, t, -1)
        stats("fL feat", lf, ref["lfeat"])
        # stage: classifier on the ORACLE's features
        lg, lo = model.classifier(ref["features"].to(dev))
        stats("head", lg, ref["logits"])
        # end to end
        stats("e2e logits", logits, ref["logits"])
        stats("e2e last", last, ref["last_out"])
        print("  class idx ours", last.argmax(1).tolist(), "ref", ref["last_out"].argmax(1).tolist())
        srt = ref["last_out"].sort(1)[0]
        print("  ref top1-top2 margin", (srt[:, -1] - srt[:, -2]).tolist())
        # forced actions: e2e with the oracle's actions -> isolates fp16 error from action flips
        for i in range(3):
            torch.cuda.synchronize()
            t0 = time.time()
            model(input=xd, scan=xd, training=False, backbone_pred=False, one_step=True, gpu=0)
            torch.cuda.synchronize()
            print(f"  forward {i}: {(time.time() - t0) * 1e3:.2f} ms for {b} clips")


if __name__ == "__main__":
    main()
