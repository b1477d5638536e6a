"""Where does the logit error of the CUDA path come from?  (GPU box; prints one JSON object.)

  e2e            : our logits vs the oracle (CPU fp32)
  trunk_only     : OUR [global | local] features (plan.feat) pushed through the ORACLE's fp32 GRU classifier
  head_only      : the ORACLE's fp32 features pushed through OUR classifier kernels
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from adafocus_b200 import synth  # noqa: E402
from adafocus_b200.models.gfv_net import GFV  # noqa: E402
from oracle import adafocus_oracle as orc  # noqa: E402


def err(got, ref):
    got, ref = got.double().cpu(), ref.double().cpu()
    d = (got - ref).abs()
    scale = max(1.0, float(ref.abs().max()))
    return {"max_abs": float(d.max()), "max_abs_over_scale": float(d.max()) / scale,
            "rel_rms": float(((got - ref).pow(2).mean().sqrt() / ref.pow(2).mean().sqrt())), "scale": scale}


def main():
    torch.set_num_threads(os.cpu_count() or 8)
    dev = torch.device("cuda", 0)
    out = {}
    for tag, over, b in (("c3_b2", {}, 2), ("t4_p96_b3", dict(num_segments=4, patch_size=96, action_dim=36,
                                                              num_classes=51), 3)):
        args = synth.act_args(**over)
        model = GFV(args)
        ck = synth.synth_checkpoint_act(model)
        synth.load_checkpoint_act(model, ck)
        model = model.to(dev)
        model.eval()
        x = synth.synth_clips(b, args.num_segments, args.input_size)
        ref = orc.act_forward(x, x, ck, args.patch_size, args.action_dim)
        xd = x.to(dev)
        logits, last = model(input=xd, scan=xd, training=False, backbone_pred=False, one_step=True, gpu=0)
        torch.cuda.synchronize()
        plan = model.last_plan
        t = args.num_segments
        r = {"e2e": err(logits, ref["logits"])}
        feat = plan.features().float().cpu().view(b, t, -1)
        r["features"] = err(feat, ref["features"])
        lg, _ = orc.recurrent_classifier(feat, ck["fc"])
        r["trunk_only"] = err(lg, ref["logits"])
        lg2, _ = model.classifier(ref["features"].to(dev))
        r["head_only"] = err(lg2, ref["logits"])
        out[tag] = r
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
