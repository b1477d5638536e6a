"""End-to-end evaluation from HOST buffers: the call a user of the reference's validate() loop makes.

`StreamingEvaluator.run(batches)` takes pinned host tensors in the reference's input contract -- (B, 3T, H, W) fp32,
ImageNet-normalised, as produced by the reference's DataLoader (ACT/ops/transforms.py:303-336) -- copies each batch
host->device, runs the fused stage-3 forward (ACT/main_dist.py:367-371) and copies the per-clip logits back to the
host.  Two input slots and two CUDA streams overlap the H2D copy of batch i+1 with the compute of batch i; the policy
loop, the crop and all ~190 layer launches replay natively from one recorded plan per slot.

`input_format="u8_hwc"` moves the last two steps of the reference's transform chain onto the device: the host batches
are then (B, H, W, 3T) uint8 -- the stacked decoded frames, before ToTorchFormatTensor / GroupNormalize
(ACT/ops/transforms.py:303-336, 64-77) -- a quarter of the PCIe bytes; `af_frames_u8_to_f32` produces the bit-identical
fp32 tensor in front of the plan.
"""
import torch

from .engine import get_engine


class StreamingEvaluator:
    def __init__(self, model, batch, device, slots=2, input_format="f32", frame_hw=None):
        assert input_format in ("f32", "u8_hwc", "u8_frames")
        self.model, self.batch, self.device = model, batch, torch.device(device)
        self.input_format = input_format
        t, s = model.num_segments, model.input_size
        self.plans = [model.fused_plan(batch, t, s, s, model.glance_size, self.device, True, slot=i)
                      for i in range(slots)]
        self.copy_stream = torch.cuda.Stream(self.device)
        self.compute_stream = torch.cuda.Stream(self.device)
        self.copied = [torch.cuda.Event() for _ in range(slots)]
        self.consumed = [torch.cuda.Event() for _ in range(slots)]
        self.num_classes = self.plans[0].num_classes
        self.t = t
        self.out_host = [torch.empty(batch, self.num_classes, dtype=torch.float32).pin_memory() for _ in range(slots)]
        self.h2d_bytes = batch * 3 * t * s * s * (4 if input_format == "f32" else 1)
        self.d2h_bytes = batch * self.num_classes * 4
        self.u8 = None
        self.raw = None
        if input_format in ("u8_hwc", "u8_frames"):
            self.eng = get_engine(self.device)
            self.u8 = [torch.empty(batch, s, s, 3 * t, dtype=torch.uint8, device=self.device) for _ in range(slots)]
            self.mean, self.std = list(model.input_mean), list(model.input_std)
        if input_format == "u8_frames":
            # decoded frames at their stored size (B*T, H, W, 3): GroupScale + GroupCenterCrop + Stack run on the device
            from .preprocess import FramePreprocessor
            if frame_hw is None:
                raise ValueError("input_format='u8_frames' needs frame_hw=(H, W) of the decoded frames")
            fh, fw = frame_hw
            self.pre = FramePreprocessor(model.scale_size, model.crop_size, self.mean, self.std, self.device)
            self.raw = [torch.empty(batch * t, fh, fw, 3, dtype=torch.uint8, device=self.device) for _ in range(slots)]
            self.h2d_bytes = batch * t * fh * fw * 3

    def _refresh_plans(self):
        """Plans are re-fetched on every run(): after a checkpoint reload the model hands back re-recorded plans over
        freshly packed weights (an O(1) key check per plan otherwise)."""
        m = self.model
        self.plans = [m.fused_plan(self.batch, self.t, m.input_size, m.input_size, m.glance_size, self.device, True,
                                   slot=i) for i in range(len(self.plans))]

    def run(self, host_batches, collect=True):
        """host_batches: sequence of pinned (B,3T,H,W) fp32 CPU tensors -- or (B,H,W,3T) uint8 with
        input_format="u8_hwc", or (B*T,Hf,Wf,3) uint8 decoded frames with input_format="u8_frames"; the last one may
        hold fewer than B clips.  Returns the list of (n_i, C) host logits
        (last time step, the reference's `pred`).  collect=True keeps every batch's result (rows of ONE pinned
        (num_batches, B, C) buffer allocated up front); collect=False recycles one pinned buffer per slot, for timing."""
        self._refresh_plans()
        host_batches = list(host_batches)
        results = []
        n = len(self.plans)
        if collect:
            ring = torch.empty(max(1, len(host_batches)), self.batch, self.num_classes, dtype=torch.float32).pin_memory()
        for s in range(n):
            self.consumed[s].record(self.compute_stream)
        for i, hb in enumerate(host_batches):
            s = i % n
            plan = self.plans[s]
            nb = hb.shape[0] // self.t if self.raw is not None else hb.shape[0]
            if nb > self.batch or (nb < self.batch and i != len(host_batches) - 1):
                raise ValueError(f"batch {i} has {nb} clips: every batch but the last must hold exactly {self.batch}")
            dst = plan.input if self.u8 is None else (self.u8[s] if self.raw is None else self.raw[s])
            with torch.cuda.stream(self.copy_stream):
                self.copy_stream.wait_event(self.consumed[s])        # slot's previous compute has read its input
                # ragged final batch: clips are independent, rows >= nb keep the slot's previous (valid) clips
                dst[:hb.shape[0]].copy_(hb, non_blocking=True)
                self.copied[s].record(self.copy_stream)
            with torch.cuda.stream(self.compute_stream):
                self.compute_stream.wait_event(self.copied[s])
                if self.raw is not None:
                    self.pre(self.raw[s], self.t, out=plan.input, u8_out=self.u8[s].view(self.batch * self.t, *self.u8[s].shape[1:3], 3))
                elif self.u8 is not None:
                    self.eng.frames_u8_to_f32(self.u8[s], self.mean, self.std, out=plan.input)
                plan.run()
                self.consumed[s].record(self.compute_stream)
                last = plan.logits.view(self.batch, self.t, -1)[:, -1, : self.num_classes]
                out = ring[i] if collect else self.out_host[s]
                out.copy_(last, non_blocking=True)
                results.append(out[:nb])
        self.compute_stream.synchronize()
        return results
