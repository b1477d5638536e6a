"""CPU-only checks: the C-ABI library loads and exports every symbol include/adafocus_b200.h declares, fails loudly
without a GPU, and the host-side logic (weight packing layout, action tables, synthetic checkpoint, parameter names)."""
import json
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from adafocus_b200 import _lib, build
    build.build()
    header = open(os.path.join(ROOT, "include", "adafocus_b200.h")).read()
    declared = set(re.findall(r"\b(af_[a-z0-9_]+)\s*\(", header))
    declared -= {"af_status", "af_act"}
    lib = _lib.load()
    assert declared == set(_lib.SIGNATURES), (declared ^ set(_lib.SIGNATURES))
    for name in declared:
        assert hasattr(lib, name)
    assert lib.af_version() == 210


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    from adafocus_b200 import _lib
    from adafocus_b200.engine import Engine
    with pytest.raises(_lib.AfError):
        _lib.Context(0)
    with pytest.raises(_lib.AfError):
        Engine("cpu")
    from adafocus_b200.models.utils import get_patch
    with pytest.raises(RuntimeError):
        get_patch(torch.zeros(1, 3, 8, 8), torch.zeros(1, 2), 4)


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "adafocus_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f


def test_pack_conv_layout():
    from adafocus_b200.engine import pack_conv, pack_stem
    w = torch.arange(2 * 3 * 3 * 3, dtype=torch.float32).reshape(2, 3, 3, 3) / 64
    w8 = torch.zeros(2, 8, 3, 3)
    w8[:, :3] = w
    pc = pack_conv(w8, torch.tensor([2.0, 3.0]), torch.tensor([0.5, -0.5]), stride=2, pad=1, act=1, block_n=16,
                   device="cpu")
    assert pc.w.shape == (16, 9 * 64) and pc.w.dtype == torch.float16
    for co in range(2):
        for r in range(3):
            for s in range(3):
                for ci in range(3):
                    assert float(pc.w[co, (r * 3 + s) * 64 + ci]) == float(w[co, ci, r, s])
    assert float(pc.w[2:].abs().max()) == 0 and float(pc.w[:, 8:64].abs().max()) == 0
    assert pc.scale[:2].tolist() == [2.0, 3.0] and pc.scale[2:].eq(1).all() and pc.bias[:2].tolist() == [0.5, -0.5]
    ps = pack_stem(w, None, None, stride=2, pad=1, act=2, device="cpu")
    assert ps.w.shape == (16, 64) and ps.stem["kpad"] == 32
    assert float(ps.w[1, (1 * 3 + 2) * 3 + 1]) == float(w[1, 1, 1, 2])
    lin = torch.randn(5, 12)
    perm = torch.arange(12).flip(0)
    pl = pack_conv(lin, None, torch.ones(5), device="cpu", cin_perm=perm)
    assert torch.equal(pl.w[:5, :12], lin[:, perm].half())


@pytest.mark.parametrize("k,pad,p,cout", [(7, 3, 32, 64), (3, 1, 24, 32), (5, 2, 20, 16)])
def test_stem_space_to_depth_packing_equals_strided_conv(k, pad, p, cout):
    """Host-side half of the tensor-core stems: the weights pack_stem_s2d produces, applied as a stride-1 R x 1 conv
    over the sliding-window view of the 2x2 space-to-depth image (emulated here with torch ops exactly as
    af_stem_s2d lays it out), equal the stride-2 k x k convolution (ACT/models/resnet.py:138, mobilenet.py:105)."""
    import torch.nn.functional as F
    from adafocus_b200.engine import pack_stem
    torch.manual_seed(k)
    w = torch.randn(cout, 3, k, k)
    x = torch.randn(2, 3, p, p)
    q = pack_stem(w, None, None, 2, pad, 0, device="cpu").s2d
    assert q is not None and q.cin == 64 and q.kw == 1 and q.kh * q.vt == (k + 1) // 2
    ref = F.conv2d(x, w, stride=2, padding=pad)
    ho = ref.shape[2]
    pe, win = 16 * q.vt, 64 // (16 * q.vt)            # channels per s2d pixel, X positions per view pixel
    hs, ws = ho + q.kh - 1, ho + win - 1
    xp = torch.zeros(2, 3, 2 * (hs + q.vt), 2 * (ws + 1))
    xp[:, :, pad:pad + p, pad:pad + p] = x
    s2d = torch.zeros(2, hs, ws, pe)
    for v in range(q.vt):
        for dy in range(2):
            for dx in range(2):
                for c in range(3):
                    s2d[..., v * 16 + (dy * 2 + dx) * 3 + c] = xp[:, c, dy + 2 * v::2, dx::2][:, :hs, :ws]
    view = torch.cat([s2d[:, :, sx:sx + ho, :] for sx in range(win)], dim=3)          # (2, hs, ho, 64)
    wq = q.w.float().reshape(-1, q.kh, 64)[:cout]
    out = sum(torch.einsum("nhwc,oc->nohw", view[:, r:r + ho], wq[:, r]) for r in range(q.kh))
    assert float((out - ref).abs().max()) <= 2e-2 * float(ref.abs().max())            # fp16 weight rounding only


def test_fused_block_packing_carries_the_expand_bias_in_two_weight_columns():
    """pack_mbconv: BN scales folded into fp16 weights; with cin + 2 <= 64 the expand bias is stored as an fp16
    high part + fp16 remainder in weight columns cin and cin + 1 (they multiply the constant-1 channel pair the kernel
    plants in the input tile), reproducing the fp32 bias to ~2^-22."""
    from adafocus_b200.engine import pack_mbconv
    torch.manual_seed(5)
    cin, cexp, cout = 24, 144, 32
    w1, wd, w2 = torch.randn(cexp, cin), torch.randn(cexp, 1, 3, 3), torch.randn(cout, cexp)
    s1, b1 = torch.rand(cexp) + 0.5, torch.randn(cexp)
    s2, b2 = torch.rand(cexp) + 0.5, torch.randn(cexp)
    s3, b3 = torch.rand(cout) + 0.5, torch.randn(cout)
    pm = pack_mbconv(w1, s1, b1, wd, s2, b2, w2, s3, b3, 1, device="cpu")
    assert pm.bias1_in_w1 and pm.w1.shape == (192, 64) and pm.w1.dtype == torch.float16
    assert torch.equal(pm.w1[:cexp, :cin], (w1 * s1[:, None]).half())
    rebuilt = pm.w1[:cexp, cin].float() + pm.w1[:cexp, cin + 1].float()
    assert float((rebuilt - b1).abs().max()) <= 2 ** -20 * float(b1.abs().max())
    assert float(pm.w1[cexp:].abs().max()) == 0 and float(pm.w1[:, cin + 2:].abs().max()) == 0
    assert torch.allclose(pm.dw[:, :cexp].t().reshape(cexp, 1, 3, 3), wd * s2.view(-1, 1, 1, 1))
    wide = pack_mbconv(torch.randn(384, 64), torch.ones(384), torch.zeros(384), torch.randn(384, 1, 3, 3), torch.ones(384),
                       torch.zeros(384), torch.randn(64, 384), torch.ones(64), torch.zeros(64), 1, device="cpu")
    assert not wide.bias1_in_w1          # no spare K columns at cin = 64: bias stays in the epilogue


@pytest.mark.parametrize("cexp,spr", [(96, 4), (144, 4), (192, 2), (384, 1), (32, 4), (96, 2), (48, 2), (16, 4)])
def test_row_kernel_lane_placement(cexp, spr):
    """af_mbconv_rows_layout (host query, no device work): every expanded channel sits on at least one TMEM lane, a
    channel's K column is the same wherever it is replicated, and K columns inside a chunk are distinct per channel."""
    from adafocus_b200.engine import mbconv_rows_layout
    lay = mbconv_rows_layout(cexp, spr)
    assert lay is not None
    nch, lane_ch, lane_kpos = lay
    assert 1 <= nch <= 3
    seen = set()
    for c in range(nch):
        ch, kp = lane_ch[c].tolist(), lane_kpos[c].tolist()
        col_of = {}
        for lane in range(128):
            if ch[lane] >= 0:
                assert 0 <= kp[lane] < 128
                assert col_of.setdefault(ch[lane], kp[lane]) == kp[lane]          # replicas share the column
                seen.add(ch[lane])
        assert len(set(col_of.values())) == len(col_of)                           # distinct channels, distinct columns
        dead = [kp[lane] for lane in range(128) if ch[lane] < 0]
        assert not (set(dead) & set(col_of.values()))                             # idle lanes never alias a live column
    for c in range(nch, 3):
        assert all(v == -1 for v in lane_ch[c].tolist()) or nch == 3
    assert seen == set(range(cexp))
    assert mbconv_rows_layout(576, 1) is None and mbconv_rows_layout(192, 4) is None and mbconv_rows_layout(24, 4) is None


@pytest.mark.parametrize("cin,cexp,cout,stride,spr", [(24, 144, 24, 1, 4), (32, 96, 24, 2, 4), (64, 384, 64, 1, 1)])
def test_row_kernel_packing_reproduces_the_block(cin, cexp, cout, stride, spr):
    """pack_mbconv_rows on the host: evaluating the block FROM THE PACKED TENSORS (lane-ordered expand weights with the
    1/6 of the saturating ReLU6 folded in, per-lane depthwise taps and biases, K-column-ordered project weights with the
    6) gives the three reference convolutions (ACT/models/mobilenet.py:42-68) up to fp16 weight rounding."""
    import torch.nn.functional as F
    from adafocus_b200.engine import mbconv_rows_layout, pack_mbconv_rows
    torch.manual_seed(cexp)
    w1, wd, w2 = torch.randn(cexp, cin) / cin ** 0.5, torch.randn(cexp, 1, 3, 3) / 3, torch.randn(cout, cexp) / cexp ** 0.5
    s1, b1 = torch.rand(cexp) + 0.5, torch.randn(cexp) * 0.2
    s2, b2 = torch.rand(cexp) + 0.5, torch.randn(cexp) * 0.2
    s3, b3 = torch.rand(cout) + 0.5, torch.randn(cout) * 0.2
    pr = pack_mbconv_rows(w1, s1, b1, wd, s2, b2, w2, s3, b3, stride, spr, device="cpu")
    nch, lane_ch, lane_kpos = mbconv_rows_layout(cexp, spr)
    assert pr.w1.shape == (nch * 128, 64) and pr.dwp.shape == (nch, 11, 128) and pr.w2.shape[1] == nch * 128
    x = torch.randn(2, cin, 12, 12).half().float()
    out = torch.zeros(2, pr.w2.shape[0], (12 - 1) // stride + 1, (12 - 1) // stride + 1)
    done = set()
    for c in range(nch):
        for lane in range(128):
            ch = int(lane_ch[c, lane])
            if ch < 0 or ch in done:
                continue
            done.add(ch)
            e = F.conv2d(x, pr.w1[c * 128 + lane, :cin].float().view(1, cin, 1, 1)) + pr.dwp[c, 9, lane]
            e = e.clamp(0, 1)                                                   # add.sat: ReLU6(x) / 6
            d = F.conv2d(e, pr.dwp[c, :9, lane].view(1, 1, 3, 3), None, stride, 1) + pr.dwp[c, 10, lane]
            d = d.clamp(0, 1)
            col = c * 128 + int(lane_kpos[c, lane])
            out += d * pr.w2[:, col].float().view(1, -1, 1, 1)
    out = out[:, :cout] + pr.b3[:cout].view(1, -1, 1, 1)
    e = (F.conv2d(x, (w1 * s1[:, None])[:, :, None, None]) + b1.view(1, -1, 1, 1)).clamp(0, 6)
    d = (F.conv2d(e, wd * s2.view(-1, 1, 1, 1), None, stride, 1, 1, cexp) + b2.view(1, -1, 1, 1)).clamp(0, 6)
    ref = F.conv2d(d, (w2 * s3[:, None])[:, :, None, None]) + b3.view(1, -1, 1, 1)
    assert float((out - ref).abs().max()) <= 4e-3 * max(1.0, float(ref.abs().max()))
    assert float(pr.w2[cout:].abs().max() if pr.w2.shape[0] > cout else 0) == 0


def test_row_kernel_choice_heuristic(monkeypatch):
    """models/mobilenet.py:_rows_choice on a 148-SM device: whole frames per CTA, so the row kernel is taken from one
    frame per SM, skipped just above a multiple of the SM count, never for the 144-channel stride-1 block (28 warps at 72
    registers spill) and never when switched off."""
    import types
    import adafocus_b200.models.mobilenet as mb
    from adafocus_b200.engine import PackedMbconvRows
    eng = types.SimpleNamespace(ctx=types.SimpleNamespace(sm_count=148))

    def entry(cin, cexp, cout, stride):
        packs = {spr: PackedMbconvRows(None, None, None, None, cin, cexp, cout, stride, spr) for spr in (1, 2, 4)}
        return {"rows": packs, "stride": stride}

    def y(n, hw, c):
        return types.SimpleNamespace(shape=(n, hw, hw, c))

    monkeypatch.setattr(mb, "mbconv_rows_supported", lambda *a: True)
    monkeypatch.setattr(mb, "_ROWS_MODE", "auto")
    b5 = entry(32, 192, 32, 1)
    assert mb._rows_choice(eng, b5, y(1024, 28, 32)) is b5["rows"][2]
    assert mb._rows_choice(eng, b5, y(147, 28, 32)) is None              # fewer frames than SMs
    assert mb._rows_choice(eng, b5, y(148, 28, 32)) is not None
    assert mb._rows_choice(eng, b5, y(150, 28, 32)) is None              # two rounds for 1.01 frames per SM
    assert mb._rows_choice(eng, b5, y(256, 28, 32)) is not None          # cfg5
    b2 = entry(32, 96, 24, 2)
    assert mb._rows_choice(eng, b2, y(74, 112, 32)) is b2["rows"][4]     # 112-wide stride 2: two half-row units per frame
    b3 = entry(24, 144, 24, 1)
    assert mb._rows_choice(eng, b3, y(1024, 56, 24)) is None
    assert mb._rows_choice(eng, entry(24, 144, 32, 2), y(1024, 56, 24)) is not None
    assert mb._rows_choice(eng, b5, types.SimpleNamespace(shape=(1024, 28, 30, 32))) is None      # square frames only
    monkeypatch.setattr(mb, "_ROWS_MODE", "force")
    assert mb._rows_choice(eng, b3, y(2, 56, 24)) is b3["rows"][4]
    monkeypatch.setattr(mb, "_ROWS_MODE", "0")
    assert mb._rows_choice(eng, b5, y(1024, 28, 32)) is None


def test_standard_action_table_matches_reference_literals():
    from adafocus_b200.models.gfv_net import standard_action_table
    t49 = standard_action_table(49)
    assert t49.shape == (49, 2) and t49.dtype == torch.float32
    assert t49[8].tolist() == [np.float32(1 / 6), np.float32(1 / 6)]      # row-major: index = iy*7 + ix
    assert t49[6].tolist() == [0.0, 1.0] and t49[42].tolist() == [1.0, 0.0]
    assert standard_action_table(25)[7].tolist() == [np.float32(1 / 4), np.float32(2 / 4)]
    with pytest.raises(ValueError):
        standard_action_table(50)


def test_parameter_names_match_reference(golden_dir):
    """State-dict keys of the mirror == keys of the reference's modules (list written by make_golden.py)."""
    from adafocus_b200 import synth
    from adafocus_b200.models.gfv_net import GFV
    ref = json.load(open(os.path.join(golden_dir, "ref_state_keys.json")))
    m = GFV(synth.act_args())
    ours = {"glancer": m.glancer.state_dict(), "focuser": m.focuser.state_dict(), "fc": m.classifier.state_dict(),
            "policy": m.focuser.policy.policy.state_dict(), "model": m.state_dict()}
    for part, sd in ours.items():
        assert list(sd.keys()) == [k for k, _ in ref[part]], part
        assert [list(v.shape) for v in sd.values()] == [s for _, s in ref[part]], part


def test_synthetic_checkpoint_is_deterministic():
    from adafocus_b200 import synth
    from adafocus_b200.models.gfv_net import GFV
    args = synth.act_args(num_segments=2, num_classes=10)
    a = synth.synth_checkpoint_act(GFV(args))
    b = synth.synth_checkpoint_act(GFV(args))
    for part in ("glancer", "focuser", "fc", "policy"):
        for k in a[part]:
            assert torch.equal(a[part][k], b[part][k]), (part, k)
    x1, x2 = synth.synth_clips(1, 2), synth.synth_clips(1, 2)
    assert torch.equal(x1, x2) and x1.shape == (1, 6, 224, 224)
    # fingerprint pins the generator across machines / torch builds (golden vectors depend on it)
    assert abs(float(a["fc"]["fc.weight"].double().sum()) - float(b["fc"]["fc.weight"].double().sum())) == 0.0


def test_gfv_quirks_and_out_of_scope_paths():
    from adafocus_b200 import synth
    from adafocus_b200.models.gfv_net import GFV
    m = GFV(synth.act_args(num_segments=2, num_classes=10))
    assert m.eval() is None and m.training is False
    assert m.scale_size == 256 and m.crop_size == 224 and m.input_mean == [0.485, 0.456, 0.406]
    assert m.classifier.input_dim == 1280 + 2048
    x = torch.zeros(1, 6, 224, 224)
    with pytest.raises(NotImplementedError):
        m(input=x, scan=x, backbone_pred=True, one_step=False, glancer=True)
    with pytest.raises(RuntimeError):
        m(input=x, scan=x, training=False, backbone_pred=False, one_step=True)     # CPU tensors: no fallback


def test_sth_parameter_names_match_reference(golden_dir):
    """STH tree: state-dict keys of the mirror (after the fc strip of STH/evaluate.py:83) == the reference's."""
    from adafocus_b200 import synth
    from adafocus_b200.models_sth.gfv_net import GFV as GFV_STH
    ref = json.load(open(os.path.join(golden_dir, "ref_state_keys_sth.json")))
    m = GFV_STH(synth.sth_args())
    synth.strip_fc_sth(m)
    ours = {"glancer": m.glancer.state_dict(), "focuser": m.focuser.state_dict(), "fc": m.classifier.state_dict(),
            "policy": m.focuser.policy.policy.state_dict(), "model": m.state_dict()}
    for part, sd in ours.items():
        assert list(sd.keys()) == [k for k, _ in ref[part]], part
        assert [list(v.shape) for v in sd.values()] == [s for _, s in ref[part]], part
    # ResNet-101: TSM on every second block of every stage (n_round = 2, STH/ops/temporal_shift.py:124-135)
    m101 = GFV_STH(synth.sth_args(base_model="resnet101"))
    shifted = [k for k in m101.focuser.state_dict() if k.endswith("conv1.net.weight")]
    assert len(shifted) == 2 + 2 + 12 + 2
    assert m.eval() is None


def test_checkpoint_ingest_formats(tmp_path):
    """Reference checkpoint formats round-trip through adafocus_b200.checkpoint (ACT resume dict; STH resume dict;
    TSM 'module.base_model.' / 'module.new_fc.' pretrained backbones)."""
    from adafocus_b200 import checkpoint, synth
    from adafocus_b200.models.gfv_net import GFV
    from adafocus_b200.models_sth.gfv_net import GFV as GFV_STH
    m = GFV(synth.act_args(num_segments=2, num_classes=10))
    ck = synth.synth_checkpoint_act(m, seed=3)
    path = tmp_path / "act.pth.tar"
    checkpoint.save_checkpoint(ck, path)
    assert not os.path.exists(str(path) + ".tmp")
    info = checkpoint.load_act_checkpoint(m, str(path))
    assert info["epoch"] == 0
    assert torch.equal(m.classifier.fc.weight, ck["fc"]["fc.weight"])
    assert torch.equal(m.focuser.policy.policy_old.actor[0].weight, ck["policy"]["actor.0.weight"])

    s = GFV_STH(synth.sth_args(num_classes=16))
    # a TSM-style training checkpoint of the backbones: DataParallel 'module.' prefix, 'new_fc' heads
    g_sd = {"module.base_model." + k: v.clone() + 1 for k, v in s.glancer.net.state_dict().items()
            if not k.startswith("classifier.")}
    g_sd["module.new_fc.weight"] = torch.full_like(s.glancer.net.classifier.weight, 0.5)
    g_sd["module.new_fc.bias"] = torch.zeros_like(s.glancer.net.classifier.bias)
    f_sd = {"module." + k: v.clone() + 2 for k, v in s.focuser.net.state_dict().items()}
    f_sd["module.new_fc.weight"] = torch.full_like(s.classifier.weight, 0.25)
    f_sd["module.new_fc.bias"] = torch.ones_like(s.classifier.bias)
    before = s.focuser.net.base_model.conv1.weight.clone()
    checkpoint.load_sth_pretrained(s, {"state_dict": g_sd}, {"state_dict": f_sd})
    assert float(s.glancer.net.classifier.weight.mean()) == 0.5
    assert float(s.classifier.weight.mean()) == 0.25 and float(s.classifier.bias.mean()) == 1.0
    assert torch.equal(s.focuser.net.base_model.conv1.weight, before + 2)
    synth.strip_fc_sth(s)
    ck_s = synth.synth_checkpoint_sth(s, seed=5)
    checkpoint.load_sth_checkpoint(s, ck_s)
    assert torch.equal(s.classifier.weight, ck_s["fc"]["weight"])


def test_packing_is_host_arithmetic_and_split_layout():
    """engine.pack_* work on host tensors (no device kernels: a CUDA-less machine can pack) and pack_conv_split lays
    out [W_hi | W_hi | W_lo] with W_hi + W_lo == W to ~2^-22."""
    from adafocus_b200.engine import fold_bn, pack_conv, pack_conv_split
    torch.manual_seed(1)
    w = torch.randn(40, 128)
    pc = pack_conv_split(w, torch.randn(40), device="cpu")
    assert pc.w.shape == (48, 384) and pc.cin == 384 and pc.w.dtype == torch.float16
    hi, hi2, lo = pc.w[:40, :128].float(), pc.w[:40, 128:256].float(), pc.w[:40, 256:].float()
    assert torch.equal(hi, hi2) and torch.equal(hi, w.half().float())
    assert float((hi + lo - w).abs().max()) <= 2.0 ** -21 * float(w.abs().max())
    bn = torch.nn.BatchNorm2d(8)
    bn.running_var.uniform_(0.5, 2)
    s, b = fold_bn(bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.eps)
    assert s.device.type == "cpu" and b.device.type == "cpu"
    pc = pack_conv(torch.randn(8, 3, 3, 3), s, b, device="cpu")
    assert pc.w.shape == (16, 9 * 64)


def test_param_key_is_constant_time_and_tracks_reloads():
    """The packed-weight cache key is O(1) (epoch + three probe tensors) and is bumped by load_state_dict on the module,
    an ancestor or a descendant, by in-place updates of the probe tensors, and by invalidate_packed()."""
    from adafocus_b200 import synth
    from adafocus_b200.models.gfv_net import GFV
    from adafocus_b200.models.mobilenet import _param_key, invalidate_packed
    m = GFV(synth.act_args(num_segments=2, num_classes=10))
    net = m.focuser.net
    k0 = _param_key(net)
    assert _param_key(net) == k0 and len(k0) <= 8
    net.load_state_dict(net.state_dict())
    k1 = _param_key(net)
    assert k1 != k0
    m.focuser.load_state_dict(m.focuser.state_dict(), strict=False)        # ancestor
    k2 = _param_key(net)
    assert k2 != k1
    net.layer3[2].conv2.load_state_dict(net.layer3[2].conv2.state_dict())  # descendant
    k3 = _param_key(net)
    assert k3 != k2
    invalidate_packed(net)
    assert _param_key(net) != k3
    assert _param_key(m.glancer.net) == _param_key(m.glancer.net)


def test_pack_cache_content_hash(tmp_path):
    """packcache keys on the CONTENT of the state_dict: equal weights in another module instance hit, changed weights
    miss; the stored object is the runner itself."""
    from adafocus_b200 import packcache
    lin = torch.nn.Linear(64, 8)
    lin2 = torch.nn.Linear(64, 8)
    lin2.load_state_dict(lin.state_dict())
    assert packcache.content_hash(lin) == packcache.content_hash(lin2)
    with torch.no_grad():
        lin2.weight[0, 0] += 1
    assert packcache.content_hash(lin) != packcache.content_hash(lin2)

    packcache.set_cache_dir(str(tmp_path))
    R = lambda: _PickleRunner(lin.weight.detach().clone())      # noqa: E731
    try:
        before = dict(packcache.stats)
        r1 = packcache.cached_runner(lin, "R", R, "k1")
        r2 = packcache.cached_runner(lin, "R", lambda: (_ for _ in ()).throw(AssertionError("must hit")), "k2")
        assert r2.key == "k2" and torch.equal(r1.t, r2.t)
        assert packcache.stats["hits"] == before["hits"] + 1 and packcache.stats["stores"] == before["stores"] + 1
    finally:
        packcache.set_cache_dir(None)


class _PickleRunner:
    def __init__(self, t):
        self.key, self.t = None, t
