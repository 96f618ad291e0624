"""GPU parity of the bf16 NHWC trunk kernels (fused BatchNorm(+residual)(+ReLU), max-pool) against torch's own ops
evaluated in fp32 on the same bf16 inputs, and of the whole fused trunk against the torchvision/cuDNN path."""
import pytest
import torch
import torch.nn.functional as F

from util_gpu import rel

pytestmark = pytest.mark.gpu


def _cl_bf16(*shape):
    return torch.randn(*shape, device="cuda").to(dtype=torch.bfloat16, memory_format=torch.channels_last)


@pytest.mark.parametrize("N,C,H,W,res,relu", [(4, 64, 56, 56, False, True), (3, 128, 28, 28, True, True),
                                               (2, 512, 7, 7, False, False), (5, 256, 14, 14, True, True),
                                               (1, 64, 112, 112, False, True)])
def test_fused_bn_act_forward_backward(N, C, H, W, res, relu):
    from soccerdiffusion_b200.ml.model.encoder.trunk import FusedBNAct

    torch.manual_seed(C + H)
    x = (_cl_bf16(N, C, H, W) * 1.7 + 0.3).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    r = _cl_bf16(N, C, H, W) if res else None
    gamma = (torch.rand(C, device="cuda") + 0.5).requires_grad_(True)
    beta = (torch.randn(C, device="cuda") * 0.1).requires_grad_(True)
    rm, rv = torch.zeros(C, device="cuda"), torch.ones(C, device="cuda")
    rm2, rv2 = rm.clone(), rv.clone()
    go = _cl_bf16(N, C, H, W)
    xg = x.clone().requires_grad_(True)
    rg = r.clone().requires_grad_(True) if res else None
    y = FusedBNAct.apply(xg, gamma, beta, rm, rv, rg, relu, True, 0.1, 1e-5)
    y.backward(go)
    # reference: torch ops in fp32 on the same (bf16-valued) inputs
    xf = x.float().requires_grad_(True)
    rf = r.float().requires_grad_(True) if res else None
    g2, b2 = gamma.detach().clone().requires_grad_(True), beta.detach().clone().requires_grad_(True)
    yf = F.batch_norm(xf, rm2, rv2, g2, b2, True, 0.1, 1e-5)
    if res:
        yf = yf + rf
    if relu:
        yf = F.relu(yf)
    yf.backward(go.float())
    assert rel(y.float(), yf) < 6e-3                       # bf16 output rounding
    assert rel(rm, rm2) < 1e-4 and rel(rv, rv2) < 1e-4      # running statistics (fp32)
    assert rel(xg.grad.float(), xf.grad) < 1.5e-2
    assert rel(gamma.grad, g2.grad) < 1.5e-2 and rel(beta.grad, b2.grad) < 1.5e-2
    if res:
        assert rel(rg.grad.float(), rf.grad) < 1e-2
    # eval mode uses the running statistics
    ye = FusedBNAct.apply(x, gamma.detach(), beta.detach(), rm, rv, r, relu, False, 0.1, 1e-5)
    yef = F.batch_norm(x.float(), rm2, rv2, g2.detach(), b2.detach(), False, 0.1, 1e-5)
    if res:
        yef = yef + r.float()
    if relu:
        yef = F.relu(yef)
    assert rel(ye.float(), yef) < 6e-3


@pytest.mark.parametrize("N,C,H,W", [(3, 64, 112, 112), (2, 64, 32, 32), (1, 16, 7, 9)])
def test_maxpool_forward_backward(N, C, H, W):
    from soccerdiffusion_b200.ml.model.encoder.trunk import MaxPool3x3s2

    x = _cl_bf16(N, C, H, W)
    xg = x.clone().requires_grad_(True)
    y = MaxPool3x3s2.apply(xg)
    xf = x.float().requires_grad_(True)
    yf = F.max_pool2d(xf, 3, 2, 1)
    assert torch.equal(y.float(), yf)
    go = _cl_bf16(*y.shape)
    y.backward(go)
    yf.backward(go.float())
    assert rel(xg.grad.float(), xf.grad) < 5e-3   # sums of <= 4 bf16 gradients, rounded once to bf16


@pytest.mark.parametrize("N,C,H,W", [(3, 64, 112, 112), (2, 64, 30, 34), (1, 16, 7, 9), (2, 64, 33, 47), (5, 64, 2, 2),
                                     (1, 64, 9, 112), (2, 64, 20, 120)])
def test_fused_stem_bn_relu_pool(N, C, H, W):
    from soccerdiffusion_b200.ml.model.encoder.trunk import StemBNReLUPool

    torch.manual_seed(H)
    x = (_cl_bf16(N, C, H, W) * 1.3 - 0.2).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    gamma = (torch.rand(C, device="cuda") + 0.5).requires_grad_(True)
    beta = (torch.randn(C, device="cuda") * 0.2).requires_grad_(True)
    rm, rv = torch.zeros(C, device="cuda"), torch.ones(C, device="cuda")
    rm2, rv2 = rm.clone(), rv.clone()
    xg = x.clone().requires_grad_(True)
    y = StemBNReLUPool.apply(xg, gamma, beta, rm, rv, True, 0.1, 1e-5)
    go = _cl_bf16(*y.shape)
    y.backward(go)
    xf = x.float().requires_grad_(True)
    g2, b2 = gamma.detach().clone().requires_grad_(True), beta.detach().clone().requires_grad_(True)
    yf = F.max_pool2d(F.relu(F.batch_norm(xf, rm2, rv2, g2, b2, True, 0.1, 1e-5)), 3, 2, 1)
    yf.backward(go.float())
    assert rel(y.float(), yf) < 6e-3
    assert rel(rm, rm2) < 1e-4 and rel(rv, rv2) < 1e-4
    assert rel(xg.grad.float(), xf.grad) < 1.5e-2
    assert rel(gamma.grad, g2.grad) < 1.5e-2 and rel(beta.grad, b2.grad) < 1.5e-2


def test_stem_pack_and_space_to_depth_conv_equal_conv1():
    from soccerdiffusion_b200.ml.model.encoder.trunk import _stem_conv_s2d

    torch.manual_seed(0)
    conv = torch.nn.Conv2d(3, 64, 7, 2, 3, bias=False).cuda()
    img = torch.randn(3, 3, 64, 96, device="cuda")
    with torch.autocast("cuda", dtype=torch.bfloat16):
        got = _stem_conv_s2d(conv, img)                        # packed by sd_stem_pack_s2d_bf16
        got2 = _stem_conv_s2d(conv, img.double().float().requires_grad_(True))   # torch-op packing path
        want = conv(img.to(memory_format=torch.channels_last))
    assert got.shape == want.shape
    assert rel(got.float(), want.float()) < 4e-3 and rel(got2.float(), want.float()) < 4e-3
    assert torch.equal(got, got2)                              # the two packings feed cuDNN identical operands
    got.float().square().mean().backward()
    g1 = conv.weight.grad.clone()
    conv.weight.grad = None
    with torch.autocast("cuda", dtype=torch.bfloat16):
        conv(img.to(memory_format=torch.channels_last)).float().square().mean().backward()
    assert rel(g1, conv.weight.grad) < 1e-2


@pytest.mark.parametrize("N,H,W", [(2, 40, 48), (3, 112, 112), (2, 31, 45)])
def test_stem_band_backward_matches_two_kernel_path(N, H, W):
    """sd_stem_bn_relu_pool_nhwc_bf16_bwd (row-band kernels: shared-memory scatter of the pooled gradient, ReLU mask
    recomputed) against the generic path (max-pool backward + BatchNorm backward with the recomputed ReLU mask), both
    fed with the tensors the band forward saved (same arg-max tap encoding)."""
    from soccerdiffusion_b200 import ops
    from soccerdiffusion_b200.ml.model.encoder.trunk import StemBNReLUPool

    torch.manual_seed(3)
    C = 64
    assert ops.stem_band_supported(H, W, C)
    x = (_cl_bf16(N, C, H, W) * 1.2 + 0.1).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    gamma = (torch.rand(C, device="cuda") + 0.5).requires_grad_(True)
    beta = (torch.randn(C, device="cuda") * 0.2).requires_grad_(True)
    rm, rv = torch.zeros(C, device="cuda"), torch.ones(C, device="cuda")
    xg = x.clone().requires_grad_(True)
    y = StemBNReLUPool.apply(xg, gamma, beta, rm, rv, True, 0.1, 1e-5)
    go = _cl_bf16(*y.shape)
    xs, idx, mean, invstd, g_, b_, _y = y.grad_fn.saved_tensors   # read before backward() frees them
    y.backward(go)                                            # band kernels
    # generic two-kernel path from the same saved tensors
    sums = torch.empty(2 * C, device="cuda", dtype=torch.float64)
    dact, dx = torch.empty_like(xs), torch.empty_like(xs)
    dg, db = torch.empty(C, device="cuda"), torch.empty(C, device="cuda")
    ops.maxpool_bwd(go, idx, dact, N, H, W, C)
    ops.bn_bwd(dact, None, xs, mean, invstd, g_, sums, dx, None, dg, db, N * H * W, C, beta_recompute=b_)
    assert rel(xg.grad.float(), dx.float()) < 1e-2
    assert rel(gamma.grad, dg) < 1e-2 and rel(beta.grad, db) < 1e-2   # the two-kernel path rounds the pooled gradient to bf16


@pytest.mark.parametrize("N,H,W", [(4, 224, 224), (3, 20, 36), (2, 64, 96), (1, 6, 6)])
def test_stem_wgrad_tcgen05_matches_cudnn(N, H, W):
    """conv1 weight gradient: tcgen05 kernel on the space-to-depth image vs aten.convolution_backward (7x7 form)."""
    from soccerdiffusion_b200.ml.model.encoder import trunk as T

    torch.manual_seed(N * H)
    w = torch.randn(64, 3, 7, 7, device="cuda") * 0.05
    img = torch.randn(N, 3, H, W, device="cuda")
    dy = _cl_bf16(N, 64, H // 2, W // 2)
    got = {}
    for mode in (True, False):
        T._USE_TC_STEM_WGRAD = mode
        try:
            wg = w.clone().requires_grad_(True)
            y, _ = T.StemConvS2D.apply(img, wg)
            y.backward(dy)
            got[mode] = wg.grad.clone()
        finally:
            T._USE_TC_STEM_WGRAD = True
    # exact reference in float64 on the bf16-rounded operands
    ref = torch.nn.grad.conv2d_weight(img.to(torch.bfloat16).double(), (64, 3, 7, 7), dy.double(), stride=2, padding=3)
    assert rel(got[True], ref) < 2e-5, rel(got[True], ref)        # fp32 accumulation of exact bf16 products
    assert rel(got[False], ref) < 1e-2                            # cuDNN result is rounded to bf16


@pytest.mark.parametrize("N,H,W", [(4, 224, 224), (2, 64, 96), (3, 16, 16)])
def test_stem_fprop_tcgen05_matches_cudnn(N, H, W):
    """conv1 forward: TMA + tcgen05 kernel on the space-to-depth image vs the library convolution, and vs float64."""
    from soccerdiffusion_b200.ml.model.encoder import trunk as T

    torch.manual_seed(N + H)
    w = torch.randn(64, 3, 7, 7, device="cuda") * 0.05
    img = torch.randn(N, 3, H, W, device="cuda")
    outs = {}
    for mode in (True, False):
        T._USE_TC_STEM_FPROP = mode
        try:
            outs[mode] = T._stem_conv_s2d_raw(img, w).float()
        finally:
            T._USE_TC_STEM_FPROP = True
    ref = torch.nn.functional.conv2d(img.to(torch.bfloat16).double(), w.to(torch.bfloat16).double(), None, 2, 3)
    assert outs[True].shape == ref.shape
    assert rel(outs[True], ref) < 4e-3 and rel(outs[False], ref) < 4e-3     # bf16 output rounding
    assert rel(outs[True], outs[False]) < 4e-3


def test_fused_trunk_matches_library_trunk_bf16():
    """Whole trunk, train mode: libsd_b200 path (space-to-depth stem, fused BN/ReLU/pool kernels) and torch's bf16
    autocast path are both compared with the fp32 trunk; the fused path must be as close to fp32 as the library
    bf16 path is (a layout/indexing bug would show as a much larger error, bf16 noise does not)."""
    import soccerdiffusion_b200 as sdb
    from oracle import synth
    from soccerdiffusion_b200 import runtime
    from util_gpu import synth_model

    hp = dict(synth.TINY_HP, image_resolution=96, image_context_length=3)
    outs = {}
    try:
        for name, prec, fused in (("fp32", "fp32", False), ("fused", "bf16", True), ("lib", "bf16", False)):
            sdb.set_precision(prec)
            runtime.set_fused_trunk(fused)
            model, _ = synth_model(hp, 3)
            model.train()
            enc = model.image_sequence_encoder.image_encoder
            imgs = torch.from_numpy(synth.normal("img", (2 * 3, 3, 96, 96), 3)).cuda()
            feat = enc.trunk(imgs)
            tgt = torch.from_numpy(synth.normal("tgt", tuple(feat.shape), 3)).cuda()
            (feat.float() * tgt).mean().backward()   # linear functional of the features: no amplification by the loss
            e = enc.encoder
            outs[name] = dict(feat=feat.float().detach(), rm=e.bn1.running_mean.clone(), rv=e.layer2[0].bn2.running_var.clone(),
                              nbt=int(e.bn1.num_batches_tracked), g_conv1=e.conv1.weight.grad.clone(),
                              g_bn1=e.bn1.weight.grad.clone(), g_l1=e.layer1[0].conv1.weight.grad.clone(),
                              g_bn=e.layer3[1].bn2.weight.grad.clone(), g_ds=e.layer4[0].downsample[0].weight.grad.clone())
    finally:
        runtime.set_fused_trunk(True)
        sdb.set_precision("fp32")
    ref, a, b = outs["fp32"], outs["fused"], outs["lib"]
    assert a["nbt"] == b["nbt"] == ref["nbt"] == 1
    for k in ("feat", "rm", "rv", "g_ds", "g_bn", "g_l1", "g_bn1", "g_conv1"):
        ea, eb = rel(a[k], ref[k]), rel(b[k], ref[k])
        print(f"{k}: fused-vs-fp32 {ea:.3e}  torch-bf16-vs-fp32 {eb:.3e}")
        assert ea < max(2.0 * eb, 2e-2), (k, ea, eb)


def test_fused_trunk_resnet50_bottleneck_blocks():
    """ImageEncoderType.RESNET50 (Bottleneck blocks: three BatchNorms per block, forked block outputs) through the fused
    trunk, train mode, against the fp32 trunk — judged like the ResNet18 test above."""
    import soccerdiffusion_b200 as sdb
    from soccerdiffusion_b200 import runtime
    from soccerdiffusion_b200.ml.model.encoder.image import ImageEncoderType, ResNetImageEncoder

    g = torch.Generator().manual_seed(17)
    imgs = torch.randn(8, 3, 128, 128, generator=g).cuda()   # 128 samples per channel in layer4: train-mode BN stays well conditioned
    outs = {}
    try:
        for name, prec, fused in (("fp32", "fp32", False), ("fused", "bf16", True), ("lib", "bf16", False)):
            sdb.set_precision(prec)
            runtime.set_fused_trunk(fused)
            torch.manual_seed(5)
            enc = ResNetImageEncoder(ImageEncoderType.RESNET50, 32, True, 128)
            # a random-initialised 50-layer net in train mode amplifies bf16 rounding to O(1) errors (both bf16 paths
            # alike); five Bottleneck blocks keep the comparison discriminating
            r = enc.encoder
            r.layer1, r.layer2, r.layer3, r.layer4 = r.layer1[:2], r.layer2[:1], r.layer3[:1], r.layer4[:1]
            enc = enc.cuda().train()
            feat = enc.trunk(imgs)
            if name == "fp32":
                tgt = torch.randn(feat.shape, generator=g).cuda()
            (feat.float() * tgt).mean().backward()
            e = enc.encoder
            outs[name] = dict(feat=feat.float().detach(), g_c3=e.layer1[0].conv3.weight.grad.clone(),
                              g_bn3=e.layer2[0].bn3.weight.grad.clone(), g_ds=e.layer3[0].downsample[0].weight.grad.clone(),
                              g_conv1=e.conv1.weight.grad.clone(), rv=e.layer1[1].bn2.running_var.clone())
    finally:
        runtime.set_fused_trunk(True)
        sdb.set_precision("fp32")
    ref, a, b = outs["fp32"], outs["fused"], outs["lib"]
    for k in ("feat", "rv", "g_c3", "g_bn3", "g_ds", "g_conv1"):
        ea, eb = rel(a[k], ref[k]), rel(b[k], ref[k])
        print(f"{k}: fused-vs-fp32 {ea:.3e}  torch-bf16-vs-fp32 {eb:.3e}")
        assert ea < max(2.0 * eb, 3e-2), (k, ea, eb)


@pytest.mark.parametrize("N,H,W", [(3, 224, 224), (2, 20, 36)])
def test_uint8_stem_pack_equals_reference_preprocessing(N, H, W):
    """Raw uint8 frames packed by sd_stem_pack_s2d_u8 == the reference's host preprocessing (torchvision v2.ToDtype(float32,
    scale=True) -> v2.Normalize(ImageNet), dataset/pytorch.py:198-204) followed by sd_stem_pack_s2d_bf16: bit for bit."""
    from torchvision.transforms import v2

    from soccerdiffusion_b200 import ops
    from soccerdiffusion_b200.ml.model.encoder.trunk import normalize_u8

    g = torch.Generator().manual_seed(5)
    u8 = torch.randint(0, 256, (N, 3, H, W), generator=g, dtype=torch.uint8)
    u8[0, :, 0, :4] = torch.tensor([0, 1, 254, 255], dtype=torch.uint8)
    pre = v2.Compose([v2.ToDtype(torch.float32, scale=True), v2.Normalize((0.485, 0.456, 0.406), (0.229, 0.224, 0.225))])
    ref = torch.stack([pre(img) for img in u8])                      # host, as the reference does per frame
    Hp, Wp = (H + 6) // 2, (W + 6) // 2
    a = torch.empty((N, Hp, Wp, 16), device="cuda", dtype=torch.bfloat16)
    b = torch.empty_like(a)
    ops.stem_pack_u8(u8.cuda(), a, N, H, W)
    ops.stem_pack(ref.cuda(), b, N, H, W)
    assert torch.equal(a.view(torch.int16), b.view(torch.int16))
    assert torch.equal(normalize_u8(u8.cuda()).cpu(), ref)            # generic (non-fused) device path: same fp32 values


@pytest.mark.parametrize("precision", ["bf16", "fp32"])
def test_model_accepts_raw_uint8_frames(precision):
    """End2EndDiffusionTransformer.forward with uint8 image_data == the same call with host-preprocessed float frames."""
    from torchvision.transforms import v2

    import soccerdiffusion_b200 as sd
    from soccerdiffusion_b200 import config, runtime

    hp = dict(config.DEFAULT)
    hp.update(image_context_length=2, image_resolution=64, image_use_final_avgpool=False, action_context_length=10,
              imu_context_length=10, joint_state_context_length=10)
    prev = sd.precision_name()
    sd.set_precision(precision)
    try:
        torch.manual_seed(0)
        model = config.build_model(hp).cuda().eval()
        batch = config.synthetic_batch(hp, 3, "cuda", seed=1, uint8_images=True)
        u8 = batch["image_data"]
        pre = v2.Compose([v2.ToDtype(torch.float32, scale=True), v2.Normalize((0.485, 0.456, 0.406), (0.229, 0.224, 0.225))])
        flt = torch.stack([pre(img) for img in u8.cpu().flatten(0, 1)]).view(*u8.shape).cuda()
        x = torch.randn(3, hp["trajectory_prediction_length"], hp["num_joints"], device="cuda")
        t = torch.tensor([5, 500, 900], device="cuda")
        with torch.no_grad():
            a = model(batch, x, t)
            b = model({**batch, "image_data": flt}, x, t)
        assert torch.equal(a, b)
        # training mode (train-mode BatchNorm, fused stem with the uint8 packing): gradients identical too
        model.train()
        runtime.set_dropout(0.0)
        grads = []
        for imgs in (u8, flt):
            model.zero_grad()
            model({**batch, "image_data": imgs}, x, t).square().mean().backward()
            grads.append(model.image_sequence_encoder.image_encoder.encoder.conv1.weight.grad.clone())
        assert rel(grads[0], grads[1]) < 1e-5   # split-K atomics: summation order differs run to run
    finally:
        runtime.set_dropout(0.1)
        sd.set_precision(prev)


@pytest.mark.parametrize("N,H,W", [(3, 224, 224), (2, 64, 96)])
def test_stem_fprop_epilogue_statistics(N, H, W):
    """sd_stem_fprop_s2d_bf16_stats: the per-channel sums accumulated in the convolution's epilogue give the same
    BatchNorm batch statistics as a pass over the stored bf16 map (they are taken before the bf16 rounding)."""
    from soccerdiffusion_b200.ml.model.encoder import trunk as T

    torch.manual_seed(11)
    img = torch.randn(N, 3, H, W, device="cuda")
    w = torch.randn(64, 3, 7, 7, device="cuda") * 0.05
    sums = torch.full((128,), 7.0, device="cuda", dtype=torch.float64)      # the call zeroes it
    y, _, have = T._stem_conv_s2d_raw(img, w, return_packed=True, sums=sums)
    assert have
    yf = y.float()
    n = yf.numel() // 64
    mean_ref, var_ref = yf.mean((0, 2, 3)).double(), yf.var((0, 2, 3), unbiased=False).double()
    mean = sums[:64] / n
    var = sums[64:] / n - mean * mean
    assert ((mean - mean_ref).abs() < 1e-3 * var_ref.sqrt()).all()
    assert ((var / var_ref - 1).abs() < 2e-3).all()


@pytest.mark.parametrize("res", [False, True])
def test_fused_bn_fork_sums_two_gradients(res):
    """FusedBNAct(fork=True) returns twin outputs; their two gradients are summed inside the backward kernels
    (sd_bn_bwd2_nhwc_bf16) — same result as BatchNorm backward on the fp32 sum; an unused twin costs nothing."""
    from soccerdiffusion_b200.ml.model.encoder.trunk import FusedBNAct

    torch.manual_seed(21)
    N, C, H, W = 3, 64, 20, 24
    x = (_cl_bf16(N, C, H, W) * 1.7 + 0.3).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    r = _cl_bf16(N, C, H, W) if res else None
    gamma = (torch.rand(C, device="cuda") + 0.5).requires_grad_(True)
    beta = (torch.randn(C, device="cuda") * 0.1).requires_grad_(True)
    ga, gb = _cl_bf16(N, C, H, W), _cl_bf16(N, C, H, W)
    for use_b in (True, False):
        gamma.grad = beta.grad = None
        xg = x.clone().requires_grad_(True)
        rg = r.clone().requires_grad_(True) if res else None
        rm, rv = torch.zeros(C, device="cuda"), torch.ones(C, device="cuda")
        y1, y2 = FusedBNAct.apply(xg, gamma, beta, rm, rv, rg, True, True, 0.1, 1e-5, True)
        assert y1.data_ptr() == y2.data_ptr()
        if use_b:
            torch.autograd.backward([y1, y2], [ga, gb])
        else:
            y2.backward(ga)                                   # only the second twin is consumed
        xf = x.float().requires_grad_(True)
        rf = r.float().requires_grad_(True) if res else None
        g2, b2 = gamma.detach().clone().requires_grad_(True), beta.detach().clone().requires_grad_(True)
        yf = F.batch_norm(xf, torch.zeros(C, device="cuda"), torch.ones(C, device="cuda"), g2, b2, True, 0.1, 1e-5)
        yf = F.relu(yf + rf if res else yf)
        yf.backward(ga.float() + gb.float() if use_b else ga.float())
        assert rel(xg.grad.float(), xf.grad) < 1.5e-2
        assert rel(gamma.grad, g2.grad) < 1.5e-2 and rel(beta.grad, b2.grad) < 1.5e-2
        if res:
            assert rel(rg.grad.float(), rf.grad) < 1e-2


@pytest.mark.parametrize("case", ["typical", "fallback_large_beta_over_gamma", "fallback_zero_gamma"])
def test_stem_backward_pooled_domain_reductions(case):
    """sd_stem_bn_relu_pool_nhwc_bf16_bwd2 with the pooled output: the per-channel reductions taken over pooling windows
    (sum dp*(y>0), sum dp*(y>0)*(y-beta)/gamma) equal the pixel-domain ones; parameters outside the safe range (|beta/gamma|
    > 16, gamma ~ 0) take the exact kernel on the device."""
    from soccerdiffusion_b200 import ops

    torch.manual_seed(9)
    N, C, H, W = 4, 64, 64, 96
    x = (_cl_bf16(N, C, H, W) * 1.2 + 0.1).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    gamma = torch.rand(C, device="cuda") + 0.5
    beta = torch.randn(C, device="cuda") * 0.3
    if case == "fallback_large_beta_over_gamma":
        gamma[5], beta[5] = 0.01, 0.9
    elif case == "fallback_zero_gamma":
        gamma[7] = 0.0
    mean, invstd = torch.empty(C, device="cuda"), torch.empty(C, device="cuda")
    sums = torch.empty(2 * C, device="cuda", dtype=torch.float64)
    rm, rv = torch.zeros(C, device="cuda"), torch.ones(C, device="cuda")
    ops.bn_stats(x, N * H * W, C, sums, 1e-5, 0.1, mean, invstd, rm, rv)
    HO, WO = H // 2, W // 2
    y = torch.empty((N, C, HO, WO), device="cuda", dtype=torch.bfloat16, memory_format=torch.channels_last)
    idx = torch.empty((N, HO, WO, C), device="cuda", dtype=torch.uint8)
    ops.stem_fwd(x, mean, invstd, gamma, beta, y, idx, N, H, W, C)
    dp = _cl_bf16(N, C, HO, WO)
    out = {}
    for use_y in (False, True):
        dx = torch.empty_like(x)
        dg, db = torch.empty(C, device="cuda"), torch.empty(C, device="cuda")
        ops.stem_bwd(dp, idx, x, mean, invstd, gamma, beta, sums, dx, dg, db, N, H, W, C, y_pooled=y if use_y else None)
        out[use_y] = (dx.float(), dg.clone(), db.clone())
    if case == "typical":
        assert rel(out[True][2], out[False][2]) < 1e-5                      # sum g: same terms, other order
        # sum g*xhat: xhat is recovered from the bf16-rounded y (relative error <= 2^-9 per term, like dp's own rounding);
        # with random dp the sum is itself only sqrt(n)-sized, so the relative difference sits at the bf16 level
        assert rel(out[True][1], out[False][1]) < 5e-3
        assert rel(out[True][0], out[False][0]) < 5e-3
    else:                                                                   # exact kernel ran in both calls
        assert rel(out[True][1], out[False][1]) < 1e-5 and rel(out[True][0], out[False][0]) < 1e-5


@pytest.mark.parametrize("n,H,cin,cout", [(6, 56, 64, 128), (5, 28, 128, 256), (7, 14, 256, 512), (3, 8, 64, 64)])
def test_downsample_conv1x1s2_dgrad_matches_library(n, H, cin, cout):
    """sd_conv1x1s2_dgrad_bf16 (TMA + tcgen05 GEMM, strided scatter) against torch's convolution_backward on the same bf16
    operands, and the autograd node of the trunk's downsample path against F.conv2d's autograd."""
    from soccerdiffusion_b200 import ops
    from soccerdiffusion_b200.ml.model.encoder.trunk import DownsampleConv1x1S2

    gen = torch.Generator().manual_seed(n + H)
    cl = torch.channels_last
    x = torch.randn(n, cin, H, H, generator=gen).cuda().to(torch.bfloat16).contiguous(memory_format=cl)
    w = (torch.randn(cout, cin, 1, 1, generator=gen) / cin ** 0.5).cuda()
    assert ops.conv1x1s2_dgrad_supported(H, H, cin, cout)
    wb4 = w.to(torch.bfloat16).contiguous(memory_format=cl)
    y = torch.ops.aten.convolution(x, wb4, None, [2, 2], [0, 0], [1, 1], False, [0, 0], 1)
    dy = torch.randn(y.shape, generator=gen).cuda().to(torch.bfloat16).contiguous(memory_format=cl)
    want_dx, want_dw, _ = torch.ops.aten.convolution_backward(dy, x, wb4, None, [2, 2], [0, 0], [1, 1], False, [0, 0], 1,
                                                              [True, True, False])
    dx = torch.full_like(x, float("nan"))
    ops.conv1x1s2_dgrad(dy, w.to(torch.bfloat16).view(cout, cin).contiguous(), dx, n, H, H, cin, cout)
    assert rel(dx.float(), want_dx.float()) < 5e-3, rel(dx.float(), want_dx.float())
    assert not dx[:, :, 1::2, :].any() and not dx[:, :, :, 1::2].any()   # the pixels the stride skips
    xr = x.clone().requires_grad_(True)
    wr = w.clone().requires_grad_(True)
    out = DownsampleConv1x1S2.apply(xr, wr)
    assert rel(out.float(), y.float()) < 1e-6
    out.backward(dy)
    assert rel(xr.grad.float(), want_dx.float()) < 5e-3
    assert wr.grad.dtype == torch.float32 and rel(wr.grad, want_dw.float()) < 5e-3


@pytest.mark.parametrize("n,H,W", [(3, 56, 56), (2, 8, 12), (5, 7, 20), (1, 4, 4), (148, 12, 12)])
def test_conv3x3_wgrad_c64_matches_library(n, H, W):
    """sd_conv3x3_wgrad_c64_bf16 (implicit GEMM over the pixels, every tap a shifted view of two TMA tiles) against torch's
    convolution_backward on the same bf16 operands; deterministic; accumulate mode."""
    from soccerdiffusion_b200 import ops

    assert ops.conv3x3_wgrad_c64_supported(H, W)
    gen = torch.Generator().manual_seed(n * 100 + H + W)
    cl = torch.channels_last
    x = torch.randn(n, 64, H, W, generator=gen).cuda().to(torch.bfloat16).contiguous(memory_format=cl)
    dy = torch.randn(n, 64, H, W, generator=gen).cuda().to(torch.bfloat16).contiguous(memory_format=cl)
    w = torch.zeros(64, 64, 3, 3, device="cuda", dtype=torch.bfloat16).contiguous(memory_format=cl)
    want = torch.ops.aten.convolution_backward(dy.float(), x.float(), w.float(), None, [1, 1], [1, 1], [1, 1], False, [0, 0], 1,
                                               [False, True, False])[1]
    dW = torch.full((64, 64, 3, 3), float("nan"), device="cuda")
    ops.conv3x3_wgrad_c64(x, dy, dW, n, H, W)
    assert rel(dW, want) < 2e-3, rel(dW, want)
    dW2 = torch.full((64, 64, 3, 3), float("nan"), device="cuda")
    ops.conv3x3_wgrad_c64(x, dy, dW2, n, H, W)
    assert torch.equal(dW, dW2)
    base = torch.randn(64, 64, 3, 3, generator=gen).cuda()
    acc = base.clone()
    ops.conv3x3_wgrad_c64(x, dy, acc, n, H, W, accumulate=True)
    assert rel(acc, base + want) < 2e-3


def test_tensor_map_kernels_as_first_call_in_a_fresh_process():
    """cuTensorMapEncodeTiled needs a context bound to the thread: each TMA-fed entry point must work when it is the FIRST call into
    the library after another library (cuDNN) has run — run in a fresh interpreter so no earlier test has initialised anything."""
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = r'''
import sys, torch
sys.path.insert(0, %r)
from soccerdiffusion_b200 import ops
which = sys.argv[1]
cl = torch.channels_last
x = torch.randn(2, 64, 16, 16, device="cuda", dtype=torch.bfloat16).contiguous(memory_format=cl)
w = torch.randn(64, 64, 3, 3, device="cuda", dtype=torch.bfloat16).contiguous(memory_format=cl)
y = torch.ops.aten.convolution(x, w, None, [1, 1], [1, 1], [1, 1], False, [0, 0], 1)   # cuDNN first
if which == "wgrad":
    dW = torch.empty(64, 64, 3, 3, device="cuda")
    ops.conv3x3_wgrad_c64(x, y, dW, 2, 16, 16)
elif which == "ds":
    dy = torch.randn(2, 128, 8, 8, device="cuda", dtype=torch.bfloat16).contiguous(memory_format=cl)
    wb = torch.randn(128, 64, device="cuda", dtype=torch.bfloat16)
    ops.conv1x1s2_dgrad(dy, wb, torch.empty_like(x), 2, 16, 16, 64, 128)
elif which == "stem":
    from soccerdiffusion_b200.ml.model.encoder import trunk as T
    out, _, have = T._stem_conv_s2d_raw(torch.randn(2, 3, 64, 64, device="cuda"), torch.randn(64, 3, 7, 7, device="cuda") * 0.05,
                                        return_packed=True, sums=torch.zeros(128, device="cuda", dtype=torch.float64))
    assert have, "the TMA stem kernel was not used"
torch.cuda.synchronize()
print("ok")
''' % root
    for which in ("wgrad", "ds", "stem"):
        r = subprocess.run([sys.executable, "-c", code, which], capture_output=True, text=True, timeout=300)
        assert r.returncode == 0 and "ok" in r.stdout, (which, r.stdout[-500:], r.stderr[-1500:])


def test_tensor_map_kernels_inside_cuda_graph_capture():
    """Every TMA-fed trunk entry point must be capturable (the training step is one CUDA graph): no call that is illegal during
    stream capture on their host paths (a cudaFree(0) there once invalidated the capture of the whole step)."""
    from soccerdiffusion_b200 import ops
    from soccerdiffusion_b200.ml.model.encoder import trunk as T

    cl = torch.channels_last
    torch.manual_seed(0)
    x = torch.randn(2, 64, 16, 16, device="cuda", dtype=torch.bfloat16).contiguous(memory_format=cl)
    dy = torch.randn(2, 64, 16, 16, device="cuda", dtype=torch.bfloat16).contiguous(memory_format=cl)
    dyd = torch.randn(2, 128, 8, 8, device="cuda", dtype=torch.bfloat16).contiguous(memory_format=cl)
    wb = torch.randn(128, 64, device="cuda", dtype=torch.bfloat16)
    img = torch.randn(2, 3, 64, 64, device="cuda")
    w7 = torch.randn(64, 3, 7, 7, device="cuda") * 0.05
    dW, dx = torch.zeros(64, 64, 3, 3, device="cuda"), torch.zeros_like(x)

    def run():
        ops.conv3x3_wgrad_c64(x, dy, dW, 2, 16, 16)
        ops.conv1x1s2_dgrad(dyd, wb, dx, 2, 16, 16, 64, 128)
        return T._stem_conv_s2d_raw(img, w7)

    want = run().clone()
    want_dW, want_dx = dW.clone(), dx.clone()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        out = run()
    dW.zero_(); dx.zero_()
    g.replay()
    torch.cuda.synchronize()
    assert torch.equal(out, want) and torch.equal(dW, want_dW) and torch.equal(dx, want_dx)
