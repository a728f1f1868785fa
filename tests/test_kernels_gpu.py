"""Per-kernel numerics on the B200: every C-ABI entry point against a plain PyTorch fp32 statement of the same
operation (the path-level parity tests against the oracle live in test_parity_gpu.py)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

DEV = "cuda"


@pytest.fixture(autouse=True)
def _strict_fp32():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield


def _ops():
    from rspnet_b200 import ops
    return ops


def rand(*shape, seed=0, scale=1.0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(DEV)


# ------------------------------------------------------------------------------------------------ MoCo kernels
@pytest.mark.parametrize("n", [0, 1, 7, 4096, 1_000_003])
def test_ema_bit_exact(n):
    ops = _ops()
    k, q = rand(n, seed=1), rand(n, seed=2)
    m = 0.999
    ref = k * m + q * (1.0 - m)  # the reference expression (builder:343)
    ops.ema_update_(k, q, m)
    assert torch.equal(k, ref)


def test_sgd_matches_torch_optim():
    ops = _ops()
    n = 100_003
    p0, g1, g2 = rand(n, seed=1), rand(n, seed=2), rand(n, seed=3)
    pr = torch.nn.Parameter(p0.clone())
    opt = torch.optim.SGD([pr], lr=0.1, momentum=0.9, weight_decay=1e-4)
    p = p0.clone()
    mom = torch.zeros_like(p)
    for i, g in enumerate((g1, g2)):
        pr.grad = g.clone()
        opt.step()
        ops.sgd_step_(p, g, mom, 0.1, 0.9, 1e-4, 1.0, first_step=(i == 0))
    torch.testing.assert_close(p, pr.data, rtol=1e-5, atol=1e-6)


def _diff_speed_ref(im_q, im_k, perm, n_s1, d):
    b, c, t, h, w = im_q.shape
    tr = t // d
    s1, s2 = perm[:n_s1], perm[n_s1:]
    sp1 = torch.arange(0, t, 1, device=DEV)[:tr]
    sp2 = torch.arange(0, t, d, device=DEV)[:tr]
    q = torch.empty(b, c, tr, h, w, device=DEV)
    k = torch.empty_like(q)
    kn = torch.empty_like(q)
    q[s1] = im_q.index_select(0, s1).index_select(2, sp1)
    q[s2] = im_q.index_select(0, s2).index_select(2, sp2)
    k[s1] = im_k.index_select(0, s1).index_select(2, sp1)
    k[s2] = im_k.index_select(0, s2).index_select(2, sp2)
    kn[s1] = im_k.index_select(0, s1).index_select(2, sp2)
    kn[s2] = im_k.index_select(0, s2).index_select(2, sp1)
    return q, k, kn


@pytest.mark.parametrize("b,t,hw,d", [(4, 8, 6, 2), (5, 12, 7, 2), (1, 4, 5, 2), (6, 8, 8, 4)])
def test_speed_gather(b, t, hw, d):
    ops = _ops()
    im_q, im_k = rand(b, 3, t, hw, hw, seed=1), rand(b, 3, t, hw, hw, seed=2)
    perm = torch.randperm(b, generator=torch.Generator().manual_seed(3)).to(DEV)
    n_s1 = int(b * 0.5)
    ref = _diff_speed_ref(im_q, im_k, perm, n_s1, d)
    got = ops.speed_gather(im_q, im_k, perm, n_s1, d, 0)
    for g, r in zip(got, ref):
        assert torch.equal(g, r)
    got1 = ops.speed_gather(im_q, im_k, perm, n_s1, d, 1)
    for g, r in zip(got1, ref):
        exp = r.permute(0, 2, 3, 4, 1).bfloat16()
        assert torch.equal(g[..., :3], exp)
        assert torch.count_nonzero(g[..., 3]) == 0


def test_gather_rows_and_layout_roundtrip():
    ops = _ops()
    src = rand(10, 3, 4, 8, seed=1)
    idx = torch.tensor([9, 0, 3, 3, 7], device=DEV)
    assert torch.equal(ops.gather_rows(src, idx), src[idx])
    x = rand(2, 3, 4, 5, 6, seed=2)
    y = ops.to_ndhwc_bf16(x)
    assert y.shape == (2, 4, 5, 6, 4)
    assert torch.equal(ops.to_ncdhw_f32(y, 3), x.bfloat16().float())


def test_queue_enqueue_ring():
    ops = _ops()
    d, k, n = 128, 1024, 64
    queue = rand(d, k, seed=1)
    ref = queue.clone()
    ptr = torch.zeros(1, dtype=torch.long, device=DEV)
    p = 0
    for step in range(k // n + 2):  # wraps around
        keys = rand(n, d, seed=10 + step)
        ops.queue_enqueue_(queue, keys, ptr)
        ref[:, p:p + n] = keys.T
        p = (p + n) % k
        assert int(ptr) == p
    assert torch.equal(queue, ref)


def _logits_ref(q_a, q_m, k_a, k_m, kn_a, kn_m, queue, T):
    l_pos_a1 = torch.einsum("nc,nc->n", q_a, k_a).unsqueeze(-1) / T
    l_pos_a2 = torch.einsum("nc,nc->n", q_a, kn_a).unsqueeze(-1) / T
    l_pos_m = torch.einsum("nc,nc->n", q_m, k_m).unsqueeze(-1) / T
    l_neg_a = torch.einsum("nc,ck->nk", q_a, queue) / T
    l_neg_m = torch.einsum("nc,nc->n", q_m, kn_m).unsqueeze(-1) / T
    return torch.cat([l_pos_a1, l_neg_a], 1), torch.cat([l_pos_a2, l_neg_a], 1), l_pos_m, l_neg_m


@pytest.mark.parametrize("n,k", [(4, 512), (64, 16384), (5, 300)])
def test_moco_logits_loss_fwd_bwd(n, k):
    ops = _ops()
    d, T, margin, A, M = 128, 0.07, 2.0, 1.0, 0.5
    feats = [F.normalize(rand(n, d, seed=s), dim=1) for s in range(6)]
    queue = F.normalize(rand(d, k, seed=9), dim=0)
    q_a, q_m = feats[0].clone().requires_grad_(True), feats[1].clone().requires_grad_(True)
    l1, l2, lpm, lnm = _logits_ref(q_a, q_m, *feats[2:], queue, T)
    tgt = torch.zeros(n, dtype=torch.long, device=DEV)
    ce = F.cross_entropy(l1, tgt) + F.cross_entropy(l2, tgt)
    rank = torch.clamp(-(lpm - lnm) + margin, min=0).mean()
    loss = A * ce + M * rank
    loss.backward()

    logits, rows, ranks = ops.moco_logits_fwd(feats[0], feats[1], *feats[2:], queue, T, materialize=True)
    assert ranks.dtype == torch.int32 and ranks.shape == (2, n)
    for i in range(2):   # the counters against the logits the same call materialised: exact
        assert torch.equal(ranks[i].long(), (logits[i][:, 1:] > logits[i][:, :1]).sum(1))
    torch.testing.assert_close(logits[0], l1.detach(), rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(logits[1], l2.detach(), rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(rows[0], lpm.detach().squeeze(1), rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(rows[1], lnm.detach().squeeze(1), rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(rows[2], torch.logsumexp(l1.detach(), 1), rtol=1e-5, atol=1e-4)
    out3 = ops.moco_loss_fwd(rows, margin, A, M)
    torch.testing.assert_close(out3, torch.stack([loss, ce, rank]).detach(), rtol=1e-3, atol=1e-5)

    g3 = torch.tensor([1.0, 0.0, 0.0], device=DEV)
    g_rows = ops.moco_loss_bwd(rows, margin, A, M, g3)
    dq_a, dq_m = ops.moco_logits_bwd(feats[0], feats[1], *feats[2:], queue, T, rows, g_rows, None, None)
    torch.testing.assert_close(dq_a, q_a.grad, rtol=1e-3, atol=1e-5)
    torch.testing.assert_close(dq_m, q_m.grad, rtol=1e-3, atol=1e-5)

    # dense path: gradient arriving through the materialised logits (a caller doing its own CE)
    lse = ops.ce0_fwd(logits[0])
    torch.testing.assert_close(lse, torch.logsumexp(l1.detach(), 1), rtol=1e-5, atol=1e-4)
    one = torch.ones(1, device=DEV)
    dl1 = ops.ce0_bwd(logits[0], lse, one)
    dl2 = ops.ce0_bwd(logits[1], ops.ce0_fwd(logits[1]), one)
    zero_rows = torch.zeros_like(rows)
    dq_a2, _ = ops.moco_logits_bwd(feats[0], feats[1], *feats[2:], queue, T, rows, zero_rows, dl1, dl2)
    q_a2 = feats[0].clone().requires_grad_(True)
    r1, r2, _, _ = _logits_ref(q_a2, feats[1], *feats[2:], queue, T)
    (F.cross_entropy(r1, tgt) + F.cross_entropy(r2, tgt)).backward()
    torch.testing.assert_close(dq_a2, q_a2.grad, rtol=1e-3, atol=1e-5)


def test_contrastive_meters_match_reference_accuracy():
    """Rank counters + device meters against the reference's accuracy() (top-k over [N, 1+K]) and AverageMeter
    arithmetic (framework/metrics/classification.py:6-20, framework/meters/average.py:23-30, pretrain.py:169-196)."""
    from rspnet_b200 import meters as M
    ops = _ops()
    n, d, k, T = 64, 128, 4096, 0.07
    queue = F.normalize(rand(d, k, seed=9), dim=0)

    def ref_accuracy(output, target, topk):
        maxk = max(topk)
        _, pred = output.topk(maxk, 1, True, True)
        correct = pred.t().eq(target[None])
        return [correct[:kk].flatten().sum(dtype=torch.float) * (100.0 / target.size(0)) for kk in topk]

    meters = M.ContrastiveMeters(DEV)
    ref_sum, ref_count, ref_val, seen_mid = torch.zeros(8, dtype=torch.float64), 0, None, 0
    tgt = torch.zeros(n, dtype=torch.long, device=DEV)
    for step in range(3):
        q_a = F.normalize(rand(n, d, seed=20 + step), dim=1)
        # keys at graded distances from the queries so that ranks cover 0, 1..4 and >= 5
        noise = F.normalize(rand(n, d, seed=30 + step), dim=1)
        scale = torch.linspace(0.0, 5.0, n, device=DEV)[:, None]
        k_a = F.normalize(q_a + scale * noise, dim=1)
        kn_a = F.normalize(q_a + scale.flip(0) * noise, dim=1)
        q_m, k_m, kn_m = [F.normalize(rand(n, d, seed=40 + 3 * step + i), dim=1) for i in range(3)]
        logits, rows, ranks = ops.moco_logits_fwd(q_a, q_m, k_a, k_m, kn_a, kn_m, queue, T, materialize=True)
        loss3 = ops.moco_loss_fwd(rows, 2.0, 1.0, 1.0)
        l1, l2 = logits[0], logits[1]
        l1._rsp_ranks = l2._rsp_ranks = ranks
        l1._rsp_slot, l2._rsp_slot = 0, 1
        lpm, lnm = rows[0].unsqueeze(1), rows[1].unsqueeze(1)
        meters.update(loss3, (l1, l2), (lpm, lnm))
        a1, a5 = ref_accuracy(l1, tgt, (1, 5))
        n1, n5 = ref_accuracy(l2, tgt, (1, 5))
        m1, = ref_accuracy(torch.cat([lpm, lnm], dim=1), tgt, (1,))
        assert 0 < float(a1) <= float(a5) < 100          # the case is not degenerate
        seen_mid += int(((ranks > 0) & (ranks < 5)).sum())
        got = M.accuracy(l1, tgt, topk=(1, 5))
        assert float(got[0]) == float(a1) and float(got[1]) == float(a5)
        got = M.accuracy(l2, tgt, topk=(1, 5))
        assert float(got[0]) == float(n1) and float(got[1]) == float(n5)
        assert float(M.accuracy(torch.cat([lpm, lnm], dim=1), tgt, topk=(1,))[0]) == float(m1)   # plain-tensor route
        ref_val = torch.tensor([float(loss3[0]), float(loss3[1]), float(a1), float(a5), float(n1), float(n5),
                                float(loss3[2]), float(m1)], dtype=torch.float64)
        ref_sum += ref_val * n
        ref_count += n
    assert seen_mid > 0                                   # some positives sit between top-1 and top-5
    s = meters.summary()
    for i, name in enumerate(M.NAMES):
        assert abs(s[name]["val"] - float(ref_val[i])) <= 1e-4 * max(1.0, abs(float(ref_val[i]))), name
        assert abs(s[name]["avg"] - float(ref_sum[i]) / ref_count) <= 1e-4 * max(1.0, abs(float(ref_sum[i]) / ref_count)), name
    assert "Acc@1_A" in str(meters)
    meters.reset()
    assert meters.summary()["Loss"]["avg"] == 0.0


# ------------------------------------------------------------------------------------------------ BN / pool / head
@pytest.mark.parametrize("c,relu,res", [(64, True, False), (128, True, True), (512, False, False), (256, True, True),
                                        (192, True, False), (576, True, True)])
def test_bn_act_fwd_bwd(c, relu, res):
    ops = _ops()
    n, t, h, w = 2, 3, 5, 7
    x = rand(n, c, t, h, w, seed=1) * 2 + 0.5
    gamma, beta = rand(c, seed=2).abs() + 0.5, rand(c, seed=3)
    r = rand(n, c, t, h, w, seed=4) if res else None
    xb = x.bfloat16().float().requires_grad_(True)
    rb = r.bfloat16().float().requires_grad_(True) if res else None
    rm, rv = torch.zeros(c, device=DEV), torch.ones(c, device=DEV)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    y = F.batch_norm(xb, rm, rv, gr, br, True, 0.1, 1e-5)
    if res:
        y = y + rb
    if relu:
        y = F.relu(y)
    dy = rand(n, c, t, h, w, seed=5)
    y.backward(dy.bfloat16().float())

    xn = ops.to_ndhwc_bf16(x, c)
    s, ss = ops.bn_stats(xn)
    rm2, rv2 = torch.zeros(c, device=DEV), torch.ones(c, device=DEV)
    scale, shift, mean, invstd = ops.bn_finalize(s, ss, xn.numel() // c, gamma, beta, 1e-5, 0.1, rm2, rv2, c)
    torch.testing.assert_close(rm2, rm, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(rv2, rv, rtol=1e-4, atol=1e-5)
    rn = ops.to_ndhwc_bf16(r, c) if res else None
    out = ops.bn_act_fwd(xn, scale, shift, rn, relu)
    torch.testing.assert_close(ops.to_ncdhw_f32(out, c), y.detach(), rtol=2e-2, atol=2e-2)
    dyn = ops.to_ndhwc_bf16(dy, c)
    dx, dres, dgamma, dbeta = ops.bn_act_bwd(dyn, out, xn, mean, invstd, gamma, relu, res)
    torch.testing.assert_close(ops.to_ncdhw_f32(dx, c), xb.grad, rtol=3e-2, atol=3e-2)
    # ReLU masks are decided on bf16-rounded outputs: compare the reductions with a loose tolerance
    torch.testing.assert_close(dgamma, gr.grad, rtol=5e-2, atol=0.3)
    torch.testing.assert_close(dbeta, br.grad, rtol=5e-2, atol=0.3)
    if res:
        torch.testing.assert_close(ops.to_ncdhw_f32(dres, c), rb.grad, rtol=2e-2, atol=2e-2)


@pytest.mark.parametrize("k,s,p,shape", [((3, 3, 3), (2, 2, 2), (1, 1, 1), (2, 64, 6, 10, 10)),
                                          ((1, 2, 2), (1, 2, 2), (0, 0, 0), (2, 64, 4, 8, 8)),
                                          ((2, 2, 2), (2, 2, 2), (0, 0, 0), (1, 128, 4, 6, 6)),
                                          ((3, 3, 3), (1, 1, 1), (1, 1, 1), (1, 64, 3, 5, 5))])
def test_maxpool_fwd_bwd(k, s, p, shape):
    ops = _ops()
    x = rand(*shape, seed=1)
    x = torch.where(x < 0, torch.zeros_like(x), x)  # post-ReLU style ties at 0
    xb = x.bfloat16().float().requires_grad_(True)
    y = F.max_pool3d(xb, k, s, p)
    dy = rand(*y.shape, seed=2)
    y.backward(dy.bfloat16().float())
    c = shape[1]
    xn = ops.to_ndhwc_bf16(x, c)
    desc = ops.pool_desc(xn.shape, k, s, p)
    yo, idx = ops.maxpool3d_fwd(desc, xn)
    assert torch.equal(ops.to_ncdhw_f32(yo, c), y.detach())
    dx = ops.maxpool3d_bwd(desc, ops.to_ndhwc_bf16(dy, c), idx)
    torch.testing.assert_close(ops.to_ncdhw_f32(dx, c), xb.grad, rtol=1e-2, atol=1e-2)


@pytest.mark.parametrize("k,s,p,shape", [((3, 3, 3), (2, 2, 2), (1, 1, 1), (2, 64, 6, 12, 10)),
                                          ((3, 3, 3), (2, 2, 2), (1, 1, 1), (1, 64, 5, 23, 56)),
                                          ((1, 2, 2), (1, 2, 2), (0, 0, 0), (2, 64, 3, 8, 8)),
                                          ((2, 2, 2), (2, 2, 2), (0, 0, 0), (1, 128, 4, 6, 6)),
                                          ((2, 2, 2), (2, 2, 2), (0, 0, 0), (2, 512, 4, 7, 7)),
                                          ((3, 3, 3), (2, 2, 2), (1, 1, 1), (3, 64, 7, 16, 56)),
                                          ((3, 2, 2), (2, 2, 2), (1, 0, 0), (2, 128, 5, 9, 12)),
                                          ((3, 3, 3), (1, 1, 1), (1, 1, 1), (1, 256, 3, 5, 5))])
def test_bn_relu_maxpool_fused(k, s, p, shape):
    """The fused BN -> ReLU -> MaxPool kernels against (a) the two-step kernels (bit-identical pooled values) and (b) torch
    autograd in fp32."""
    ops = _ops()
    n, c = shape[0], shape[1]
    x = rand(*shape, seed=1) * 1.5 + 0.3
    gamma, beta = rand(c, seed=2).abs() + 0.5, rand(c, seed=3) * 0.5
    xn = ops.to_ndhwc_bf16(x, c)
    ssum, ssq = ops.bn_stats(xn)
    scale, shift, mean, invstd = ops.bn_finalize(ssum, ssq, xn.numel() // c, gamma, beta, 1e-5, 0.1, None, None, c)
    desc = ops.pool_desc(xn.shape, k, s, p)
    assert ops.bn_relu_maxpool_supported(desc)
    act = ops.bn_act_fwd(xn, scale, shift, None, True)
    y2, idx2 = ops.maxpool3d_fwd(desc, act)
    y1, idx1, xmax1 = ops.bn_relu_maxpool_fwd(desc, xn, scale, shift)
    y0 = ops.bn_relu_maxpool_fwd(desc, xn, scale, shift, aux=False)[0]       # the no-grad (key encoder) variant
    assert torch.equal(y0, y1)
    # x_max is the raw conv output at the argmax: the activation of it is the pooled value wherever that is positive
    # (the kernel uses one fused multiply-add: a bf16 ulp of slack)
    act_max = torch.relu(xmax1.float() * scale + shift)
    torch.testing.assert_close(act_max[y1.float() > 0], y1.float()[y1.float() > 0], rtol=1e-2, atol=1e-3)
    # same values bit for bit; the argmax may differ only where the winner is not unique after the activation (ReLU-clamped
    # windows, two inputs rounding to the same bf16) — positions whose gradient is masked or equivalent
    assert torch.equal(y1, y2)
    live = y2.float() > 0
    assert (idx1[live] == idx2[live]).float().mean().item() > 0.99
    dy = ops.to_ndhwc_bf16(rand(n, c, *y1.shape[1:4], seed=4), c)
    dx1, dgamma1, dbeta1 = ops.bn_relu_maxpool_bwd(desc, dy, idx1, xmax1, xn, scale, shift, mean, invstd, gamma)
    dpool = ops.maxpool3d_bwd(desc, dy, idx2)               # rounds the routed gradient to bf16 (the fused path does not)
    dx2, _, dgamma2, dbeta2 = ops.bn_act_bwd(dpool, act, xn, mean, invstd, gamma, True, False)
    # windows whose winner differs (activation-rounding ties, see above) route their gradient elsewhere: robust statistic
    scale_dx = dx2.float().abs().max().item()
    err12 = (dx1.float() - dx2.float()).abs()
    assert err12.median().item() <= 1e-2 * scale_dx
    assert (err12 > 3e-2 * scale_dx).float().mean().item() < 0.02
    # sums over thousands of positions of gradients that differ by one bf16 rounding each
    torch.testing.assert_close(dgamma1, dgamma2, rtol=3e-2, atol=0.3)
    torch.testing.assert_close(dbeta1, dbeta2, rtol=3e-2, atol=0.3)
    # torch fp32 statement of the same block
    xb = xn.float().permute(0, 4, 1, 2, 3).contiguous().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    yt = F.max_pool3d(F.relu(F.batch_norm(xb, None, None, gr, br, True, 0.1, 1e-5)), k, s, p)
    yt.backward(dy.float().permute(0, 4, 1, 2, 3))
    torch.testing.assert_close(ops.to_ncdhw_f32(y1, c), yt.detach(), rtol=2e-2, atol=2e-2)
    # argmax decisions are taken on bf16-rounded activations: a few windows may route differently -> robust statistic
    err = (ops.to_ncdhw_f32(dx1, c) - xb.grad).abs()
    assert err.median().item() < 1e-2 * xb.grad.abs().max().item()
    assert (err > 5e-2 * xb.grad.abs().max().item()).float().mean().item() < 0.02
    torch.testing.assert_close(dgamma1, gr.grad, rtol=5e-2, atol=0.5)
    torch.testing.assert_close(dbeta1, br.grad, rtol=5e-2, atol=0.5)


def test_head_fwd_bwd():
    ops = _ops()
    b, c, d = 5, 512, 128
    feat = rand(b, c, 1, 4, 4, seed=1).abs()
    w1, b1, w2, b2 = rand(d, c, seed=2, scale=0.05), rand(d, seed=3), rand(d, c, seed=4, scale=0.05), rand(d, seed=5)
    fb = feat.bfloat16().float().requires_grad_(True)
    params = [t.clone().requires_grad_(True) for t in (w1, b1, w2, b2)]
    pooled = fb.mean(dim=(2, 3, 4))
    o1 = F.normalize(F.linear(pooled, params[0], params[1]), dim=1)
    o2 = F.normalize(F.linear(pooled, params[2], params[3]), dim=1)
    g1, g2 = rand(b, d, seed=6), rand(b, d, seed=7)
    (o1 * g1).sum().backward(retain_graph=True)
    (o2 * g2).sum().backward()
    fn = ops.to_ndhwc_bf16(feat, c)
    out1, out2, pl, raw = ops.head_fwd(fn, c, w1, b1, w2, b2)
    torch.testing.assert_close(out1, o1.detach(), rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(out2, o2.detach(), rtol=1e-4, atol=1e-5)
    dw1, db1, dw2, db2, dfeat = ops.head_bwd(g1, g2, pl, raw, fn.shape, w1, w2)
    for got, ref in zip((dw1, db1, dw2, db2), params):
        torch.testing.assert_close(got, ref.grad, rtol=1e-3, atol=1e-5)
    torch.testing.assert_close(ops.to_ncdhw_f32(dfeat, c), fb.grad, rtol=2e-2, atol=1e-4)


# ------------------------------------------------------------------------------------------------ conv (tcgen05)
CONV_CASES = [
    (2, 64, 64, (4, 8, 8), (3, 3, 3), (1, 1, 1), (1, 1, 1)),
    (1, 64, 128, (3, 9, 7), (3, 3, 3), (1, 1, 1), (1, 1, 1)),
    (2, 128, 128, (4, 10, 10), (3, 3, 3), (2, 2, 2), (1, 1, 1)),
    (2, 64, 128, (4, 8, 8), (1, 1, 1), (2, 2, 2), (0, 0, 0)),
    (2, 128, 256, (2, 6, 6), (3, 3, 3), (1, 1, 1), (1, 1, 1)),
    (1, 3, 64, (6, 20, 20), (7, 7, 7), (1, 2, 2), (3, 3, 3)),
    # row-paired RGB stem: several 8-row iterations per frame, ragged last iteration, clipped frame taps at both ends
    (2, 3, 64, (9, 44, 36), (7, 7, 7), (1, 2, 2), (3, 3, 3)),
    (2, 3, 45, (3, 32, 28), (1, 7, 7), (1, 2, 2), (0, 3, 3)),
    # R(2+1)D-vcop stem: 83 mid channels = two 64-channel output groups (one stem launch each), ragged second group
    (2, 3, 83, (3, 32, 28), (1, 7, 7), (1, 2, 2), (0, 3, 3)),
    (1, 3, 128, (8, 40, 36), (7, 7, 7), (1, 2, 2), (3, 3, 3)),
    # wide rows (S3D-G at 224 x 224): column tiles of up to 60 output pixels, one stem launch each; ragged last tile
    (1, 3, 64, (2, 30, 224), (1, 7, 7), (1, 2, 2), (0, 3, 3)),
    (1, 3, 64, (3, 20, 122), (1, 7, 7), (1, 2, 2), (0, 3, 3)),
    (1, 3, 83, (2, 18, 136), (1, 7, 7), (1, 2, 2), (0, 3, 3)),
    (2, 3, 64, (4, 12, 12), (3, 3, 3), (1, 1, 1), (1, 1, 1)),
    (3, 256, 512, (2, 7, 7), (3, 3, 3), (2, 2, 2), (1, 1, 1)),
    # one-frame tensors (R3D-18 layer4): the outer frame taps only read padding and are skipped as whole K blocks
    (5, 128, 128, (1, 4, 4), (3, 3, 3), (1, 1, 1), (1, 1, 1)),
    (9, 64, 128, (1, 5, 4), (3, 3, 3), (1, 1, 1), (1, 1, 1)),
    # 3x3x3 unit-stride RGB stem (C3D conv1): the even/odd raw-row kernel, ragged H and odd T
    (1, 3, 64, (3, 10, 14), (3, 3, 3), (1, 1, 1), (1, 1, 1)),
    (2, 3, 64, (5, 33, 112), (3, 3, 3), (1, 1, 1), (1, 1, 1)),
    # shapes that take the direct (im2col-free) kernel: unit stride, >= 512 (Co=64) / 256 (Co%128==0) padded positions
    (2, 64, 64, (3, 28, 28), (3, 3, 3), (1, 1, 1), (1, 1, 1)),
    (1, 128, 128, (3, 20, 22), (3, 3, 3), (1, 1, 1), (1, 1, 1)),
    (2, 64, 192, (2, 24, 20), (1, 3, 3), (1, 1, 1), (0, 1, 1)),
    (1, 128, 64, (2, 30, 26), (3, 3, 3), (1, 1, 1), (1, 1, 1)),
    # temporal 3x1x1 filters of R(2+1)D: the dgrad into 192 / 320 stored channels takes the direct kernel (kh*kw = 1)
    (2, 144, 64, (5, 24, 20), (3, 1, 1), (1, 1, 1), (1, 0, 0)),
    (1, 288, 128, (3, 28, 28), (3, 1, 1), (1, 1, 1), (1, 0, 0)),
    (1, 64, 64, (7, 20, 20), (7, 1, 1), (1, 1, 1), (3, 0, 0)),
    # unit-stride 1x1x1 (S3D-G inception branches) through the same persistent TMA pipeline
    (2, 192, 64, (4, 14, 14), (1, 1, 1), (1, 1, 1), (0, 0, 0)),
    (1, 64, 128, (3, 28, 28), (1, 1, 1), (1, 1, 1), (0, 0, 0)),
    (1, 256, 192, (2, 20, 12), (1, 1, 1), (1, 1, 1), (0, 0, 0)),
    # strided pointwise / temporal filters (residual shortcuts of R3D-18 / R(2+1)D, R(2+1)D's 3x1x1 s(2,1,1)) at plane
    # sizes where the direct kernel takes their unit-stride counterparts
    (2, 64, 128, (4, 28, 28), (1, 1, 1), (2, 2, 2), (0, 0, 0)),
    (2, 64, 64, (3, 56, 56), (1, 1, 1), (1, 2, 2), (0, 0, 0)),
    (2, 64, 128, (6, 14, 14), (1, 1, 1), (2, 1, 1), (0, 0, 0)),
    (1, 128, 64, (5, 27, 29), (1, 1, 1), (2, 2, 2), (0, 0, 0)),
    (2, 256, 128, (8, 28, 28), (3, 1, 1), (2, 1, 1), (1, 0, 0)),
    (1, 64, 64, (7, 20, 20), (3, 1, 1), (2, 1, 1), (1, 0, 0)),
]


@pytest.mark.parametrize("n,ci,co,dims,k,s,p", CONV_CASES)
def test_conv3d_fprop_dgrad_wgrad(n, ci, co, dims, k, s, p):
    """bf16 inputs, fp32 accumulation: tolerance = 1e-2 of the tensor's max magnitude (bf16 output rounding is 2^-9)."""
    ops = _ops()
    x = rand(n, ci, *dims, seed=1)
    w = rand(co, ci, *k, seed=2, scale=(ci * k[0] * k[1] * k[2]) ** -0.5)
    bias = rand(co, seed=3)
    xr = x.bfloat16().float().requires_grad_(True)
    wr = w.bfloat16().float().requires_grad_(True)
    yref = F.conv3d(xr, wr, bias, s, p)
    dy = rand(*yref.shape, seed=4)
    yref.backward(dy.bfloat16().float())

    def close(got, ref, tol):
        err = (got.float() - ref.float()).abs().max().item()
        assert err <= tol * (ref.abs().max().item() + 1e-6), f"max err {err} vs scale {ref.abs().max().item()}"

    cis, cos = ops.pad_channels(ci), ops.pad_channels(co)
    xn = ops.to_ndhwc_bf16(x, cis)
    desc = ops.conv_desc(xn.shape, cos, k, s, p)
    stats = torch.zeros(2, cos, device=DEV)
    y = ops.conv3d_fprop(desc, xn, ops.conv3d_pack_weight(desc, w, 0), bias, stats=stats)
    close(ops.to_ncdhw_f32(y, co), yref.detach(), 1e-2)
    if ops.conv3d_fprop.stats_done:   # BN statistics fused into the epilogue describe the stored bf16 tensor
        yf = y.float().reshape(-1, cos)
        torch.testing.assert_close(stats[0], yf.sum(0), rtol=1e-3, atol=1e-2)
        torch.testing.assert_close(stats[1], (yf * yf).sum(0), rtol=1e-3, atol=1e-2)
    else:
        assert torch.count_nonzero(stats) == 0
    dyn = ops.to_ndhwc_bf16(dy, cos)
    dw = ops.conv3d_wgrad(desc, xn, dyn, w.shape)
    close(dw, wr.grad, 1e-2)
    if cis % 64 == 0:
        dx = ops.conv3d_dgrad(desc, dyn, ops.conv3d_pack_weight(desc, w, 1))
        close(ops.to_ncdhw_f32(dx, ci), xr.grad, 1e-2)


WGRAD_DIRECT_CASES = [
    # (n, ci, co, dims): 3x3x3 / 1x3x3 unit-stride filters through the plane-run wgrad (forced with rsp_debug_wgrad(2))
    (3, 64, 64, (4, 28, 28), (3, 3, 3), (1, 1, 1)),     # R3D-18 layer1 plane, several chunks per plane
    (2, 64, 64, (3, 9, 7), (3, 3, 3), (1, 1, 1)),       # one short chunk, Wp = 9 < 16
    (2, 128, 64, (2, 14, 14), (3, 3, 3), (1, 1, 1)),    # two ci chunks
    (1, 64, 128, (5, 12, 20), (3, 3, 3), (1, 1, 1)),    # two co chunks, odd frame count
    (2, 64, 64, (2, 16, 16), (1, 3, 3), (0, 1, 1)),     # separable spatial filter (kt = 1)
]


@pytest.mark.parametrize("n,ci,co,dims,k,p", WGRAD_DIRECT_CASES)
def test_conv3d_wgrad_direct(n, ci, co, dims, k, p):
    """Plane-run filter gradient (conv_wgrad_direct.cu) vs torch fp32 on bf16-rounded inputs, 1e-2 of the tensor max, and
    vs the generic kernel."""
    ops = _ops()
    from rspnet_b200 import _lib
    x = rand(n, ci, *dims, seed=1)
    w = rand(co, ci, *k, seed=2)
    xr = x.bfloat16().float()
    wr = w.clone().requires_grad_(True)
    yref = F.conv3d(xr, wr, None, (1, 1, 1), p)
    dy = rand(*yref.shape, seed=4)
    yref.backward(dy.bfloat16().float())
    xn = ops.to_ndhwc_bf16(x, ci)
    dyn = ops.to_ndhwc_bf16(dy, co)
    desc = ops.conv_desc(xn.shape, co, k, (1, 1, 1), p)
    lib = _lib.load()
    try:
        lib.rsp_debug_wgrad(2)
        dw = ops.conv3d_wgrad(desc, xn, dyn, w.shape)
        lib.rsp_debug_wgrad(1)
        dw_generic = ops.conv3d_wgrad(desc, xn, dyn, w.shape)
    finally:
        lib.rsp_debug_wgrad(0)
    scale = wr.grad.abs().max().item()
    assert (dw - wr.grad).abs().max().item() <= 1e-2 * scale
    assert (dw - dw_generic).abs().max().item() <= 2e-3 * scale   # same bf16 operands, different summation order


# ------------------------------------------------------------------------------------------------ S3D-G pieces
@pytest.mark.parametrize("c_l", [64, 208, 24])
def test_gate_fwd_bwd(c_l):
    ops = _ops()
    n, t, h, w = 3, 2, 5, 4
    x = rand(n, c_l, t, h, w, seed=1)
    wt, b = rand(c_l, c_l, 1, 1, 1, seed=2, scale=c_l ** -0.5), rand(c_l, seed=3)
    xr = x.bfloat16().float().requires_grad_(True)
    wr, br = wt.clone().requires_grad_(True), b.clone().requires_grad_(True)
    g = torch.sigmoid(F.conv3d(xr.mean(dim=(2, 3, 4), keepdim=True), wr, br))
    y = g * xr
    dy = rand(n, c_l, t, h, w, seed=4)
    y.backward(dy.bfloat16().float())
    cs = ops.pad_channels(c_l)
    xn = ops.to_ndhwc_bf16(x, cs)
    yo, pooled, gate = ops.gate_fwd(xn, c_l, wt, b)
    torch.testing.assert_close(ops.to_ncdhw_f32(yo, c_l), y.detach(), rtol=2e-2, atol=2e-2)
    dx, dw, db = ops.gate_bwd(ops.to_ndhwc_bf16(dy, cs), xn, c_l, wt, pooled, gate)
    torch.testing.assert_close(ops.to_ncdhw_f32(dx, c_l), xr.grad, rtol=2e-2, atol=2e-2)
    torch.testing.assert_close(dw, wr.grad.view(c_l, c_l), rtol=2e-2, atol=2e-3)
    torch.testing.assert_close(db, br.grad, rtol=2e-2, atol=2e-3)
    if cs != c_l:
        assert torch.count_nonzero(yo[..., c_l:]) == 0


def test_concat_channels_fwd_bwd():
    from rspnet_b200 import nn as rnn
    ops = _ops()
    widths = (64, 208, 48, 64)
    xs = [rand(2, c, 2, 3, 3, seed=i) for i, c in enumerate(widths)]
    xn = [ops.to_ndhwc_bf16(x).requires_grad_(True) for x in xs]
    out = rnn.concat_channels(xn, widths)
    ref = torch.cat([x.bfloat16().float() for x in xs], 1)
    assert out.shape[-1] == 384
    assert torch.equal(ops.to_ncdhw_f32(out, 384), ref)
    gout = rand(2, 384, 2, 3, 3, seed=9)
    out.backward(ops.to_ndhwc_bf16(gout, 384))
    off = 0
    for x, c in zip(xn, widths):
        assert torch.equal(ops.to_ncdhw_f32(x.grad, c), gout[:, off:off + c].bfloat16().float())
        off += c
