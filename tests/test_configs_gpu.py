"""Parity on the other BASELINE.json configurations (sizes the oracle finishes in seconds) and on
edge cases: config 1 (selector, 32 refs x 12 rotation bins), detector with 64 references and a
2-frame batch at a non-/32 frame size, refiner batches below / above the small-M FC switch (qn = 3,
qn = 9), and the size-independent properties used where the oracle is too slow."""
import numpy as np
import pytest
import torch

from golden import cases
from gen6d_b200.network import name2network
from gen6d_b200.weights import seeded_state_dict
from oracle import gen6d_oracle as O

pytestmark = pytest.mark.gpu


def build(name, cfg):
    net = name2network[name](cfg)
    sd = seeded_state_dict(net, cases.WEIGHT_SEED)
    net.load_state_dict(sd, strict=True)
    return net.cuda().eval(), sd


def test_config1_selector_32refs_12bins():
    """BASELINE configs[0]: selector forward, 1 query 128x128, 32 refs, 12 rotation bins."""
    c = cases.selector_case(seed=51, rfn=32, an=12)
    net, sd = build('selector', c['cfg'])
    assert tuple(net.angle_predict[0].weight.shape) == (512, 515 * 12, 1)
    net.load_ref_imgs(c['ref_imgs'], c['ref_poses'], c['object_center'], c['object_vert'])
    feats, embed = O.sel_load_refs(sd, cases.u8_to_nchw(c['ref_imgs']), torch.from_numpy(c['ref_poses']),
                                   torch.from_numpy(c['object_center']), torch.from_numpy(c['object_vert']))
    logits, angles = O.sel_forward(sd, cases.u8_to_nchw(c['que_imgs']), feats, embed)
    res = net.select_que_imgs(c['que_imgs'])
    top2 = torch.topk(logits, 2, 1)[0]
    margin = float(top2[0, 0] - top2[0, 1])
    print('config-1 selector margin (oracle top-1 minus top-2):', margin)
    assert margin > 0.2                      # seed 51 was chosen for a clear margin (0.24): the index check below is meaningful
    np.testing.assert_allclose(res['scores'], logits.numpy(), atol=5e-4)
    idx, ang = O.sel_select(logits, angles)
    assert res['ref_idx'].tolist() == idx.tolist()                              # bit-exact viewpoint, unconditionally
    np.testing.assert_allclose(res['angles'], angles.numpy()[np.arange(1), res['ref_idx']], atol=5e-4)


def test_detector_64refs_batch2_odd_size():
    """64 reference views (two lanes-per-ref passes in the fused head), qn = 2, frame 104x136."""
    c = cases.detector_case(seed=63, rfn=64, hq=104, wq=136, qn=2)      # seed 63: oracle margins 0.17 / 0.13 on the two frames
    net, sd = build('detector', {'name': 'd', 'network': 'detector', **c['cfg']})
    net.load_ref_imgs(c['ref_imgs'])
    ref_feats = O.det_load_refs(sd, cases.u8_to_nchw(c['ref_imgs']))
    want = O.det_detect(sd, c['cfg'], cases.u8_to_nchw(c['que_imgs']), ref_feats)
    got = net.detect_impl(cases.u8_to_nchw(c['que_imgs']).cuda())
    np.testing.assert_allclose(got['scores'].cpu().numpy(), want['scores'].numpy(), atol=5e-4)
    pos, scl, idx = O.det_parse(want['scores'], want['select_pr_scale'], want['select_pr_offset'])
    top2 = torch.topk(want['scores'].flatten(1), 2, 1)[0]
    margins = (top2[:, 0] - top2[:, 1]).tolist()
    print('detector margins', margins)
    ws = want['scores'].shape[-1]
    sel = got['que_select_id'].cpu()
    assert min(margins) > 0.1
    for qi in range(len(margins)):
        assert int(sel[qi, 1] * ws + sel[qi, 0]) == int(idx[qi])               # bit-exact detection cell, both frames
    res = net.detect_que_imgs(c['que_imgs'])
    np.testing.assert_allclose(res['scales'], scl.numpy(), rtol=3e-3)


@pytest.mark.parametrize('qn', [3, 9])
def test_refiner_batches(qn):
    """qn = 3 uses the weight-streaming FC kernel (M <= 8), qn = 9 the GEMM path; both must agree
    with the oracle and with each other on the shared poses."""
    c = cases.refiner_case(seed=71, qn=qn)
    net, sd = build('refiner', {})
    T = torch.from_numpy
    n_or = min(qn, 3)          # the oracle on 3 poses takes ~3 s; the rest is checked by consistency
    want = O.ref_forward(sd, cases.u8_to_nchw(c['que_imgs'][:n_or]), T(c['que_Ks'][:n_or]), T(c['que_poses'][:n_or]),
                         cases.u8_to_nchw(c['ref_imgs'][:n_or]), T(c['ref_Ks'][:n_or]), T(c['ref_poses'][:n_or]), 32)
    data = {'que_imgs_info': {'imgs': cases.u8_to_nchw(c['que_imgs']).cuda(), 'Ks_in': T(c['que_Ks']).cuda(),
                              'poses_in': T(c['que_poses']).cuda()},
            'ref_imgs_info': {'imgs': cases.u8_to_nchw(c['ref_imgs']).cuda(), 'Ks': T(c['ref_Ks']).cuda(),
                              'poses': T(c['ref_poses']).cuda()}, 'inference': True}
    res = net(data)
    np.testing.assert_allclose(res['rotation'][:n_or].cpu().numpy(), want['rotation'].numpy(), atol=3e-4)
    np.testing.assert_allclose(res['offset'][:n_or].cpu().numpy(), want['offset'].numpy(), atol=3e-4)
    np.testing.assert_allclose(res['scale'][:n_or].cpu().numpy(), want['scale'].numpy(), atol=3e-4)
    # poses are independent (per-sample InstanceNorm): a batch must equal its members run alone
    one = {k: ({kk: vv[qn - 1:qn] for kk, vv in v.items()} if isinstance(v, dict) else v) for k, v in data.items()}
    alone = net(one)
    np.testing.assert_allclose(alone['rotation'].cpu().numpy(), res['rotation'][qn - 1:].cpu().numpy(), atol=2e-5)


def test_selector_score_properties_at_scale():
    """Size-independent properties of the S2 kernel at BASELINE configs[3] per-GPU size (64 refs x 36
    bins = 2304 slices, 1.6 GB): scale covariance score(a*q) = a*score(q), and agreement of the
    3-level pass with the per-level kernel."""
    from gen6d_b200 import ops
    S = 2304
    g = torch.Generator(device='cuda').manual_seed(3)
    refs = [torch.rand(S, P, 512, device='cuda', generator=g) for P in (256, 64, 16)]
    qs = [torch.rand(P, 512, device='cuda', generator=g) for P in (256, 64, 16)]
    s3 = ops.sel_corr_score3(refs, qs)
    for l in range(3):
        s1 = ops.sel_corr_score(refs[l], qs[l])
        np.testing.assert_allclose(s3[l].cpu().numpy(), s1.cpu().numpy(), rtol=2e-6)
    s3b = ops.sel_corr_score3(refs, [q * 4.0 for q in qs])
    np.testing.assert_allclose(s3b.cpu().numpy(), 4.0 * s3.cpu().numpy(), rtol=1e-6)   # power-of-two scale: exact up to sum order


def test_config4_angle_bins_36():
    """BASELINE configs[3] per-shard shape in miniature: 36 rotation bins (angle_predict 515*36 wide)."""
    c = cases.selector_case(seed=81, rfn=8, an=36)
    net, sd = build('selector', c['cfg'])
    net.load_ref_imgs(c['ref_imgs'], c['ref_poses'], c['object_center'], c['object_vert'])
    feats, embed = O.sel_load_refs(sd, cases.u8_to_nchw(c['ref_imgs']), torch.from_numpy(c['ref_poses']),
                                   torch.from_numpy(c['object_center']), torch.from_numpy(c['object_vert']))
    logits, angles = O.sel_forward(sd, cases.u8_to_nchw(c['que_imgs']), feats, embed)
    res = net.select_que_imgs(c['que_imgs'])
    np.testing.assert_allclose(res['scores'], logits.numpy(), atol=5e-4)
    top2 = torch.topk(logits, 2, 1)[0]
    margin = float(top2[0, 0] - top2[0, 1])
    print('36-bin selector margin (oracle):', margin)
    assert margin > 1.0                      # seed 81: 1.69
    assert res['ref_idx'].tolist() == torch.argmax(logits, 1).tolist()          # bit-exact viewpoint, unconditionally


def test_config5_refiner_batch32_is_poses_independent():
    """BASELINE configs[4] per-GPU batch (256 poses / 8 GPUs = 32): the batched forward must equal
    its members run one at a time (per-sample InstanceNorm, no cross-pose term)."""
    c = cases.refiner_case(seed=91, qn=32)
    net, _ = build('refiner', {})
    T = torch.from_numpy
    data = {'que_imgs_info': {'imgs': cases.u8_to_nchw(c['que_imgs']).cuda(), 'Ks_in': T(c['que_Ks']).cuda(),
                              'poses_in': T(c['que_poses']).cuda()},
            'ref_imgs_info': {'imgs': cases.u8_to_nchw(c['ref_imgs']).cuda(), 'Ks': T(c['ref_Ks']).cuda(),
                              'poses': T(c['ref_poses']).cuda()}, 'inference': True}
    res = net(data)
    for qi in (0, 17, 31):
        one = {k: ({kk: vv[qi:qi + 1] for kk, vv in v.items()} if isinstance(v, dict) else v) for k, v in data.items()}
        alone = net(one)
        for key in ('rotation', 'offset', 'scale'):
            np.testing.assert_allclose(alone[key].cpu().numpy(), res[key][qi:qi + 1].cpu().numpy(), atol=3e-5)
