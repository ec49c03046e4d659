"""Reference-sharded selector and pose-sharded refiner with 2 ranks: one rank per GPU over NCCL when
the box has >= 2 GPUs (the collectives are then captured into the select stage's CUDA graph and ride
the branch streams); on a 1-GPU box both ranks share cuda:0 and talk over gloo (tensors staged through
the host, eager)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path[:0] = [os.path.dirname(here), here]
    multi = torch.cuda.device_count() >= world
    if multi:
        torch.cuda.set_device(rank)
        dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    else:
        dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from golden import cases
        from gen6d_b200 import ops
        from gen6d_b200.dist import Comm, pose_shard, shard_selector
        from gen6d_b200.network import name2network
        from gen6d_b200.weights import seeded_state_dict
        torch.cuda.set_device(rank if multi else 0)
        comm = Comm()
        assert comm.capturable == multi

        def build(name, cfg):
            net = name2network[name](cfg)
            net.load_state_dict(seeded_state_dict(net, cases.WEIGHT_SEED))
            return net.cuda().eval()

        c = cases.selector_case(rfn=8, an=5)
        full = build('selector', c['cfg'])
        full.load_ref_imgs(c['ref_imgs'], c['ref_poses'], c['object_center'], c['object_vert'])
        want = full.select_que_imgs(c['que_imgs'])
        sharded = shard_selector(build('selector', c['cfg']), comm)
        sharded.load_ref_imgs(c['ref_imgs'], c['ref_poses'], c['object_center'], c['object_vert'])
        assert sharded.ref_shape == (4, 5) and sharded.rfn_total == 8
        got = sharded.select_que_imgs(c['que_imgs'])
        got = sharded.select_que_imgs(c['que_imgs'])            # second call: graph replay on the NCCL path
        assert got['ref_idx'].tolist() == want['ref_idx'].tolist()
        np.testing.assert_allclose(got['scores'], want['scores'], atol=2e-4)       # exact statistics: not ~0.1 off
        np.testing.assert_allclose(got['angles'], want['angles'], atol=2e-4)

        rc = cases.refiner_case(qn=2)
        rfr = build('refiner', {})
        dev = lambda a: torch.from_numpy(a).cuda()
        que = ops.preprocess_u8(dev(rc['que_imgs']), 4, True)
        ref = ops.preprocess_u8(dev(rc['ref_imgs']), 4, True)
        args = [que, dev(rc['que_Ks']), dev(rc['que_poses']), ref, dev(rc['ref_Ks']), dev(rc['ref_poses'])]
        whole = rfr._forward_nhwc(*args)
        split = pose_shard(rfr._forward_nhwc, args, comm)
        np.testing.assert_allclose(split.cpu().numpy(), whole.cpu().numpy(), atol=1e-5)
        q.put((rank, 'ok nccl' if multi else 'ok gloo'))
    except Exception as e:  # noqa: BLE001
        import traceback
        q.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


def test_sharded_selector_and_pose_shard_match_unsharded():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1].startswith('ok') for r in res), res
    print('backend:', res[0][1])
