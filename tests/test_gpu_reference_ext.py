"""GPU: three-way bit-exact index parity — (1) the REFERENCE'S OWN CUDA kernels (oracle/_ref, built from the
sources under /root/reference by oracle/build_ref.py), (2) the C restatement the oracle uses, (3) our kernels.
This is what pins the oracle's index ops to the real reference on the target hardware."""
import numpy as np
import pytest
import torch

from oracle import build_ref
from oracle import genpose_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ref_ext():
    mod = build_ref.load_ref()
    if mod is None:
        pytest.skip("oracle/_ref/pointnet2_cuda_ref.so not present (build it where /root/reference exists)")
    return mod


def _clouds(seed, B, n):
    rs = np.random.RandomState(seed)
    xyz = (rs.standard_normal((B, n, 3)) * 0.05 + np.array([0.1, -0.1, 0.9])).astype(np.float32)
    k = max(1, n // 7)
    xyz[1] = np.concatenate([np.tile(xyz[1, :k], (n // k, 1)), xyz[1, : n % k]], 0)    # tiled duplicates
    xyz[2] = xyz[2, 0]                                                                  # fully degenerate
    return xyz


@pytest.mark.parametrize("n,m", [(1024, 512), (512, 256), (256, 128), (1000, 200), (100, 100)])
def test_fps_three_way(ref_ext, n, m):
    from genpose_b200 import ops
    xyz = _clouds(n + m, 4, n)
    d = torch.from_numpy(xyz).cuda()
    idx = torch.zeros(4, m, dtype=torch.int32, device="cuda")
    temp = torch.full((4, n), 1e10, dtype=torch.float32, device="cuda")
    ref_ext.furthest_point_sampling_wrapper(4, n, m, d, temp, idx)
    torch.cuda.synchronize()
    ref = idx.cpu().numpy()
    assert np.array_equal(O.furthest_point_sample(torch.from_numpy(xyz), m).numpy(), ref), "C oracle != reference kernel"
    assert np.array_equal(ops.furthest_point_sample(d, m).cpu().numpy(), ref), "our kernel != reference kernel"


@pytest.mark.parametrize("n,m,r,ns", [(1024, 512, 0.02, 16), (1024, 512, 0.04, 32), (512, 256, 0.04, 16),
                                      (512, 256, 0.08, 32), (256, 128, 0.08, 16), (256, 128, 0.16, 32)])
def test_ball_query_three_way(ref_ext, n, m, r, ns):
    from genpose_b200 import ops
    xyz = _clouds(n + ns, 4, n)
    new_xyz = np.ascontiguousarray(xyz[:, :m])
    d, dn = torch.from_numpy(xyz).cuda(), torch.from_numpy(new_xyz).cuda()
    idx = torch.zeros(4, m, ns, dtype=torch.int32, device="cuda")
    ref_ext.ball_query_wrapper(4, n, m, r, ns, dn, d, idx)
    torch.cuda.synchronize()
    ref = idx.cpu().numpy()
    assert np.array_equal(O.ball_query(r, ns, torch.from_numpy(xyz), torch.from_numpy(new_xyz)).numpy(), ref)
    assert np.array_equal(ops.ball_query(r, ns, d, dn).cpu().numpy(), ref)


def test_gather_group_three_way(ref_ext):
    from genpose_b200 import ops
    rs = np.random.RandomState(0)
    pts = torch.from_numpy(rs.standard_normal((3, 19, 200)).astype(np.float32)).cuda()
    idx = torch.from_numpy(rs.randint(0, 200, (3, 50)).astype(np.int32)).cuda()
    out = torch.zeros(3, 19, 50, device="cuda")
    ref_ext.gather_points_wrapper(3, 19, 200, 50, pts, idx, out)
    assert torch.equal(ops.gather_points(pts, idx), out)
    gidx = torch.from_numpy(rs.randint(0, 200, (3, 20, 8)).astype(np.int32)).cuda()
    gout = torch.zeros(3, 19, 20, 8, device="cuda")
    ref_ext.group_points_wrapper(3, 19, 200, 20, 8, pts, gidx, gout)
    assert torch.equal(ops.group_points(pts, gidx), gout)
