"""View-batch sharding on CPU: the partition rule, and a world_size-2 gloo run in which every rank prepares its
block of views with the host-side renderer mirror and the gathered per-view digests must equal a single-process run."""
import os
import socket

import numpy as np
import pytest

from square_wheel_b200 import batch


def test_partition_is_contiguous_balanced_and_complete():
    for num_views in (0, 1, 7, 64, 8192, 8193):
        for world in (1, 2, 3, 4, 8):
            seen = []
            sizes = []
            for rank in range(world):
                r = batch.partition_views(num_views, world, rank)
                assert r.step == 1
                seen.extend(r)
                sizes.append(len(r))
            assert seen == list(range(num_views))
            assert max(sizes) - min(sizes) <= 1
    for i in range(64):
        assert i in batch.partition_views(64, 8, batch.owner_of_view(i, 64, 8))
    with pytest.raises(ValueError):
        batch.partition_views(8, 2, 2)


def test_resampled_path_is_seeded_and_dense():
    base = np.zeros((4, 5), np.float32)
    base[:, 0] = [0, 100, 100, 0]
    base[:, 1] = [0, 0, 100, 100]
    base[:, 3] = [0.0, 1.5, 3.1, -1.6]
    a = batch.resample_camera_path(base, 16)
    b = batch.resample_camera_path(base, 16)
    c = batch.resample_camera_path(base, 16, seed=1)
    assert a.shape == (64, 5)
    assert np.array_equal(a, b) and not np.array_equal(a, c)
    assert np.abs(a[16, :3] - base[1, :3]).max() <= 4.0  # jitter bound


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, num_views, out_dir):
    import torch.distributed as dist

    from square_wheel_b200 import scene as S

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sc = S.Scene(seed=11, grid=(8, 8, 2), num_rooms=5, num_materials=4, texture_size_log2=5)
        poses = batch.resample_camera_path(sc.camera_path(8), num_views // 8)
        mine = batch.partition_views(len(poses), world, rank)
        hr = S.HostRenderer(sc)
        digests = []
        for i in mine:
            hf = hr.build_frame(320, 180, poses[i])
            digests.append(batch.frame_digest(hf.jobs, hf.cmds, hf.polygon_points))
        hr.close()
        dist.barrier()
        elapsed = batch.max_over_ranks(1.0 + rank, dist)  # the slowest rank defines the elapsed time
        all_digests = batch.gather_digests(digests, mine, len(poses), dist)
        if rank == 0:
            np.save(os.path.join(out_dir, "digests.npy"), np.array(all_digests))
            np.save(os.path.join(out_dir, "elapsed.npy"), np.array([elapsed]))
        sc.close()
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_batch_matches_single_process(tmp_path):
    import torch.multiprocessing as mp

    from square_wheel_b200 import scene as S

    num_views, world = 24, 2
    mp.spawn(_worker, args=(world, _free_port(), num_views, str(tmp_path)), nprocs=world, join=True)
    got = list(np.load(tmp_path / "digests.npy"))
    assert float(np.load(tmp_path / "elapsed.npy")[0]) == 2.0

    sc = S.Scene(seed=11, grid=(8, 8, 2), num_rooms=5, num_materials=4, texture_size_log2=5)
    poses = batch.resample_camera_path(sc.camera_path(8), num_views // 8)
    hr = S.HostRenderer(sc)
    want = []
    for p in poses:
        hf = hr.build_frame(320, 180, p)
        want.append(batch.frame_digest(hf.jobs, hf.cmds, hf.polygon_points))
    hr.close()
    sc.close()
    assert got == want
    assert len(set(want)) > 1  # the views really differ
