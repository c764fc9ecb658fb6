"""N > 1 path on CPU: two processes, gloo backend, frames of each batch sharded across ranks with one
all-reduce of the partial w_init sums per batch.  Uses the kernel-logic emulator (tests/emu)."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, emu_path, tmp, cc=False, coupled=False):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "tests"))
    import torch.distributed as dist
    import flowreg3d_b200 as F
    from flowreg3d_b200 import _lib
    _lib._select_for_tests(emu_path)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = np.load(ROOT / "tests" / "golden" / "sequence.npz")
    opts = F.OFOptions(min_level=3, iterations=8, update_lag=4, buffer_size=5, weight=[0.5, 0.5])
    v = g["video"][:, :12, :24, :28]
    r = g["ref"][:12, :24, :28]
    if coupled:   # options that couple the frames of a batch: temporal pre-filter + update_reference
        opts = F.OFOptions(min_level=3, iterations=8, update_lag=4, buffer_size=5, weight=[0.5, 0.5],
                           sigma=[[1.0, 1.0, 1.0, 0.8], [1.0, 1.0, 1.0, 0.8]], update_reference=True)
    if cc:   # rigid cross-correlation pre-alignment: single channel (as in the reference), per-rank estimator
        opts = F.OFOptions(min_level=3, iterations=8, update_lag=4, buffer_size=5, cc_initialization=True,
                           cc_hw=(20, 24), cc_up=10)
        v, r = v[..., :1], r[..., :1]
    reg, w, idx = F.compensate_arr_3D_sharded(v, r, opts, cc_prealign=cc)
    np.savez(os.path.join(tmp, f"r{rank}.npz"), reg=reg, w=w, idx=idx)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_with_frame_coupling_options(emu_backend, tmp_path):
    """Temporal pre-filter (sigma_t = 0.8: radius 3 frames, so every shard needs halo frames of its batch) and
    update_reference (the fixed volume re-averaged from ALL frames of a batch: all-reduced float64 partial sums) on a
    batch sharded over two ranks equal the single-process result."""
    from emu.build_emu import build
    import flowreg3d_b200 as F
    emu = str(build())
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), emu, str(tmp_path), False, True), nprocs=world, join=True)
    g = np.load(ROOT / "tests" / "golden" / "sequence.npz")
    opts = F.OFOptions(min_level=3, iterations=8, update_lag=4, buffer_size=5, weight=[0.5, 0.5], output_typename=None,
                       sigma=[[1.0, 1.0, 1.0, 0.8], [1.0, 1.0, 1.0, 0.8]], update_reference=True)
    v = g["video"][:, :12, :24, :28]
    r = g["ref"][:12, :24, :28]
    reg1, w1 = F.compensate_arr_3D(v, r, opts)
    seen = []
    for rank in range(world):
        d = np.load(tmp_path / f"r{rank}.npz")
        seen.extend(d["idx"].tolist())
        e = np.sqrt(((d["w"].astype(np.float64) - w1[d["idx"]]) ** 2).sum(-1))
        assert e.mean() <= 1e-5 and e.max() <= 1e-3, (rank, e.mean(), e.max())
        assert np.linalg.norm(d["reg"] - reg1[d["idx"]]) <= 1e-5 * np.linalg.norm(reg1[d["idx"]])
    assert sorted(seen) == list(range(v.shape[0]))


@pytest.mark.parametrize("cc", [False, True])
def test_two_rank_sharding_matches_single_process(emu_backend, tmp_path, cc):
    from emu.build_emu import build
    import flowreg3d_b200 as F
    emu = str(build())
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), emu, str(tmp_path), cc), nprocs=world, join=True)
    g = np.load(ROOT / "tests" / "golden" / "sequence.npz")
    opts = F.OFOptions(min_level=3, iterations=8, update_lag=4, buffer_size=5, weight=[0.5, 0.5],
                       output_typename=None)
    v = g["video"][:, :12, :24, :28]
    r = g["ref"][:12, :24, :28]
    if cc:
        opts = F.OFOptions(min_level=3, iterations=8, update_lag=4, buffer_size=5, cc_initialization=True,
                           cc_hw=(20, 24), cc_up=10, output_typename=None)
        v, r = v[..., :1], r[..., :1]
    reg1, w1 = F.compensate_arr_3D(v, r, opts, cc_prealign=cc)
    seen = []
    for rank in range(world):
        d = np.load(tmp_path / f"r{rank}.npz")
        seen.extend(d["idx"].tolist())
        # the all-reduced w_init sums associate differently from numpy's sequential float32 mean:
        # 1-ulp float32 differences in w_init, far below the 0.01 / 0.05 voxel tolerance
        e = np.sqrt(((d["w"].astype(np.float64) - w1[d["idx"]]) ** 2).sum(-1))
        assert e.mean() <= 1e-5 and e.max() <= 1e-3
        assert np.linalg.norm(d["reg"] - reg1[d["idx"]]) <= 1e-5 * np.linalg.norm(reg1[d["idx"]])
    assert sorted(seen) == list(range(v.shape[0]))   # every frame processed exactly once
    # batches of 5 and 2 frames over 2 ranks: 3+2 and 1+1
    assert len(np.load(tmp_path / "r0.npz")["idx"]) == 4 and len(np.load(tmp_path / "r1.npz")["idx"]) == 3


def _pipeline_worker(rank, world, port, emu_path, tmp, zslab=False):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "tests"))
    import torch.distributed as dist
    import flowreg3d_b200 as F
    from flowreg3d_b200 import _lib, device as dev
    from flowreg3d_b200.multigpu import get_displacement_pipelined
    _lib._select_for_tests(emu_path)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = np.load(ROOT / "tests" / "golden" / "flow_small.npz")
    fixed, moving = g["fixed"][:16, :28, :32].astype(np.float32), g["moving"][:16, :28, :32].astype(np.float32)
    fp = F.FlowParams(alpha=(0.25, 0.3, 0.2), update_lag=3, iterations=14, min_level=0, levels=100, eta=0.8,
                      a_smooth=1.0, a_data=0.45)
    reg = F.Registration(fixed.shape[:3], fixed.shape[3], fp, max_batch=2)
    reg.set_reference(fixed)
    mv = np.stack([moving, np.roll(moving, 1, 2)], 0)
    if zslab:
        from flowreg3d_b200.multigpu import get_displacement_zslab
        flow = get_displacement_zslab(reg, mv, min_slots=0)                  # z-slabs + halo exchange on every level
    else:
        flow = get_displacement_pipelined(reg, mv, min_slots=0, n_chunks=5)  # pipeline every level
    reg.sync()
    np.save(os.path.join(tmp, f"p{rank}.npy"), dev.to_host(flow))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sweep_pipelined_solve_is_bit_identical(emu_backend, tmp_path, world):
    """One volume on several ranks: sweeps split over the ranks, increments streamed rank -> rank+1 in
    hyperplane blocks.  Every rank ends with exactly the single-process flow (iterations 14 / lag 3 gives
    uneven sweep ranges: 2 ranks -> [0,9) [9,14); 3 ranks -> [0,6) [6,12) [12,14))."""
    from emu.build_emu import build
    import flowreg3d_b200 as F
    from flowreg3d_b200 import device as dev
    from flowreg3d_b200.multigpu import pipeline_schedule, sweep_partition
    assert sweep_partition(14, 3, 2) == [(0, 9), (9, 14)]
    assert sweep_partition(14, 3, 3) == [(0, 6), (6, 12), (12, 14)]
    assert sweep_partition(10, 5, 4) == [(0, 5), (5, 10), (10, 10), (10, 10)]
    # schedule invariants: every rank runs each of its waves exactly once, in order; receives of rank r are
    # the sends of rank r-1; a wave never runs before the hyperplanes it reads have arrived
    S, T = 40, 14
    parts = sweep_partition(T, 3, 3)
    active, sched = pipeline_schedule(S, T, parts, 5)
    for i, r in enumerate(active):
        t0, t1 = parts[r]
        st = sched[r]
        assert st[0][1] == 2 * t0 and st[-1][2] == S - 1 + 2 * (t1 - 1) + 1
        assert all(a[2] == b[1] for a, b in zip(st, st[1:]))
        got = 0 if i else S
        for rcv, q0, q1, snd in st:
            if rcv is not None:
                assert rcv[0] == got
                got = rcv[1]
            assert q1 <= got + 2 * t0 - 1 or got == S
        if i:
            assert [s[0] for s in st if s[0]] == [s[3] for s in sched[active[i - 1]] if s[3]]
    emu = str(build())
    mp.spawn(_pipeline_worker, args=(world, _free_port(), emu, str(tmp_path)), nprocs=world, join=True)
    g = np.load(ROOT / "tests" / "golden" / "flow_small.npz")
    fixed, moving = g["fixed"][:16, :28, :32].astype(np.float32), g["moving"][:16, :28, :32].astype(np.float32)
    fp = F.FlowParams(alpha=(0.25, 0.3, 0.2), update_lag=3, iterations=14, min_level=0, levels=100, eta=0.8,
                      a_smooth=1.0, a_data=0.45)
    reg = F.Registration(fixed.shape[:3], fixed.shape[3], fp, max_batch=2)
    reg.set_reference(fixed)
    ref = dev.to_host(reg.get_displacement(np.stack([moving, np.roll(moving, 1, 2)], 0)))
    for rank in range(world):
        assert np.array_equal(np.load(tmp_path / f"p{rank}.npy"), ref), rank


@pytest.mark.parametrize("world", [2, 3])
def test_zslab_solve_with_halo_exchange_is_bit_identical(emu_backend, tmp_path, world):
    """One volume on several ranks, z-slab decomposition: every rank sweeps its own planes wave by wave and trades
    its boundary planes with both z-neighbours after every wave (S + 2(T-1) exchanges per level).  Same update order
    as one process -> bit-identical flows on every rank (16 planes over 3 ranks gives uneven slabs 6/5/5; the
    coarsest levels have fewer planes than ranks on some configurations and are then solved redundantly)."""
    from emu.build_emu import build
    import flowreg3d_b200 as F
    from flowreg3d_b200 import device as dev
    emu = str(build())
    mp.spawn(_pipeline_worker, args=(world, _free_port(), emu, str(tmp_path), True), nprocs=world, join=True)
    g = np.load(ROOT / "tests" / "golden" / "flow_small.npz")
    fixed, moving = g["fixed"][:16, :28, :32].astype(np.float32), g["moving"][:16, :28, :32].astype(np.float32)
    fp = F.FlowParams(alpha=(0.25, 0.3, 0.2), update_lag=3, iterations=14, min_level=0, levels=100, eta=0.8,
                      a_smooth=1.0, a_data=0.45)
    reg = F.Registration(fixed.shape[:3], fixed.shape[3], fp, max_batch=2)
    reg.set_reference(fixed)
    ref = dev.to_host(reg.get_displacement(np.stack([moving, np.roll(moving, 1, 2)], 0)))
    for rank in range(world):
        assert np.array_equal(np.load(tmp_path / f"p{rank}.npy"), ref), rank
