"""CUDA-vs-reference parity AT THE SIZES THE BENCH PUBLISHES (BASELINE.json configs 2 and 4).

config 2: one frame of the 2-channel 32x512x512 recording at OFOptions defaults, compared with
  (i) the golden computed by the LIVE reference (tests/golden/config2.npz, gen_golden.py:gen_config2) and
  (ii) the oracle run on the spot at full size (~30 s of CPU),
for both solver state precisions (float64 default, float32 option) and for the factored gather option.
config 4: a z-heavy reduced volume (64x256x256, min_level 2: finest level 41x164x164) against the oracle.

Tolerances are the north star's (mean EPE <= 0.01, max <= 0.05 voxel, relative L2 of the corrected volume
<= 1e-4) tightened to what the float64 path is expected to hold (1e-4 / 5e-3 / 1e-5).
"""
import numpy as np
import pytest

from conftest import epe_stats, rel_l2
from oracle import oracle as O
from tests_inputs import smooth_flow, synth_volume

pytestmark = pytest.mark.gpu

SHAPE2 = (32, 512, 512)


@pytest.fixture(scope="module")
def config2_inputs(golden):
    Z, Y, X = SHAPE2
    ref = np.stack([synth_volume(SHAPE2, 10 + c) for c in range(2)], -1)
    g = smooth_flow(SHAPE2, 1000, 2.0, 12.0)
    r64 = ref.astype(np.float64)
    mov = O.imregister_wrapper(r64, -g[..., 0], -g[..., 1], -g[..., 2], r64, "linear")
    mov = (mov + 0.01 * np.random.default_rng(2000).standard_normal(mov.shape)).astype(np.float32)
    gd = golden("config2")
    chk = gd["moving_checksum"]
    assert abs(mov.astype(np.float64).sum() - chk[0]) <= 1e-6 * abs(chk[0]) and float(mov[7, 11, 13, 1]) == chk[1]
    return ref, mov, g


@pytest.fixture(scope="module")
def config2_oracle(config2_inputs):
    """The oracle at full size: pre-filter, get_displacement at OFOptions defaults, cubic compensation warp."""
    ref, mov, _ = config2_inputs
    r64 = ref.astype(np.float64)
    sigma = np.array([[1.0, 1.0, 1.0, 0.1]] * 2)
    fp = O.preprocess(r64, sigma)
    mp = O.preprocess(mov[None], sigma, r64)[0]
    flow = O.get_displacement(fp, mp, alpha=(0.25,) * 3, update_lag=5, iterations=100, min_level=5, levels=100,
                              eta=0.8, a_smooth=1.0, a_data=0.45, weight=np.full(ref.shape, 0.5))
    f32 = flow.astype(np.float32)
    reg = O.imregister_wrapper(mov, f32[..., 0], f32[..., 1], f32[..., 2], r64, "cubic")
    return f32, np.asarray(reg, np.float32)


def _cuda_config2(ref, mov, state, factored=False):
    import flowreg3d_b200 as F
    from flowreg3d_b200 import _lib, core, device as dev
    old = core.STATE_DTYPE
    core.STATE_DTYPE = np.float32 if state == "f32" else np.float64
    try:
        seq = F.SequenceCorrector(ref, F.OFOptions(buffer_size=1), max_batch=1)
    finally:
        core.STATE_DTYPE = old
    if factored:
        core._check(seq.reg.ctx.h, seq.reg.ctx.lib.fr3d_set_option(seq.reg.ctx.h, _lib.OPT_WARP_FACTORED, 1))
    proc = seq.reg.preprocess(mov[None], seq.lo, seq.den)
    flow_d = seq.reg.get_displacement(proc)
    reg_d = seq.reg.compensate(mov[None], flow_d)
    flow, reg = dev.to_host(flow_d)[0], dev.to_host(reg_d)[0]
    seq.close()
    return flow, reg


# (state, factored gather) -> (mean EPE, max EPE, rel-L2) bounds against the reference.
# float64 state: the product default, expected at float64 rounding of the reference (the final flow is float32).
# float32 state: SURVEY 7.3-D "safe" row; round 1 measured 7.8e-6 / 1.1e-2 / 1.0e-5 against the float64 state here.
BOUNDS = {("f64", False): (1e-4, 5e-3, 1e-5), ("f32", False): (1e-3, 5e-2, 1e-4), ("f64", True): (1e-4, 5e-3, 1e-5)}


@pytest.mark.parametrize("state,factored", list(BOUNDS))
def test_config2_frame_vs_live_golden_and_oracle(cuda_backend, golden, config2_inputs, config2_oracle, state, factored):
    ref, mov, g = config2_inputs
    flow, reg = _cuda_config2(ref, mov, state, factored)
    gd = golden("config2")
    bm, bx, bl = BOUNDS[(state, factored)]
    # (i) live reference, sub-sampled golden
    mean, mx = epe_stats(flow[::2, ::4, ::4], gd["flow_s"])
    l2 = rel_l2(reg[::2, ::4, ::4], gd["reg_s"])
    print(f"config2 {state} factored={factored} vs LIVE reference: EPE mean {mean:.2e} max {mx:.2e}, rel-L2 {l2:.2e}")
    assert mean <= bm and mx <= bx and l2 <= bl, (state, factored, mean, mx, l2)
    # (ii) oracle, every voxel
    oflow, oreg = config2_oracle
    mean, mx = epe_stats(flow, oflow)
    l2 = rel_l2(reg, oreg)
    print(f"config2 {state} factored={factored} vs oracle (all voxels): EPE mean {mean:.2e} max {mx:.2e}, rel-L2 {l2:.2e}")
    assert mean <= bm and mx <= bx and l2 <= bl, (state, factored, mean, mx, l2)
    # sanity: the estimate explains the synthetic motion as well as the reference does (0.079 mean EPE vs ground truth)
    assert epe_stats(flow, g)[0] <= gd["stats"][0] * 1.05


def test_config2_oracle_matches_live_golden(golden, config2_oracle):
    """Pins the oracle itself at the published size (runs on the GPU box because it shares the 30 s oracle run)."""
    gd = golden("config2")
    oflow, oreg = config2_oracle
    mean, mx = epe_stats(oflow[::2, ::4, ::4], gd["flow_s"])
    assert mean <= 1e-6 and mx <= 1e-4, (mean, mx)
    assert rel_l2(oreg[::2, ::4, ::4], gd["reg_s"]) <= 1e-6


@pytest.mark.parametrize("state", ["f64", "f32"])
def test_config4_reduced_z_heavy_vs_oracle(cuda_backend, state):
    """Config 4 (single large 1-channel volume, high-displacement field) at a z-heavy reduced size the oracle finishes
    in about a minute: 64x256x256, min_level 2 (11 levels, finest 41x164x164), OFOptions defaults otherwise."""
    import flowreg3d_b200 as F
    from flowreg3d_b200 import core
    Z, Y, X = 64, 256, 256
    fixed = synth_volume((Z, Y, X), 4)
    g = smooth_flow((Z, Y, X), 5, 3.0, 10.0)
    f64 = fixed.astype(np.float64)
    moving = O.imregister_wrapper(f64, -g[..., 0], -g[..., 1], -g[..., 2], f64, "linear")
    kw = dict(alpha=(0.25,) * 3, update_lag=5, iterations=100, min_level=2, levels=100, eta=0.8, a_smooth=1.0,
              a_data=0.45)
    key = "_config4_oracle"
    if not hasattr(test_config4_reduced_z_heavy_vs_oracle, key):
        setattr(test_config4_reduced_z_heavy_vs_oracle, key, O.get_displacement(fixed, moving, **kw))
    ref = getattr(test_config4_reduced_z_heavy_vs_oracle, key)
    old = core.STATE_DTYPE
    core.STATE_DTYPE = np.float32 if state == "f32" else np.float64
    try:
        flow = F.get_displacement(fixed, moving, **kw)
    finally:
        core.STATE_DTYPE = old
    mean, mx = epe_stats(flow, ref)
    print(f"config4 reduced {state}: EPE vs oracle mean {mean:.2e} max {mx:.2e}")
    bm, bx = (1e-4, 5e-3) if state == "f64" else (1e-3, 5e-2)
    assert mean <= bm and mx <= bx, (state, mean, mx)
    f32 = flow.astype(np.float32)
    reg = F.imregister_wrapper(moving, f32[..., 0], f32[..., 1], f32[..., 2], fixed, "cubic")
    oreg = O.imregister_wrapper(moving, f32[..., 0], f32[..., 1], f32[..., 2], fixed, "cubic")
    assert rel_l2(reg, oreg) <= 1e-6
