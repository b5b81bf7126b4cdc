"""Parity of the CUDA path (through the C ABI) against the CPU oracle.  Needs a B200.

Tolerances (BASELINE.json north_star): loss and gradients rel 1e-4 in fp32; `valid` bit-exact;
`sel` equal outside near-ties (gradient parity is measured against the
oracle run with the kernel's own arg-min decision and (a, b))."""
import pytest
import torch

import coivo_b200
from coivo_b200.synthetic import make_triplets, make_sequence
from oracle import photometric as O
from conftest import record_parity

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TOL = 1e-4


def relinf(a, b):
    return (a.cpu() - b).abs().max().item() / max(b.abs().max().item(), 1e-30)


def assert_close_upto_kinks(got, ref, kinks, what, atol=0.0):
    """max-norm parity at TOL, except for at most 8 elements per L1-kink sample (oracle.l1_kink_count: the
    sample's own depth texel(s) / four bilinear taps), each bounded by 1e-2 of the max.  Strict when kinks == 0.
    Returns (elements beyond TOL, rel max error) for the record."""
    err = (got - ref).abs()
    scale = ref.abs().max().item()
    bad = err > TOL * scale + atol
    assert int(bad.sum()) <= 8 * kinks, f"{what}: {int(bad.sum())} elements beyond {TOL} (rel {relinf(got, ref)}), {kinks} L1 kinks"
    assert err.max().item() <= 1e-2 * scale + atol, f"{what}: outlier {relinf(got, ref)}"
    return int(bad.sum()), err.max().item() / max(scale, 1e-30)


def run_cuda(d, **kw):
    depth = [x.to(DEV).requires_grad_() for x in d["depth"]]
    pose = d["pose"].to(DEV).requires_grad_()
    srcs = d["srcs"].to(DEV).requires_grad_()
    out = coivo_b200.photometric_loss(depth, pose, d["K"].to(DEV), d["tgt"].to(DEV), srcs, return_masks=True, **kw)
    loss, valid, sel, ab = out
    loss.backward()
    torch.cuda.synchronize()
    return loss, valid.cpu(), sel.cpu(), ab.cpu(), [x.grad.cpu() for x in depth], pose.grad.cpu(), srcs.grad.cpu()


# SURVEY.md section 7.4 H2 asks for `sel` equal to the oracle's except where the two smallest candidates differ by < 1e-6.
# That band assumed fp32 evaluations of pe agree to ~1e-7; they do not on flat regions: sigma = E[x^2] - mu^2 carries
# ~6e-8 absolute rounding against C2 = 9e-4, i.e. up to ~1e-4 relative noise in the SSIM term.  So instead of widening
# the band on faith, EVERY mismatch is adjudicated against the candidates evaluated in float64: the candidate the kernel
# chose must lie within SEL_EXCESS of the true (fp64) minimum, and the fp32 oracle's own arg-min is held to the same
# yardstick and reported next to it (measured: kernel <= 2.6e-5 at 1080x1350, the fp32 oracle itself 2.3e-5 at config 2).
SEL_EXCESS = 5e-5


def adjudicate(d, sel, s0, ab, name, **kw):
    mism = sel != s0
    n_mism = int(mism.sum())
    akw = dict(alpha=kw.get("alpha", 0.85), lcc=kw.get("lcc", True))
    ex_k, sel64 = O.adjudicate_sel(d["depth"], d["pose"], d["K"], d["tgt"], d["srcs"], sel, ab, **akw)
    ex_o, _ = O.adjudicate_sel(d["depth"], d["pose"], d["K"], d["tgt"], d["srcs"], s0, ab, **akw)
    stats = dict(pixels=sel.numel(), sel_mism=n_mism, kernel_vs_fp64=int((sel != sel64).sum()),
                 oracle32_vs_fp64=int((s0 != sel64).sum()), kernel_excess_max=f"{ex_k.max().item():.2e}",
                 oracle32_excess_max=f"{ex_o.max().item():.2e}",
                 mism_beyond_1e6=int((ex_k[mism] > 1e-6).sum()) if n_mism else 0)
    assert ex_k.max().item() <= SEL_EXCESS, f"{name}: kernel chose a candidate {ex_k.max().item():.3e} above the fp64 minimum"
    assert mism.float().mean().item() < 1e-3
    return stats


def check_against_oracle(d, depth_atol=0.0, name=None, scatter=None, **kw):
    N, S = d["srcs"].shape[1], len(d["depth"])
    name = name or f"B{d['tgt'].shape[0]}_{d['tgt'].shape[2]}x{d['tgt'].shape[3]}_N{N}_S{S}" + "".join(f"_{k}={v}" for k, v in kw.items())
    loss, valid, sel, ab, gd, gT, gs = run_cuda(d, **(dict(kw, scatter=scatter) if scatter else kw))
    with torch.no_grad():
        l0, v0, s0, ab0 = O.photometric_loss(d["depth"], d["pose"], d["K"], d["tgt"], d["srcs"], return_masks=True, **kw)
    assert torch.equal(valid, v0), "valid mask must be bit-exact"
    assert torch.allclose(ab, ab0, rtol=1e-5, atol=1e-6)
    stats = adjudicate(d, sel, s0, ab, name, **kw)
    assert abs(loss.item() - l0.item()) <= TOL * abs(l0.item()), (loss.item(), l0.item())
    od = [x.clone().requires_grad_() for x in d["depth"]]
    op = d["pose"].clone().requires_grad_()
    osr = d["srcs"].clone().requires_grad_()
    l1 = O.photometric_loss(od, op, d["K"], d["tgt"], osr, sel_override=sel, ab_override=ab, **kw)
    l1.backward()
    kinks = O.l1_kink_count(d["depth"], d["pose"], d["K"], d["tgt"], d["srcs"], sel, ab) if kw.get("alpha", 0.85) < 1 else 0
    outliers, worst = 0, 0.0
    for k in range(S):
        n, e = assert_close_upto_kinks(gd[k], od[k].grad, kinks, f"grad_depth[{k}]", depth_atol)
        outliers += n
        worst = max(worst, e if depth_atol == 0.0 else 0.0)
    assert relinf(gT[:, :, :3], op.grad[:, :, :3]) < TOL, f"grad_pose {relinf(gT, op.grad)}"
    assert gT[:, :, 3].abs().max().item() == 0
    n, e = assert_close_upto_kinks(gs, osr.grad, kinks, "grad_srcs")
    record_parity(name, loss_rel=f"{abs(loss.item() - l0.item()) / abs(l0.item()):.1e}", **stats, l1_kinks=kinks,
                  grad_outliers=outliers + n, grad_rel_max=f"{max(worst, e, relinf(gT[:, :, :3], op.grad[:, :, :3])):.1e}")


@pytest.mark.parametrize("B,H,W,N,S", [
    (1, 256, 320, 2, 1),      # BASELINE config 1
    (2, 64, 96, 2, 4),
    (1, 37, 53, 2, 4),        # ragged: neither a tile multiple nor divisible by 2^k
    (3, 16, 24, 1, 2),        # single source
    (1, 8, 40, 2, 3),         # one tile row
])
def test_parity_small(B, H, W, N, S):
    check_against_oracle(make_triplets(B, H, W, N=N, S=S, seed=B + H))


@pytest.mark.parametrize("kw", [dict(lcc=False), dict(lcc_detach=True), dict(alpha=0.5, smooth_weight=0.1),
                                dict(smooth_weight=0.0)])
def test_parity_flags(kw):
    check_against_oracle(make_triplets(2, 48, 64, seed=21), **kw)


@pytest.mark.parametrize("B,H,W,N,S", [(2, 64, 96, 2, 4), (1, 37, 53, 1, 3), (2, 256, 320, 2, 4)])
def test_parity_warp_aggregated_scatter(B, H, W, N, S):
    """COLVO_F_SCATTER_MERGE (north_star: a scatter-add that avoids contended global atomics by warp aggregation): the same
    parity protocol with `scatter="merged"`; ragged sizes exercise warps whose active lanes are a strict prefix."""
    check_against_oracle(make_triplets(B, H, W, N=N, S=S, seed=71 + H), name=f"merged_scatter_B{B}_{H}x{W}_N{N}_S{S}", scatter="merged")


def test_parity_config2_one_triplet_full_size():
    # BASELINE config 2 shape (256x320, N=2, S=4) at a batch the oracle finishes in seconds
    check_against_oracle(make_triplets(2, 256, 320, seed=0))


def test_parity_config2_full_batch():
    # BASELINE config 2 as it is benchmarked: 12 triplets 256x320, N=2, S=4, every gradient against the oracle
    check_against_oracle(make_triplets(12, 256, 320, seed=1), name="config2_B12_256x320")


def test_parity_highres_slice():
    # BASELINE config 4 geometry (W = 1350 is not a multiple of 4/8/32; pyramid by floor), cropped in H
    check_against_oracle(make_triplets(1, 136, 1350, seed=4))


def test_parity_config4_full_resolution():
    # BASELINE config 4 at its real frame size (1080x1350, pyramid 540x675, 270x337, 135x168 by floor), one triplet
    check_against_oracle(make_triplets(1, 1080, 1350, seed=4), name="config4_B1_1080x1350")


def test_identity_pose_edge_case():
    # KAT-1 on the GPU: whole border rows/columns sit exactly on 0 and W-1 (zero coordinate gradient)
    d = make_triplets(1, 32, 48, seed=9)
    d["pose"] = torch.eye(4).reshape(1, 1, 4, 4).repeat(1, 2, 1, 1).contiguous()
    # With zero motion u' = rx*fx + cx does not depend on depth: the photometric depth gradient is an exact
    # cancellation (terms ~5e-3, fp32 residue ~5e-10 on either side) on top of a ~5e-7 smoothness gradient,
    # so the depth comparison carries an absolute floor here.
    check_against_oracle(d, depth_atol=2e-9)


def test_behind_camera_all_invalid():
    d = make_triplets(1, 32, 48, seed=10)
    d["pose"][:, 0, 2, 3] = -5.0           # source 0: every point behind the camera -> n = 0 -> (a, b) = (1, 0)
    loss, valid, sel, ab, gd, gT, gs = run_cuda(d)
    assert valid[:, 0].sum().item() == 0
    assert torch.equal(ab[:, 0], torch.tensor([1.0, 0.0]).expand_as(ab[:, 0]))
    check_against_oracle(d)


def test_no_grad_and_partial_grad_paths():
    d = make_triplets(2, 32, 48, seed=12)
    args = [[x.to(DEV) for x in d["depth"]], d["pose"].to(DEV), d["K"].to(DEV), d["tgt"].to(DEV), d["srcs"].to(DEV)]
    with torch.no_grad():
        l_ng = coivo_b200.photometric_loss(*args)
    depth = [x.clone().requires_grad_() for x in args[0]]
    l = coivo_b200.photometric_loss(depth, args[1], args[2], args[3], args[4])
    l.backward()                              # srcs and pose do not require grad: scatter skipped
    assert abs(l.item() - l_ng.item()) < 1e-7
    od = [x.clone().requires_grad_() for x in d["depth"]]
    O.photometric_loss(od, d["pose"], d["K"], d["tgt"], d["srcs"]).backward()
    assert relinf(depth[0].grad, od[0].grad) < 5e-3      # without sel/ab override: near-tie noise only
    # grad_loss scaling flows through
    depth2 = [x.clone().requires_grad_() for x in args[0]]
    (3.0 * coivo_b200.photometric_loss(depth2, args[1], args[2], args[3], args[4])).backward()
    assert relinf(depth2[1].grad, 3.0 * depth[1].grad.cpu()) < 1e-5


def test_determinism_of_forward_and_non_scatter_grads():
    d = make_triplets(2, 64, 96, seed=13)
    a = run_cuda(d)
    b = run_cuda(d)
    assert a[0].item() == b[0].item()
    assert torch.equal(a[2], b[2]) and torch.equal(a[3], b[3])
    for x, y in zip(a[4], b[4]):
        assert torch.equal(x, y)              # depth gradients use no atomics
    assert torch.equal(a[5], b[5])            # pose gradients: fixed-order reduction
    assert relinf(a[6], b[6]) < 1e-5          # grad_srcs: float REDG order is not fixed


def test_full_size_properties_config2():
    """BASELINE config 2 at full batch (12 x 256x320): size-independent properties instead of the oracle."""
    d = make_triplets(12, 256, 320, seed=1)
    loss, valid, sel, ab, gd, gT, gs = run_cuda(d)
    assert torch.isfinite(loss) and loss.item() > 0
    # batch-shard additivity: loss(B) == mean of the two half-batch losses
    halves = []
    for s in (slice(0, 6), slice(6, 12)):
        sub = {k: ([x[s].contiguous() for x in v] if isinstance(v, list) else v[s].contiguous()) for k, v in d.items()}
        halves.append(run_cuda(sub))
    assert abs(loss.item() - 0.5 * (halves[0][0].item() + halves[1][0].item())) < 2e-6
    assert torch.equal(valid[:6], halves[0][1]) and torch.equal(sel[6:], halves[1][2])
    # gradients of a batch mean: each half's gradient is twice the full-batch one on its samples
    assert relinf(2 * gd[0][:6], halves[0][4][0]) < 1e-5
    # sum of grad_srcs over a source frame equals the analytic dL/db-type invariant: finite and bounded
    assert torch.isfinite(gs).all() and torch.isfinite(gT).all()


def test_consistency_sweep_matches_oracle():
    s = make_sequence(9, 64, 80, seed=2)
    got = coivo_b200.consistency(s["depth"].to(DEV), s["pose"].to(DEV), s["K"].to(DEV), s["frames"].to(DEV)).cpu()
    ref = O.consistency(s["depth"], s["pose"], s["K"], s["frames"])
    assert torch.allclose(got[:, 3], ref[:, 3], atol=0, rtol=0), "valid fraction comes from the bit-exact mask"
    assert torch.allclose(got[:, 1:3], ref[:, 1:3], rtol=1e-5, atol=1e-6)
    assert torch.allclose(got[:, 0], ref[:, 0], rtol=1e-4, atol=1e-7)
    Kp = s["K"].reshape(1, 3, 3).repeat(8, 1, 1).contiguous()
    got2 = coivo_b200.consistency(s["depth"].to(DEV), s["pose"].to(DEV), Kp.to(DEV), s["frames"].to(DEV), lcc=False).cpu()
    ref2 = O.consistency(s["depth"], s["pose"], Kp, s["frames"], lcc=False)
    assert torch.allclose(got2, ref2, rtol=1e-4, atol=1e-6)


def test_consistency_sweep_config5_frame_size_across_a_pass_boundary():
    """BASELINE config 5 at its real frame size (256x320), 230 frames: the sweep runs in passes of at most 256 MB of
    cached warped frames = 195 pairs at this size, so pairs 194 / 195 straddle a pass boundary."""
    F = 230
    s = make_sequence(F, 256, 320, seed=7)
    got = coivo_b200.consistency(s["depth"].to(DEV), s["pose"].to(DEV), s["K"].to(DEV), s["frames"].to(DEV)).cpu()
    ref = O.consistency(s["depth"], s["pose"], s["K"], s["frames"])
    assert got.shape == (F - 1, 4)
    assert torch.equal(got[:, 3], ref[:, 3]), "valid fraction comes from the bit-exact mask"
    assert torch.allclose(got[:, 1:3], ref[:, 1:3], rtol=1e-5, atol=1e-6)
    assert torch.allclose(got[:, 0], ref[:, 0], rtol=1e-4, atol=1e-7)
    record_parity("config5_F230_256x320", pairs=F - 1, pe_rel_max=f"{((got[:, 0] - ref[:, 0]).abs() / ref[:, 0].abs()).max().item():.1e}",
                  ab_abs_max=f"{(got[:, 1:3] - ref[:, 1:3]).abs().max().item():.1e}")


@pytest.mark.parametrize("chunks", [1, 2])
def test_host_stepper_matches_device_path(chunks):
    d = make_triplets(2, 64, 96, seed=14)
    st = coivo_b200.HostStepper(2, 2, 4, 64, 96, device=DEV, chunks=chunks)
    pin = lambda t: t.pin_memory()
    st.step([pin(x) for x in d["depth"]], pin(d["pose"]), pin(d["K"]), pin(d["tgt"]), pin(d["srcs"]))
    h = st.finish()
    loss, valid, sel, ab, gd, gT, gs = run_cuda(d)
    assert abs(h.item() - loss.item()) < 1e-6
    if chunks == 1:
        for k in range(4):
            assert torch.equal(st.h_grad_depth[k], gd[k])
        assert torch.equal(st.h_grad_T, gT)
    else:                                    # per-chunk scaling rounds differently in the last bit
        for k in range(4):
            assert relinf(st.h_grad_depth[k], gd[k]) < 1e-5
        assert relinf(st.h_grad_T, gT) < 1e-5
    assert relinf(st.h_grad_srcs, gs) < 1e-5
    # the same step with the gradients left on the device (only the loss is read back)
    sd = coivo_b200.HostStepper(2, 2, 4, 64, 96, device=DEV, chunks=chunks, grads="device")
    sd.step(*[pin(x) if not isinstance(x, list) else [pin(y) for y in x] for x in (d["depth"], d["pose"], d["K"], d["tgt"], d["srcs"])])
    l3 = sd.finish()
    assert abs(l3.item() - loss.item()) <= 1e-6 * abs(loss.item())
    assert sd.d2h_bytes() == 4 * len(sd.spans)
    for (lo, hi), (gd_c, gT_c, gs_c) in zip(sd.spans, sd.device_grads()):
        for k in range(4):
            assert relinf(gd_c[k], gd[k][lo:hi]) < 1e-5
        assert relinf(gT_c, gT[lo:hi]) < 1e-5 and relinf(gs_c, gs[lo:hi]) < 1e-5
    assert st.d2h_bytes() == 4 * (chunks + sum(x.numel() for x in gd) + gT.numel() + gs.numel())


def test_host_stepper_uint8_frames():
    """COLVO_F_HOST_U8: uint8 frames widened on the device, x = u8 * (1/255) -- bit-identical to the fp32 step on the
    same widened values."""
    d = make_triplets(3, 37, 53, seed=17)             # ragged size: the scalar tail of the widening kernel is exercised
    to_u8 = lambda t: (t * 255.0).round().clamp_(0, 255).to(torch.uint8)
    tgt8, srcs8 = to_u8(d["tgt"]), to_u8(d["srcs"])
    c = torch.tensor(1.0 / 255.0, dtype=torch.float32)
    d["tgt"], d["srcs"] = tgt8.float() * c, srcs8.float() * c
    pin = lambda t: t.pin_memory()
    st = coivo_b200.HostStepper(3, 2, 4, 37, 53, device=DEV, chunks=1, images="u8")
    st.step([pin(x) for x in d["depth"]], pin(d["pose"]), pin(d["K"]), pin(tgt8), pin(srcs8))
    h = st.finish()
    loss, valid, sel, ab, gd, gT, gs = run_cuda(d)
    assert h.item() == loss.item()
    for k in range(4):
        assert torch.equal(st.h_grad_depth[k], gd[k])
    assert torch.equal(st.h_grad_T, gT)
    assert relinf(st.h_grad_srcs, gs) < 1e-5
    assert st.h2d_bytes([pin(x) for x in d["depth"]], d["pose"], d["K"], tgt8, srcs8) < 0.4 * st.h2d_bytes(d["depth"], d["pose"], d["K"], d["tgt"], d["srcs"])
    with pytest.raises(ValueError):
        st.step(d["depth"], d["pose"], d["K"], d["tgt"], d["srcs"])      # float frames into a uint8 stepper


def test_graphed_step_replays_the_eager_result():
    d = make_triplets(2, 64, 96, seed=15)
    mk = lambda: ([x.to(DEV).requires_grad_() for x in d["depth"]], d["pose"].to(DEV).requires_grad_(),
                  d["K"].to(DEV), d["tgt"].to(DEV), d["srcs"].to(DEV).requires_grad_())
    depth, pose, K, tgt, srcs = mk()
    g = coivo_b200.GraphedStep(depth, pose, K, tgt, srcs)
    l1 = g.replay().item()
    l2 = g.replay().item()
    torch.cuda.synchronize()
    loss, valid, sel, ab, gd, gT, gs = run_cuda(d)
    assert l1 == l2 == loss.item()
    assert torch.equal(depth[0].grad.cpu(), gd[0]) and torch.equal(pose.grad.cpu(), gT)
    assert relinf(srcs.grad.cpu(), gs) < 1e-5
    # new data through the static buffers
    d2 = make_triplets(2, 64, 96, seed=16)
    with torch.no_grad():
        tgt.copy_(d2["tgt"]); srcs.copy_(d2["srcs"]); pose.copy_(d2["pose"])
        for a, b in zip(depth, d2["depth"]):
            a.copy_(b)
    l3 = g.replay().item()
    torch.cuda.synchronize()
    assert abs(l3 - run_cuda(d2)[0].item()) < 1e-7


@pytest.mark.parametrize("B,H,W,N,S", [(2, 48, 64, 2, 4), (1, 37, 53, 1, 3)])
def test_geometric_consistency_term_parity(B, H, W, N, S):
    """the extra term, its gradients to depth / pose and to the source depth maps."""
    d = make_triplets(B, H, W, N=N, S=S, seed=41)
    g = torch.Generator().manual_seed(5)
    sd = (1.0 + 0.5 * torch.rand(B, N, 1, H, W, generator=g)).contiguous()
    depth = [x.to(DEV).requires_grad_() for x in d["depth"]]
    pose = d["pose"].to(DEV).requires_grad_()
    srcs = d["srcs"].to(DEV).requires_grad_()
    sdg = sd.to(DEV).requires_grad_()
    loss, valid, sel, ab, occ = coivo_b200.photometric_loss(depth, pose, d["K"].to(DEV), d["tgt"].to(DEV), srcs,
                                                            src_depth=sdg, geo_weight=0.5, return_occlusion=True)
    loss.backward()
    with torch.no_grad():
        l_plain = coivo_b200.photometric_loss([x.detach() for x in depth], pose.detach(), d["K"].to(DEV), d["tgt"].to(DEV),
                                              srcs.detach())
    assert loss.item() > l_plain.item() + 1e-3            # the term is really there
    od = [x.clone().requires_grad_() for x in d["depth"]]
    op = d["pose"].clone().requires_grad_()
    osr = d["srcs"].clone().requires_grad_()
    osd = sd.clone().requires_grad_()
    l_ref, v_ref, _, _, occ_ref = O.photometric_loss(od, op, d["K"], d["tgt"], osr, src_depth=osd, geo_weight=0.5,
                                                     sel_override=sel.cpu(), ab_override=ab.cpu(), return_occlusion=True)
    l_ref.backward()
    # the soft occlusion mask (SC-Depth's 1 - diff; 0 where invalid): a constant by-product of the term
    assert torch.equal(valid.cpu(), v_ref)
    assert occ.shape == (B, N, S, H, W) and not occ.requires_grad
    assert (occ.cpu() - occ_ref).abs().max().item() <= 1e-6
    assert torch.equal(occ.cpu() == 0, (v_ref == 0) | (occ_ref == 0))
    with pytest.raises(ValueError):
        coivo_b200.photometric_loss(depth, pose, d["K"].to(DEV), d["tgt"].to(DEV), srcs, return_occlusion=True)
    assert abs(loss.item() - l_ref.item()) <= TOL * abs(l_ref.item())
    for k in range(S):
        assert relinf(depth[k].grad, od[k].grad) < TOL, f"grad_depth[{k}]"
    assert relinf(pose.grad[:, :, :3], op.grad[:, :, :3]) < TOL
    assert relinf(srcs.grad, osr.grad) < TOL
    assert relinf(sdg.grad, osd.grad) < TOL


def test_geometric_term_switched_off_gives_no_source_depth_gradient():
    """with `geo_weight == 0` (the default, or the start of a weight ramp) the source depth maps take no part:
    their gradient is None -- not an uninitialised buffer -- and the loss equals the plain one."""
    d = make_triplets(2, 32, 48, seed=44)
    sd = (1.0 + 0.5 * torch.rand(2, 2, 1, 32, 48)).to(DEV).requires_grad_()
    depth = [x.to(DEV).requires_grad_() for x in d["depth"]]
    args = (d["pose"].to(DEV), d["K"].to(DEV), d["tgt"].to(DEV), d["srcs"].to(DEV))
    loss = coivo_b200.photometric_loss(depth, *args, src_depth=sd, geo_weight=0.0)
    loss.backward()
    assert sd.grad is None and depth[0].grad is not None and torch.isfinite(depth[0].grad).all()
    with torch.no_grad():
        assert loss.item() == coivo_b200.photometric_loss([x.detach() for x in depth], *args).item()


@pytest.mark.parametrize("B,H,W,N,S", [(2, 48, 64, 2, 4), (1, 37, 53, 1, 3)])
def test_packed_bf16_image_storage_parity(B, H, W, N, S):
    """images stored as RGBA bf16 (8 B/pixel), fp32 arithmetic.  The oracle runs on
    the same quantised values widened to fp32, so the tolerances are the fp32 ones; only the inputs differ from
    the fp32 storage case (|x - bf16(x)| <= 2^-9 per pixel value)."""
    d = make_triplets(B, H, W, N=N, S=S, seed=43)
    tgt_p, srcs_p = coivo_b200.pack_images(d["tgt"]), coivo_b200.pack_images(d["srcs"])
    tgt_q, srcs_q = coivo_b200.unpack_images(tgt_p), coivo_b200.unpack_images(srcs_p)
    assert (tgt_q - d["tgt"]).abs().max().item() <= 2.0 ** -8
    depth = [x.to(DEV).requires_grad_() for x in d["depth"]]
    pose = d["pose"].to(DEV).requires_grad_()
    loss, valid, sel, ab = coivo_b200.photometric_loss(depth, pose, d["K"].to(DEV), tgt_p.to(DEV), srcs_p.to(DEV),
                                                       return_masks=True)
    loss.backward()
    with torch.no_grad():
        l0, v0, s0, ab0 = O.photometric_loss(d["depth"], d["pose"], d["K"], tgt_q, srcs_q, return_masks=True)
        gap = O.candidate_gap(d["depth"], d["pose"], d["K"], tgt_q, srcs_q)
    assert torch.equal(valid.cpu(), v0)
    assert (gap[sel.cpu() != s0] < 1e-4).all()
    assert abs(loss.item() - l0.item()) <= TOL * abs(l0.item())
    assert torch.allclose(ab.cpu(), ab0, rtol=1e-5, atol=1e-6)
    od = [x.clone().requires_grad_() for x in d["depth"]]
    op = d["pose"].clone().requires_grad_()
    O.photometric_loss(od, op, d["K"], tgt_q, srcs_q, sel_override=sel.cpu(), ab_override=ab.cpu()).backward()
    for k in range(S):
        assert relinf(depth[k].grad, od[k].grad) < TOL
    assert relinf(pose.grad[:, :, :3], op.grad[:, :, :3]) < TOL
    # and the same values through the planar fp32 path give the same loss
    l_planar = coivo_b200.photometric_loss([x.detach() for x in depth], pose.detach(), d["K"].to(DEV), tgt_q.to(DEV), srcs_q.to(DEV))
    assert abs(l_planar.item() - loss.item()) <= 1e-6 * abs(loss.item())
    with pytest.raises(ValueError):
        coivo_b200.photometric_loss(depth, pose, d["K"].to(DEV), tgt_p.to(DEV), srcs_p.to(DEV).requires_grad_())


def test_concurrent_streams_are_reentrant():
    """The C ABI keeps no state and enqueues on the caller's stream only (programmatic dependent launches included):
    two different batches on two streams at once give the results of running them one after the other."""
    ds = [make_triplets(2, 64, 96, seed=31), make_triplets(3, 48, 64, seed=32)]
    ref = [run_cuda(d) for d in ds]
    streams = [torch.cuda.Stream(DEV), torch.cuda.Stream(DEV)]
    outs = [None, None]
    for rep in range(3):
        ins = []
        for d in ds:
            ins.append(([x.to(DEV).requires_grad_() for x in d["depth"]], d["pose"].to(DEV).requires_grad_(), d["K"].to(DEV),
                        d["tgt"].to(DEV), d["srcs"].to(DEV).requires_grad_()))
        torch.cuda.synchronize()
        for i, st in enumerate(streams):
            with torch.cuda.stream(st):
                depth, pose, K, tgt, srcs = ins[i]
                loss = coivo_b200.photometric_loss(depth, pose, K, tgt, srcs)
                loss.backward()
                outs[i] = (loss, depth, pose, srcs)
        torch.cuda.synchronize()
        for i in range(2):
            loss, depth, pose, srcs = outs[i]
            assert loss.item() == ref[i][0].item()
            for k in range(4):
                assert torch.equal(depth[k].grad.cpu(), ref[i][4][k])       # deterministic outputs: bit-equal
            assert torch.equal(pose.grad.cpu(), ref[i][5])
            assert relinf(srcs.grad, ref[i][6]) < 1e-5                      # float atomics


def test_backward_as_first_cuda_call_of_a_fresh_thread():
    """The backward builds a TMA descriptor through a driver entry point, which needs the primary context current on the
    calling thread.  A fresh host thread whose FIRST CUDA call is that backward (no source gradient: no zero-fill launch
    in front of it) must work -- this is what autograd's own thread looks like in a new process."""
    import ctypes
    import threading
    from coivo_b200 import _lib
    d = make_triplets(2, 24, 40, seed=51)
    depth = [x.to(DEV).requires_grad_() for x in d["depth"]]
    pose = d["pose"].to(DEV).requires_grad_()
    tp, sp = coivo_b200.pack_images(d["tgt"]).to(DEV), coivo_b200.pack_images(d["srcs"]).to(DEV)
    loss = coivo_b200.photometric_loss(depth, pose, d["K"].to(DEV), tp, sp)
    out = {}

    def worker():
        try:
            loss.backward()
            torch.cuda.synchronize()
            out["ok"] = True
        except Exception as e:          # noqa: BLE001
            out["err"] = str(e)
    t = threading.Thread(target=worker)
    t.start(); t.join()
    assert out.get("ok"), out.get("err")
    assert torch.isfinite(depth[0].grad).all() and torch.isfinite(pose.grad).all()
