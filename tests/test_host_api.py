"""Host-side operator: argument validation mirrors the oracle's, and the product path refuses
CPU tensors loudly (there is no CPU fallback)."""
import pytest
import torch

import coivo_b200
from coivo_b200.synthetic import make_triplets, make_sequence, pyramid_shapes


def test_cpu_tensors_are_rejected():
    d = make_triplets(1, 16, 24, seed=0)
    with pytest.raises(ValueError, match="CUDA-only"):
        coivo_b200.photometric_loss(d["depth"], d["pose"], d["K"], d["tgt"], d["srcs"])
    s = make_sequence(3, 16, 24)
    with pytest.raises(ValueError, match="CUDA-only"):
        coivo_b200.consistency(s["depth"], s["pose"], s["K"], s["frames"])


@pytest.mark.parametrize("mutate,exc", [
    (lambda d: d.update(depth=d["depth"][::-1]), ValueError),
    (lambda d: d.update(pose=d["pose"][:, :1]), ValueError),
    (lambda d: d.update(K=d["K"][:, :2]), ValueError),
    (lambda d: d.update(srcs=d["srcs"][:, :, :2]), ValueError),
    (lambda d: d.update(tgt=d["tgt"].double()), TypeError),
    (lambda d: d.update(depth=[]), ValueError),
])
def test_shape_and_dtype_validation_happens_before_any_launch(mutate, exc):
    d = make_triplets(1, 16, 24, seed=0)
    mutate(d)
    with pytest.raises(exc):
        coivo_b200.photometric_loss(d["depth"], d["pose"], d["K"], d["tgt"], d["srcs"])


def test_synthetic_generator_is_deterministic_and_shaped():
    a, b = make_triplets(2, 32, 48, seed=3), make_triplets(2, 32, 48, seed=3)
    assert torch.equal(a["tgt"], b["tgt"]) and torch.equal(a["pose"], b["pose"])
    assert [tuple(x.shape[-2:]) for x in a["depth"]] == pyramid_shapes(32, 48, 4)
    assert a["srcs"].shape == (2, 2, 3, 32, 48) and a["tgt"].min() >= 0 and a["tgt"].max() <= 1
    R = a["pose"][..., :3, :3]
    assert torch.allclose(R @ R.transpose(-1, -2), torch.eye(3).expand_as(R), atol=1e-6)
    assert not torch.equal(a["tgt"], make_triplets(2, 32, 48, seed=4)["tgt"])


def test_bench_algorithmic_bytes_match_survey():
    import bench
    ab = bench.alg_bytes_per_triplet(256, 320)
    assert ab["step"] == 9_169_920 and ab["fwd"] == 3_384_320
    assert bench.alg_bytes_per_triplet(1080, 1350)["step"] == 163_202_040
    assert bench.alg_bytes_per_pair(256, 320) == 2_293_760            # config 5
    assert set(bench.CONFIGS) == {2, 4, 5} and bench.CONFIGS[2]["B"] == 12 and bench.CONFIGS[4]["W"] == 1350


def test_source_hash_is_stable_and_sees_the_kernel_sources():
    from coivo_b200 import _lib
    h = _lib.source_hash()
    assert len(h) == 16 and h == _lib.source_hash()
