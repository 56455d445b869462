"""CPU tests of the qtorch-compatible surface's host arithmetic: quantize / dequantize / VECT_C views follow the
reference's definitions (python/qtorch/nn/functional/quantization.py:27-152, utils.py:5-30)."""
import numpy as np
import torch


def _q():
    import lowbitdnn_project_b200  # noqa: F401  (registers the package under its importable name)
    import lowbitdnn_project_b200.qtorch as q
    return q


def test_quantize_is_symmetric_round_half_even_and_remembers_its_scale():
    q = _q()
    x = torch.tensor([[-2.0, -0.5, 0.0, 0.25, 1.0, 0.0117647059 * 42.5]]).reshape(1, 1, 2, 3)
    xq = q.quantize(x)
    assert xq.dtype == torch.int8 and xq in q.QUANTIZATION_PARAMETERS
    scale = q.QUANTIZATION_PARAMETERS[xq]
    assert abs(scale - 4.0 / 255.0) < 1e-12                     # (max - min) / (qmax - qmin) with the range forced symmetric
    inv = 255.0 / 4.0
    want = torch.tensor([v * inv for v in x.flatten().tolist()], dtype=torch.float32).clamp(-128, 127).round().to(torch.int8)
    assert torch.equal(xq.flatten(), want)
    assert int(xq.flatten()[0]) == -128                          # -127.5 clamps, then rounds half-to-even to -128
    assert q.quantize(xq) is xq                                  # already quantized: returned as it is (quantization.py:116)
    back = q.dequantize(xq)
    assert back.dtype == torch.float32 and torch.allclose(back, xq.float() * scale)
    assert (back - x).abs().max() <= scale / 2 + 1e-6


def test_vect_c_views_round_trip_and_match_the_reference_definition():
    q = _q()
    t = torch.arange(2 * 8 * 3 * 5, dtype=torch.int32).reshape(2, 8, 3, 5)
    v = q.to_vect_c(t)
    assert tuple(v.shape) == (2, 2, 3, 5, 4)
    assert int(v[1, 1, 2, 3, 2]) == int(t[1, 1 * 4 + 2, 2, 3])
    assert torch.equal(q.from_vect_c(v), t)
    try:
        q.to_vect_c(torch.zeros(1, 6, 2, 2))
    except NotImplementedError:
        pass
    else:
        raise AssertionError("C % 4 != 0 must be refused, as in the reference")


def test_stochastic_rounding_stays_within_one_step():
    q = _q()
    torch.manual_seed(0)
    x = torch.rand(1, 4, 8, 8) * 2 - 1
    a, b = q.quantize(x.clone()), q.quantize(x.clone(), stochastic=True)
    assert (a.int() - b.int()).abs().max() <= 1
    assert q.QUANTIZATION_PARAMETERS[a] == q.QUANTIZATION_PARAMETERS[b]
