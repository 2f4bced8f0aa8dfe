"""CPU: host-side logic of the ctypes shim (diff_gaussian_rasterization/_C.py) that needs no GPU -- the layout of the
flat gradient buffer and the deferred-R bookkeeping of the speculative instance buffer (the argument contract of
GaussianRasterizer is covered by test_abi_cpu.py::test_shim_argument_contract)."""
import pytest
import torch


def _C():
    import diff_gaussian_rasterization as dgr

    return dgr._C


@pytest.mark.parametrize("P,M,flags", [(1000, 16, (True, True, False, False)), (37, 16, (False, False, True, True)),
                                       (5, 4, (True, True, True, True)), (1, 16, (True, False, False, True))])
def test_gradient_views_layout(P, M, flags):
    """Eight gradients = disjoint, 16-byte aligned slices of one flat buffer, in the documented order; the size helper
    covers the largest case."""
    c = _C()
    has_sh, has_scales, has_colors, has_cov = flags
    widths = (3, 3 * M if has_sh else 0, 1, 3 if has_scales else 0, 4 if has_scales else 0, 3 if has_colors else 0,
              6 if has_cov else 0, 3)
    total = sum((P * w + 3) // 4 * 4 for w in widths)
    assert total <= c.gradient_buffer_floats(P, M)
    flat = torch.arange(total, dtype=torch.float32)
    m2, m3, op, col, cov, sh, sc, rot = c._grad_views(flat, P, M, has_sh, has_scales, has_colors, has_cov)
    assert m2.shape == (P, 3) and m3.shape == (P, 3) and op.shape == (P, 1)
    assert (sh is not None) == has_sh and (sc is not None) == has_scales == (rot is not None)
    assert (col is not None) == has_colors and (cov is not None) == has_cov
    seen = torch.zeros(total, dtype=torch.int32)
    for v in (m2, m3, op, col, cov, sh, sc, rot):
        if v is None:
            continue
        assert v._base is flat and v.is_contiguous()
        assert (v.storage_offset() * 4) % 16 == 0
        seen[v.storage_offset():v.storage_offset() + v.numel()] += 1
    assert int(seen.max()) == 1  # disjoint
    # order inside the buffer: means3D first, means2D last (view_parallel sums [0, total) in one go)
    assert m3.storage_offset() == 0 and m2.storage_offset() + m2.numel() <= total
    if has_sh:
        assert sh.shape == (P, M, 3) and float(sh.reshape(-1)[0]) == float(sh.storage_offset())


class _FakeEvent:
    def __init__(self):
        self.syncs = 0

    def synchronize(self):
        self.syncs += 1


def test_pending_r_validation_and_bookkeeping():
    """PendingR (GVD_SPECULATE=defer): resolves once to a Counts (R with V attached), returns its slot, grows the history,
    raises SpeculationOverflow (every time) when the frame outgrew the instance buffer OR the chunk histogram."""
    c = _C()
    st = {"max_R": 100, "max_V": 10, "free": [], "open": []}
    slot = {"event": _FakeEvent(), "host": [1234, 77]}
    p = c.PendingR(slot, c._capacity(100), 500, st)
    r = p.resolve()
    assert int(p) == 1234 and r == 1234 and r.visible == 77 and slot["event"].syncs == 1
    assert st["free"] == [slot] and st["max_R"] == 1234 and st["max_V"] == 77

    for host, cap, vcap in (([5000, 50], 4096, 500), ([100, 900], 4096, 500)):
        st = {"max_R": 10, "max_V": 10, "free": [], "open": []}
        slot = {"event": _FakeEvent(), "host": host}
        p = c.PendingR(slot, cap, vcap, st)
        for _ in range(2):
            with pytest.raises(c.SpeculationOverflow, match="speculative buffers"):
                p.resolve()
        assert st["max_R"] == max(10, host[0]) and st["max_V"] == max(10, host[1]) and st["free"] == [slot]  # the retry will fit
    assert c._capacity(5000) >= 2 * 5000 and c._visible_capacity(100, 150) == 150
    assert issubclass(c.SpeculationOverflow, RuntimeError)
    assert not c.DEFER, "exact sizing is the default; GVD_SPECULATE=defer is the opt-in"
