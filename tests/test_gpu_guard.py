"""Guard bands around every buffer handed to the C ABI: the library must write only inside the sizes it
reports (nb_packed_bytes, nb_workspace_bytes) and the N rows of the outputs.  Stands in for a memcheck pass."""
import ctypes as C

import numpy as np
import pytest

from nbody_streams_b200 import _lib

pytestmark = pytest.mark.gpu

GUARD = 4096          # bytes on each side
PATTERN = 0xA5


class Guarded:
    """uint8 device buffer with a patterned margin on both sides of the payload."""

    def __init__(self, nbytes, device):
        import torch

        self.n = int(nbytes)
        self.buf = torch.full((self.n + 2 * GUARD,), PATTERN, dtype=torch.uint8, device=device)

    @property
    def ptr(self):
        return self.buf.data_ptr() + GUARD

    def intact(self):
        return bool((self.buf[:GUARD] == PATTERN).all()) and bool((self.buf[GUARD + self.n:] == PATTERN).all())

    def payload(self, dtype, shape):
        return self.buf[GUARD:GUARD + self.n].view(dtype).reshape(shape)


@pytest.mark.parametrize("prec", [_lib.PREC_F32, _lib.PREC_F32_COMP, _lib.PREC_F64])
@pytest.mark.parametrize("kernel_id", [1, 4])
@pytest.mark.parametrize("n", [1, 257, 5000, 70001])
def test_nb_direct_stays_inside_its_buffers(prec, kernel_id, n):
    import torch

    lib = _lib.load()
    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(n)
    es = 8 if prec == _lib.PREC_F64 else 4
    tdt = torch.float64 if es == 8 else torch.float32
    pos = torch.from_numpy(rng.standard_normal((n, 3))).to(dev)
    mass = torch.from_numpy(rng.uniform(0.5, 2.0, n)).to(dev)
    h = torch.from_numpy(rng.uniform(0.01, 0.1, n)).to(dev)
    grid = C.c_int(0)
    _lib.check(lib.nb_sweep_grid(prec, kernel_id, _lib.MODE_BOTH, C.byref(grid)))
    ws_bytes = lib.nb_workspace_bytes(prec, n, grid.value, 1)
    packed = Guarded(lib.nb_packed_bytes(prec, n), dev)
    ws = Guarded(ws_bytes, dev)
    acc = Guarded(es * 3 * n, dev)
    phi = Guarded(es * n, dev)
    stream = torch.cuda.current_stream(dev).cuda_stream
    _lib.check(lib.nb_direct(prec, kernel_id, _lib.MODE_BOTH, pos.data_ptr(), _lib.DT_F64, mass.data_ptr(), _lib.DT_F64,
                             0.0, h.data_ptr(), _lib.DT_F64, 0.0, n, 1.0, 1e-15, packed.ptr, ws.ptr, ws_bytes,
                             grid.value, acc.ptr, phi.ptr, stream))
    torch.cuda.synchronize()
    for name, g in (("packed", packed), ("workspace", ws), ("acc", acc), ("phi", phi)):
        assert g.intact(), f"write outside the {name} buffer"
    a = acc.payload(tdt, (n, 3))
    p = phi.payload(tdt, (n,))
    assert torch.isfinite(a).all() and torch.isfinite(p).all()
    if n > 1:
        assert (p < 0).all()


@pytest.mark.parametrize("prec", [_lib.PREC_F32_COMP, _lib.PREC_F64])
def test_kick_drift_pack_and_energy_stay_inside_their_buffers(prec):
    import torch

    lib = _lib.load()
    dev = torch.device("cuda", 0)
    n = 3001
    rng = np.random.default_rng(1)
    es = 8 if prec == _lib.PREC_F64 else 4
    adt = _lib.DT_F64 if es == 8 else _lib.DT_F32
    tdt = torch.float64 if es == 8 else torch.float32
    pos, vel = Guarded(8 * 3 * n, dev), Guarded(8 * 3 * n, dev)
    pos.payload(torch.float64, (n, 3)).copy_(torch.from_numpy(rng.standard_normal((n, 3))))
    vel.payload(torch.float64, (n, 3)).copy_(torch.from_numpy(rng.standard_normal((n, 3))))
    acc = torch.from_numpy(rng.standard_normal((n, 3))).to(dev).to(tdt)
    packed = Guarded(lib.nb_packed_bytes(prec, n), dev)
    stream = torch.cuda.current_stream(dev).cuda_stream
    # a full pack first (masses, softening, metadata), then the fused kick+drift that rewrites x,y,z in place
    mass = torch.full((n,), 2.0, dtype=torch.float64, device=dev)
    _lib.check(lib.nb_pack(prec, 1, packed.ptr, 0, pos.ptr, _lib.DT_F64, mass.data_ptr(), _lib.DT_F64, 0.0, None, 0, 0.05,
                           n, lib.nb_padded_count(n), 1e-15, stream))
    _lib.check(lib.nb_kick_drift_pack(prec, pos.ptr, vel.ptr, acc.data_ptr(), adt, None, packed.ptr, 0, n, 1e-3, 1e-3,
                                      stream))
    _lib.check(lib.nb_kick(vel.ptr, acc.data_ptr(), adt, None, n, 1e-3, stream))
    part = Guarded(lib.nb_energy_workspace_bytes(n), dev)
    out = Guarded(8 * 8, dev)
    phi = torch.zeros(n, dtype=tdt, device=dev)
    _lib.check(lib.nb_energy(vel.ptr, mass.data_ptr(), _lib.DT_F64, phi.data_ptr(), adt, None, 0, n, part.ptr, out.ptr,
                             stream))
    torch.cuda.synchronize()
    for name, g in (("pos", pos), ("vel", vel), ("packed", packed), ("energy partials", part), ("energy out", out)):
        assert g.intact(), f"write outside the {name} buffer"


@pytest.mark.parametrize("prec", [_lib.PREC_F32, _lib.PREC_F32_COMP, _lib.PREC_F64])
@pytest.mark.parametrize("kernel_id", [1, 4])
@pytest.mark.parametrize("n", [1, 257, 5000, 70001])
def test_nb_direct_sym_stays_inside_its_buffers(prec, kernel_id, n):
    """Same for the pair-symmetric path: its atomics go to 4 doubles per padded slot of the workspace and nowhere
    else (per-particle masses / softening: general tiles; the last block is partial and padded)."""
    import torch

    lib = _lib.load()
    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(n)
    es = 8 if prec == _lib.PREC_F64 else 4
    tdt = torch.float64 if es == 8 else torch.float32
    pos = torch.from_numpy(rng.standard_normal((n, 3))).to(dev)
    mass = torch.from_numpy(rng.uniform(0.5, 2.0, n)).to(dev)
    h = torch.from_numpy(rng.uniform(0.01, 0.1, n)).to(dev)
    grid = C.c_int(0)
    _lib.check(lib.nb_sym_grid(prec, kernel_id, _lib.MODE_BOTH, C.byref(grid)))
    ws_bytes = lib.nb_sym_workspace_bytes(lib.nb_padded_count(n))
    assert ws_bytes == 32 * lib.nb_padded_count(n)
    packed = Guarded(lib.nb_packed_bytes(prec, n), dev)
    ws = Guarded(ws_bytes, dev)
    acc = Guarded(es * 3 * n, dev)
    phi = Guarded(es * n, dev)
    stream = torch.cuda.current_stream(dev).cuda_stream
    _lib.check(lib.nb_direct_sym(prec, kernel_id, _lib.MODE_BOTH, pos.data_ptr(), _lib.DT_F64, mass.data_ptr(),
                                 _lib.DT_F64, 0.0, h.data_ptr(), _lib.DT_F64, 0.0, n, 1.0, 1e-15, packed.ptr, ws.ptr,
                                 ws_bytes, grid.value, acc.ptr, phi.ptr, stream))
    torch.cuda.synchronize()
    for name, g in (("packed", packed), ("workspace", ws), ("acc", acc), ("phi", phi)):
        assert g.intact(), f"write outside the {name} buffer"
    a = acc.payload(tdt, (n, 3))
    p = phi.payload(tdt, (n,))
    assert torch.isfinite(a).all() and torch.isfinite(p).all()
    if n > 1:
        assert (p < 0).all()
    # a too small workspace is refused, not overrun
    rc = lib.nb_direct_sym(prec, kernel_id, _lib.MODE_BOTH, pos.data_ptr(), _lib.DT_F64, mass.data_ptr(), _lib.DT_F64, 0.0,
                           h.data_ptr(), _lib.DT_F64, 0.0, n, 1.0, 1e-15, packed.ptr, ws.ptr, ws_bytes - 32, grid.value,
                           acc.ptr, phi.ptr, stream)
    assert rc == -2      # NB_E_WORKSPACE
