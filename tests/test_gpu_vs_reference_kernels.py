"""This package against the REFERENCE'S OWN GPU KERNELS run on the same B200 (cubins of the reference's
CUDA-C templates, precompiled by tools/make_ref_gpu_baseline.py into the git-ignored baseline/_ref/).
Skipped when the cubins are not present."""
import numpy as np
import pytest

import ref_gpu_kernels as ref
from conftest import rel_max, rel_rms

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not ref.available(), reason="baseline/_ref cubins not built")]
KERNELS = ["newtonian", "plummer", "dehnen_k1", "dehnen_k2", "spline"]


@pytest.fixture(scope="module")
def nb():
    import nbody_streams_b200 as nb

    return nb


@pytest.fixture(scope="module")
def system(nb):
    import torch

    xd, md = nb.make_plummer_sphere(6000, M_total=1e9, a=1.0, seed=5)
    xs, ms = nb.make_plummer_sphere(2123, M_total=1e7, a=0.3, seed=6)
    pos = np.vstack([xd[:, :3], xs[:, :3]])
    mass = np.concatenate([md, ms])
    h = np.concatenate([np.full(6000, 0.05), np.linspace(0.01, 0.03, 2123)])
    return pos, mass, h


@pytest.mark.parametrize("kernel", KERNELS)
@pytest.mark.parametrize("precision", ["float32", "float32_kahan", "float64"])
def test_forces_and_potential_match_reference_gpu_kernels(nb, system, kernel, precision):
    import torch

    pos, mass, h = system
    dt = torch.float64 if precision == "float64" else torch.float32
    pos_t, mass_t, h_t = (torch.from_numpy(a).to(dt).cuda() for a in (pos, mass, h))
    kid = nb.KERNEL_MAP[kernel]
    r_acc = ref.launch("force", precision, pos_t, mass_t, h_t, nb.G_DEFAULT, kid).cpu().numpy()
    r_phi = ref.launch("potential", precision, pos_t, mass_t, h_t, nb.G_DEFAULT, kid).cpu().numpy()
    # same eps2 convention as the reference GPU path (1e-15 in every precision, fields.py:825,983)
    acc, phi = nb.compute_nbody_forces_potential_gpu(pos_t, mass_t, h_t, precision=precision, kernel=kernel,
                                                     eps2=1e-15)
    if precision == "float64":
        assert rel_max(acc, r_acc) <= 1e-12 and rel_max(phi, r_phi) <= 1e-12
    else:
        assert rel_rms(acc, r_acc) <= 1e-5 and rel_rms(phi, r_phi) <= 1e-5
