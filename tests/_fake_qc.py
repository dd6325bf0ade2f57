"""Duck-typed stand-ins for the two quantum-chemistry packages the reference gets its AO integrals
from (pyscf.gto.Mole, psi4.core.*).  Neither package is installed in this image; the adapters of
pymolresponse_b200.interfaces only call a handful of methods on them, which these fakes provide from
seeded arrays, so the adapter plumbing (labels -> intor names, component counts, spin-orbit sum over
nuclei, gauge origins, ao_eri hand-off) is exercised for real."""

import sys
import types

import numpy as np


class FakeMole:
    """The subset of pyscf.gto.Mole the adapters touch."""

    def __init__(self, n, nelec, eri=None, seed=3):
        r = np.random.default_rng(seed)
        self.n = n
        self.nelec = nelec
        self.natm = 3
        self._coords = r.standard_normal((3, 3))
        self._charges = np.array([8, 1, 1])
        self._masses = np.array([15.9949, 1.00783, 1.00783])
        self._eri = eri
        self._rinv_orig = None
        self.calls = []
        sym = lambda a: 0.5 * (a + a.transpose(0, 2, 1))      # noqa: E731
        asym = lambda a: 0.5 * (a - a.transpose(0, 2, 1))     # noqa: E731
        self._ints = {
            "cint1e_r_sph": sym(r.standard_normal((3, n, n))),
            "cint1e_cg_irxp_sph": asym(r.standard_normal((3, n, n))),
            "cint1e_giao_irjxp_sph": asym(r.standard_normal((3, n, n))),
            "cint1e_ipovlp_sph": asym(r.standard_normal((3, n, n))),
            "int1e_ovlp": np.eye(n),
        }
        self._prinvxp = asym(r.standard_normal((3, 3, n, n)).reshape(9, n, n)).reshape(3, 3, n, n)

    def intor(self, name, comp=None, aosym=None):
        self.calls.append((name, comp, aosym))
        if name == "int2e":
            assert aosym == "s1"
            return self._eri
        if name == "cint1e_prinvxp_sph":
            atom = int(np.argmin(np.abs(self._coords - self._rinv_orig).sum(axis=1)))
            return self._prinvxp[atom]
        return self._ints[name]

    def set_rinv_orig(self, xyz):
        self._rinv_orig = np.asarray(xyz)

    def set_common_orig(self, xyz):
        self.common_orig = np.asarray(xyz)

    def atom_coord(self, i):
        return self._coords[i]

    def atom_coords(self):
        return self._coords

    def atom_charge(self, i):
        return int(self._charges[i])

    def atom_charges(self):
        return self._charges

    def atom_mass_list(self, isotope_avg=False):
        return self._masses


def install_fake_psi4(monkeypatch, n, eri, seed=4):
    """Put a minimal ``psi4`` module into sys.modules (core.Wavefunction.build, core.MintsHelper with
    ao_eri / ao_dipole / ao_angular_momentum, core.get_global_option)."""
    r = np.random.default_rng(seed)
    dip = [0.5 * (a + a.T) for a in r.standard_normal((3, n, n))]
    ang = [0.5 * (a - a.T) for a in r.standard_normal((3, n, n))]

    class Mints:
        def __init__(self, basis):
            self.basis = basis

        def ao_eri(self):
            return eri

        def ao_dipole(self):
            return dip

        def ao_angular_momentum(self):
            return ang

    class Wfn:
        @staticmethod
        def build(molecule, basis_name):
            w = Wfn()
            w.molecule, w.basis_name = molecule, basis_name
            return w

        def basisset(self):
            return ("basis", self.basis_name)

    core = types.SimpleNamespace(Wavefunction=Wfn, MintsHelper=Mints,
                                 get_global_option=lambda key: "sto-3g")
    mod = types.ModuleType("psi4")
    mod.core = core
    monkeypatch.setitem(sys.modules, "psi4", mod)
    return np.stack(dip), np.stack(ang)
