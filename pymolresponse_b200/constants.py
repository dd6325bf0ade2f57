"""Physical constants used by the excitation-energy reports and ECD (reference
``pymolresponse/constants.py``; values from ``scipy.constants``, CODATA)."""

import scipy.constants as spc

alpha = spc.alpha

HARTREE_TO_EV = spc.physical_constants["Hartree energy in eV"][0]
HARTREE_TO_INVCM = spc.physical_constants["hartree-inverse meter relationship"][0] * (1 / 100)

# dipole moment: atomic units -> Debye
convfac_au_to_debye = 2.541746230211

# ECD: ESUECD = e * a0[Angstrom] * c * 1e36 * e * hbar / m_e   (constants.py:25-31)
echarge = spc.elementary_charge
xtang = spc.physical_constants["atomic unit of length"][0] * 1.0e10
ccm = spc.c
hbar = spc.hbar
emass = spc.electron_mass
esuecd = echarge * xtang * ccm * 1.0e36 * echarge * hbar / emass
