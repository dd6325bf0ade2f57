"""DALTON integral label -> :class:`Operator` metadata (reference
``pymolresponse/interfaces/dalton/utils.py:4-111``).  No arithmetic: needed because the reference's
own disk-fixture tests (tests/test_calculators.py:41-42) describe operators by DALTON label."""

from __future__ import annotations

from pymolresponse_b200.operators import Operator

_AXIS = {"x": 0, "y": 1, "z": 2}
_PAIR = {"xx": 0, "xy": 1, "xz": 2, "yy": 3, "yz": 4, "zz": 5, "yx": 1, "zx": 2, "zy": 4}


def clean_dalton_label(original_label: str) -> str:
    """'PSO 002' -> 'pso_002'."""
    return original_label.lower().replace(" ", "_")


def dalton_label_to_operator(label: str) -> Operator:
    label = clean_dalton_label(label)
    spin_dependent = False
    if "diplen" in label:
        name, idx, imaginary = "dipole", _AXIS[label[0]], False
    elif "dipvel" in label:
        name, idx, imaginary = "dipvel", _AXIS[label[0]], True
    elif "angmom" in label:
        name, idx, imaginary = "angmom", _AXIS[label[0]], True
    elif "spnorb" in label:
        name, idx, imaginary, spin_dependent = "spinorb", _AXIS[label[0]], True, True
        nelec = label[1]
        if nelec in ("1", "2"):
            name += nelec
        elif nelec in (" ", "_"):
            name += "c"
    elif "fc" in label:
        name, idx, imaginary, spin_dependent = "fermi", int(label[6:8]) - 1, False, True
    elif "sd" in label:
        coord_atom = int(label[3:6])
        atom, first = divmod(coord_atom - 1, 3)
        pair = "xyz"[first] + label[7]
        name, idx, imaginary, spin_dependent = "sd", 6 * atom + _PAIR[pair], False, True
    elif "pso" in label:
        # the reference cannot map PSO labels to a slice either (interfaces/dalton/utils.py:86-93)
        raise AssertionError("PSO labels carry no slice index")
    else:
        raise RuntimeError(f"Unhandled DALTON operator label: {label}")
    return Operator(label=name, is_imaginary=imaginary, is_spin_dependent=spin_dependent,
                    slice_idx=idx, ao_integrals=None)
