"""Orbital-space bookkeeping (reference: ``pymolresponse/indices.py:4-37``)."""

from __future__ import annotations

from dataclasses import dataclass

# (nocc_alpha, nvirt_alpha, nocc_beta, nvirt_beta)
Occupations = tuple[int, int, int, int]


@dataclass(frozen=True)
class Indices:
    """ROHF-style (closed, active, secondary) index pairs, as full-MO indices."""

    indices_closed_act: list[tuple[int, int]]
    indices_closed_secondary: list[tuple[int, int]]
    indices_act_secondary: list[tuple[int, int]]


def form_indices_from_occupations(occupations: Occupations) -> Indices:
    if len(occupations) != 4:
        raise AssertionError("occupations must be (nocc_a, nvirt_a, nocc_b, nvirt_b)")
    na, va, nb, vb = occupations
    assert na + va == nb + vb
    norb = na + va
    n_active = abs(int(na - nb))
    n_closed = (na + nb - n_active) // 2
    closed = range(0, n_closed)
    active = range(n_closed, n_closed + n_active)
    secondary = range(n_closed + n_active, norb)
    return Indices(
        indices_closed_act=[(i, t) for i in closed for t in active],
        indices_closed_secondary=[(i, a) for i in closed for a in secondary],
        indices_act_secondary=[(t, a) for t in active for a in secondary],
    )
