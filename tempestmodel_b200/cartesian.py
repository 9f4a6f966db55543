"""Cartesian GLL grids (reference src/atm/GridCartesianGLL.{h,cpp}): node
identification for the averaging groups of the DSS.

The reference fills one-node halos through its exchange connectivity - periodic
boundaries wrap the global element indices (GridCartesianGLL.cpp:380-432) - and
then averages across every element edge (ApplyDSS, :508-654).  Here a node is
named by its periodic global index, so that duplicates across element edges,
patch edges and periodic boundaries land in one averaging group.  With a single
element across a periodic direction (XZ slices: one element in y) the two
edges of the same element are duplicates of each other, as in the reference.
"""
import numpy as np


def node_ids(nelem_a, nelem_b, elem_a0, elem_b0, ne_a, ne_b, np_,
             periodic_a=True, periodic_b=True):
    """Global ids [nelem_a*np][nelem_b*np] of a patch whose first element is
    (elem_a0, elem_b0) on a domain of ne_a x ne_b elements."""
    if not (periodic_a and periodic_b):
        raise NotImplementedError(
            "non-periodic Cartesian boundaries (GridPatchCartesianGLL::"
            "ApplyBoundaryConditions) are not implemented")
    ea = elem_a0 + np.arange(nelem_a * np_) // np_
    ia = np.arange(nelem_a * np_) % np_
    eb = elem_b0 + np.arange(nelem_b * np_) // np_
    jb = np.arange(nelem_b * np_) % np_
    na = ne_a * (np_ - 1)
    nb = ne_b * (np_ - 1)
    ua = (ea * (np_ - 1) + ia) % na
    ub = (eb * (np_ - 1) + jb) % nb
    A, B = np.meshgrid(ua, ub, indexing="ij")
    return (A.astype(np.int64) * nb + B).astype(np.int64)
