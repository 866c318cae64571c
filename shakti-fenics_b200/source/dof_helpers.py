"""Map dofs saved in parallel ordering to a serial ordering by coordinate matching
(post-processing helper of the reference, source/dof_helpers.py:5-13; same signature).

The B200 path saves fields in the caller's (serial) vertex numbering already, so for its
output this returns the identity; it is kept for results written by the reference."""
import numpy as np


def dofs_to_serial(nodes_parallel, nodes_serial):
    tol = 1e-2
    close = np.all(np.abs(nodes_parallel - nodes_serial) < 1, axis=1)
    map_dofs = np.arange(nodes_parallel.shape[0])
    for j in np.where(~close)[0]:
        hit = np.where((np.abs(nodes_parallel[:, 0] - nodes_serial[j, 0]) < tol)
                       & (np.abs(nodes_parallel[:, 1] - nodes_serial[j, 1]) < tol))[0]
        map_dofs[j] = hit[0]
    return map_dofs
