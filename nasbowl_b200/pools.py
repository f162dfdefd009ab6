"""Candidate pools on the device with the reference's de-duplication.

`mutation_pool` is the device form of the reference's `mutation(..., allow_isomorphism=False)`
(bayesopt/generate_test_graphs.py:393-539) for NB101-shaped cells: children of the best parents are
generated in batches (nbw_mutate_cells), a child isomorphic to a parent or to an earlier member of the pool
is dropped (nbw_pool_dedup: a WL isomorphism test instead of the reference's pairwise networkx search), and
the pool is topped up with random cells when the children run out (:527-533)."""
from __future__ import annotations

from typing import Optional, Tuple

import torch


def dedup_pool(eng, cells: torch.Tensor, n_max: int, labeled: bool = True,
               protected: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, int]:
    """Records with every cell isomorphic to an earlier one (or to a `protected` cell) removed; order kept.
    labeled=False reproduces the reference's structure-only test (is_isomorphic without node_match)."""
    n_prot = 0 if protected is None else protected.shape[0]
    both = cells if n_prot == 0 else torch.cat([protected, cells], dim=0).contiguous()
    keep, _, n_kept = eng.pool_dedup(both, n_max, labeled=labeled, n_protected=n_prot)
    return both[n_prot:][keep[n_prot:].bool()], n_kept


def mutation_pool(eng, parents: torch.Tensor, pool_size: int, seed: int = 0, n_mutate: Optional[int] = None,
                  allow_isomorphism: bool = False, labeled: bool = False, patience: int = 50,
                  family: int = 0) -> torch.Tensor:
    """BANANAS-style evaluation pool of `pool_size` NB101-shaped cells: up to n_mutate (default pool_size / 2)
    distinct mutated children of `parents` (device records), the rest random cells.  With
    allow_isomorphism=False (the reference's default) no child is isomorphic to a parent or to an earlier child."""
    assert family == 0, "cell-level mutation is implemented for NB101-shaped cells (use mutate_encoded for NB201 / DARTS)"
    if n_mutate is None:
        n_mutate = int(0.5 * pool_size)
    assert pool_size >= n_mutate, " pool_size must be larger or equal to n_mutate"
    kept = None
    first, rounds = 0, 0
    while (0 if kept is None else kept.shape[0]) < n_mutate and rounds < patience:
        want = n_mutate - (0 if kept is None else kept.shape[0])
        batch = max(64, 2 * want)
        kids = eng.mutate_cells(family, parents, seed, first, batch)
        first += batch
        rounds += 1
        if allow_isomorphism:
            new = kids
        else:
            prot = parents if kept is None else torch.cat([parents, kept], dim=0).contiguous()
            new, _ = dedup_pool(eng, kids, 8, labeled=labeled, protected=prot)
        kept = new if kept is None else torch.cat([kept, new], dim=0)
    kept = kept[:n_mutate] if kept is not None else parents[:0]
    n_random = max(pool_size - kept.shape[0], 0)
    if n_random:
        kept = torch.cat([kept, eng.generate_cells(family, seed ^ 0x5EED, 1_000_000, n_random)], dim=0)
    return kept.contiguous()
