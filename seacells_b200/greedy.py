"""Greedy adaptive CSSP initialisation (cpu.py:342-423) - "next" row of the scope table.

Not part of the two hot paths.  The oracle restates it (oracle.greedy_centers, pinned by the golden vectors); the GPU
version is scheduled after the hot paths.  Until then this raises instead of silently running on the CPU.
"""


def greedy_centers(model, n_centers):
    raise NotImplementedError(
        "greedy archetype initialisation is not built for the B200 engine yet; pass `initial_archetypes=` to "
        "fit()/initialize() (e.g. indices from a previous run, random cells, or palantir waypoints).")
