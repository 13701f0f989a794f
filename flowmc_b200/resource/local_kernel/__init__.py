"""Alias package: ``flowMC.resource.local_kernel`` is how BASELINE.json's north_star (and flowMC before 0.4) spells
``flowMC.resource.kernel`` (src/flowMC/resource/kernel/ in the 0.4.5 checkout).  The submodules MALA, HMC,
Gaussian_random_walk, NF_proposal and base re-export the same classes under this spelling."""
