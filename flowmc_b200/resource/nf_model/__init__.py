"""Alias package: ``flowMC.resource.nf_model`` is how BASELINE.json's north_star spells
``flowMC.resource.model.nf_model`` (src/flowMC/resource/model/nf_model/ in the 0.4.5 checkout).  The submodules
rqSpline, realNVP and base re-export the same classes under this spelling."""
