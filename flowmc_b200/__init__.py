"""flowmc_b200 -- B200-native implementation of flowMC's sampling hot path.

Same resource/strategy plugin API as kazewong/flowMC (v0.4.5); the work underneath is
hand-written sm_100a CUDA reached through the C ABI in ``include/flowmc_b200.h``.
Module layout mirrors the reference (``flowmc_b200.resource.kernel.MALA`` <->
``flowMC.resource.kernel.MALA`` and so on); arrays are CUDA ``torch.Tensor``s, PRNG keys are
host ``numpy.uint32[2]`` arrays with jax.random's threefry semantics (``flowmc_b200.random``).
"""
__version__ = "0.1.0"
