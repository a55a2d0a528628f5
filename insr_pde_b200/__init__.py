"""insr_pde_b200 -- B200-native (sm_100a) implementation of the INSR-PDE per-timestep
optimisation hot path: fused SIREN field evaluation, spatial derivatives and back-propagation
behind the reference's own ``MLP`` / ``diff_ops`` call signatures.

    from insr_pde_b200 import MLP, get_network, gradient, divergence, laplace, jacobian

``insr_pde_b200.patch.install()`` rebinds those names inside the reference's ``base`` package
so that its ``main.py`` runs unchanged on the fused kernels.
"""
from .networks import MLP, Sine, get_network, sine_init, first_layer_sine_init  # noqa: F401
from .diff_ops import gradient, divergence, laplace, jacobian, hessian  # noqa: F401
from .sampling import (sample_uniform, sample_random, sample_boundary,  # noqa: F401
                       sample_boundary2D_separate, shard)

__version__ = "0.1.0"
