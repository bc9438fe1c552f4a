"""unitaria_b200 -- a B200-native statevector engine behind unitaria's simulate seam.

Hot path replaced: ``Circuit.simulate`` -> ``tq.simulate(..., backend="qulacs")``
(/root/reference/src/unitaria/circuit.py:42-70) plus ``Subspace.project`` / normalization / norm
(/root/reference/src/unitaria/nodes/node.py:363-403, subspace.py:297-335).

Public surface:

* :func:`install` / :func:`uninstall` -- patch unitaria in place (drop-in backend).
* :func:`simulate_circuit`, :func:`simulate_projected`, :func:`simulate_norm`,
  :func:`estimate_norm_sampled` (sampling mode), :func:`block_encoding_compute` (device-resident
  ``BlockEncoding.compute``) -- the same calls on any tequila-like circuit object.
* :class:`StateVector` -- one device-resident register; :class:`SubspaceProgram`; :func:`lower_gates`.

There is no CPU fallback: the CUDA library must be built (``python -m unitaria_b200.build``) and a
CUDA device must be present for any compute call.
"""

from .adapter import (  # noqa: F401
    block_encoding_compute,
    clear_caches,
    configure,
    estimate_norm_sampled,
    install,
    seed_sampling,
    simulate_circuit,
    simulate_norm,
    simulate_projected,
    uninstall,
)
from .cabi import B2svError, EngineUnavailable, PlanOpts  # noqa: F401
from .engine import StateVector  # noqa: F401
from .fixtures import GateRecord, load_circuit, save_circuit  # noqa: F401
from .lowering import LoweringError, lower_gates  # noqa: F401
from .subspace_prog import SubspaceProgram  # noqa: F401

__version__ = "0.1.0"
