"""The product-layer dimension gate of the reference (preprocessing/config.py:87-92), with text-embedding-3-large admitted.

The reference refuses a collection whose dimension is not in SUPPORTED_DIMENSIONS = {128, 256, 768, 960, 1536}; nothing on the
search / build path needs that (any D with D % M == 0 works, 3072 is what BASELINE configs[4] runs).  A deployment that swaps
`pydiskann` for this package either applies integration/0001-accept-3072-dimensional-collections.patch to the reference, or
rebinds the two names from here:

    import preprocessing.config as cfg, diskrag_b200.config_gate as gate
    cfg.SUPPORTED_DIMENSIONS = gate.SUPPORTED_DIMENSIONS; cfg.validate_vector_dimension = gate.validate_vector_dimension
"""
from .build_index import adaptive_pq_subvectors

SUPPORTED_DIMENSIONS = {128, 256, 768, 960, 1536, 3072}


def validate_vector_dimension(dimension: int) -> bool:
    """preprocessing/config.py:90-92, same contract: is this dimension accepted for PQ-quantised collections?"""
    return dimension in SUPPORTED_DIMENSIONS


def install(config_module) -> None:
    """Rebind the reference module's gate in place (idempotent)."""
    config_module.SUPPORTED_DIMENSIONS = set(config_module.SUPPORTED_DIMENSIONS) | SUPPORTED_DIMENSIONS
    config_module.validate_vector_dimension = validate_vector_dimension


def pq_subvectors_for(dimension: int, n_points: int = 1_000_000, target_quality: str = "balanced") -> int:
    """Every admitted dimension must get a usable sub-vector count from the adaptive heuristic (adaptive_pq.py:81-108)."""
    return adaptive_pq_subvectors(n_points, dimension, target_quality)
