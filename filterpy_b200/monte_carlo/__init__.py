"""GPU mirrors of ``filterpy.monte_carlo`` for the hot path."""
from .resampling import systematic_resample, stratified_resample, ResamplePlan, normalize_weights  # noqa: F401
