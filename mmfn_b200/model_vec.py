"""Entry-point module mirroring team_code/mmfn_utils/models/model_vec.py: `mmfn_b200.model_vec:MMFN`."""
from .model_rad import MMFNVec as MMFN  # noqa: F401
