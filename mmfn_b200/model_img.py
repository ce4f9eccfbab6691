"""Entry-point module mirroring team_code/mmfn_utils/models/model_img.py (the default
train_agent.entry_point, run_steps/config/train.yaml:13): `mmfn_b200.model_img:MMFN`."""
from .model_rad import MMFNImg as MMFN  # noqa: F401
