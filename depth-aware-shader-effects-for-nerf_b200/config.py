"""Configuration object with the reference's attribute names (the reference's config.py:3-36), so that `NeRF(Config())`
and scripts written against the reference keep working.

The ray-marching kernels are specialised for the architecture group below (include/nerfw.h: NERFW_HIDDEN, NERFW_LAYERS,
NERFW_SKIP, NERFW_POS_LEVELS, NERFW_DIR_LEVELS, NERFW_APP_DIM); `nerfw.NeRF` raises for any other value.  Everything else
is carried for compatibility only and keeps the reference's defaults."""
import torch

# attribute -> default, grouped by who reads it
_ARCHITECTURE = dict(pos_enc_levels=10, dir_enc_levels=4, hidden_dim=256, num_layers=8, skip_connect_layers=[4],
                     use_appearance=True, appearance_dim=32)
_SAMPLING = dict(near=2.0, far=6.0, num_samples=64, num_importance=64)
_TRAINING = dict(batch_size=1024, learning_rate=5e-4, num_iterations=30000, scheduler_step_size=10000, scheduler_gamma=0.5)
_DATASET = dict(dataset_type="nerf_synthetic", dataset_path="data/nerf_synthetic", scene="lego")


class Config:
    """Class attributes, like the reference (callers read `config.hidden_dim` on the class or on an instance)."""
    device = torch.device("cuda" if torch.cuda.is_available() else "cpu")


for _group in (_ARCHITECTURE, _SAMPLING, _TRAINING, _DATASET):
    for _name, _value in _group.items():
        setattr(Config, _name, _value)
