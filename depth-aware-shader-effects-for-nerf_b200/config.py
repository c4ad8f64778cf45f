"""Hot-path configuration constants with the reference's attribute names (config.py:3-36).

Only what the ray-marching path reads; dataset/training-schedule attributes keep the reference defaults so scripts
written against the reference's Config keep working."""
import torch


class Config:
    dataset_type = "nerf_synthetic"
    dataset_path = "data/nerf_synthetic"
    scene = "lego"

    hidden_dim = 256
    num_layers = 8
    skip_connect_layers = [4]
    num_samples = 64
    num_importance = 64

    use_appearance = True
    appearance_dim = 32

    batch_size = 1024
    learning_rate = 5e-4
    num_iterations = 30000
    scheduler_step_size = 10000
    scheduler_gamma = 0.5

    near = 2.0
    far = 6.0

    pos_enc_levels = 10
    dir_enc_levels = 4

    device = torch.device("cuda" if torch.cuda.is_available() else "cpu")
