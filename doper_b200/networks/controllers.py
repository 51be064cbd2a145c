"""The reference's controller (doper/networks/controllers.py:6-22): a 25-50-100-2 ReLU MLP
(6,602 parameters).  It stays ``torch.nn`` — it is not on the rollout hot path;
it is here so that the trainer adapter (doper_b200.trainers) is self-contained, and because its
gradients are the one thing the multi-GPU path all-reduces."""
__all__ = ["LinearReLUController"]

import torch.nn as nn


class LinearReLUController(nn.Module):
    def __init__(self, model_config):
        super().__init__()
        d_in, d_h, d_out = model_config["input_dim"], model_config["hidden_dim"], model_config["output_dim"]
        self.controller = nn.Sequential(
            nn.Linear(d_in, d_h), nn.ReLU(inplace=True),
            nn.Linear(d_h, 2 * d_h), nn.ReLU(inplace=True),
            nn.Linear(2 * d_h, d_out),
        )

    def forward(self, controller_input):
        return self.controller(controller_input)
