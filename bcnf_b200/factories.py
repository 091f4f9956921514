"""String -> object factories with the reference's names (reference src/bcnf/factories.py).

Only what the coupling-stack path and its callers need: feature networks of the BASELINE
configs, Adam and ReduceLROnPlateau.  Anything else raises NotImplementedError, like the
reference does for unknown names (factories.py:20, :30, :58, :73).
"""
from __future__ import annotations

from typing import Any, Iterator

import torch
from torch import nn

from .feature_network import (ConcatenateCondition, FeatureNetwork, FrExpFeatureNetwork,
                              FullyConnectedFeatureNetwork, LSTMFeatureNetwork, Transformer)

_OUT_OF_SCOPE = {"CNN", "DualDomainLSTM", "DualDomainTransformer", "DualDomainFC"}


class FeatureNetworkFactory:
    @staticmethod
    def get_feature_network(network: str | None, network_kwargs: Any) -> FeatureNetwork | nn.Module:
        table = {"FullyConnected": FullyConnectedFeatureNetwork, "LSTM": LSTMFeatureNetwork,
                 "Transformer": Transformer, "ConcatenateCondition": ConcatenateCondition,
                 "FrExpFeatureNetwork": FrExpFeatureNetwork}
        if network is None:
            return nn.Identity()
        if network in table:
            return table[network](**network_kwargs)
        if network in _OUT_OF_SCOPE:
            raise NotImplementedError(f"Feature network {network} is outside this build's scope (video / "
                                      "dual-domain encoders; SURVEY.md section 2) -- pass an nn.Module instance instead")
        raise NotImplementedError(f"Feature network {network} not implemented")


class OptimizerFactory:
    @staticmethod
    def get_optimizer(optimizer: str, parameters: Iterator[nn.Parameter], optimizer_kwargs: Any) -> torch.optim.Optimizer:
        if optimizer == "Adam":
            return torch.optim.Adam(parameters, **optimizer_kwargs)
        raise NotImplementedError(f"Optimizer {optimizer} not implemented")


class SchedulerFactory:
    @staticmethod
    def get_scheduler(scheduler: str, optimizer: torch.optim.Optimizer, scheduler_kwargs: Any):
        if scheduler == "ReduceLROnPlateau":
            return torch.optim.lr_scheduler.ReduceLROnPlateau(optimizer, **scheduler_kwargs)
        raise NotImplementedError(f"Scheduler {scheduler} not implemented")
