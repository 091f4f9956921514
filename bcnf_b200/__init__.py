"""bcnf_b200 -- B200-native CondRealNVP_v2 coupling stack behind the psaegert/bcnf Python API."""
from .cnf import (ActNorm, ConditionalAffineCouplingLayer, ConditionalInvertibleLayer,  # noqa: F401
                  ConditionalNestedNeuralNetwork, CondRealNVP_v2, InvertibleLayer,
                  OrthonormalTransformation, PackedFlow)
from .factories import FeatureNetworkFactory, OptimizerFactory, SchedulerFactory  # noqa: F401
from .feature_network import (ConcatenateCondition, FeatureNetwork, FeatureNetworkStack,  # noqa: F401
                              FrExpFeatureNetwork, FullyConnectedFeatureNetwork, LSTMFeatureNetwork,
                              Transformer)
from .calibration import compute_CDF_residuals, compute_y_hat_ranks  # noqa: F401
from .resimulation import physics_ODE_simulation_batch, resimulate  # noqa: F401
from .train import FlatAdam, Trainer, fused_nll  # noqa: F401
from .utils import ParameterIndexMapping, inn_nll_loss, load_config  # noqa: F401

__version__ = "0.1.0"
