"""pybnesian_b200 — B200-native (sm_100a) implementation of PyBNesian's KDE / CKDE
log-likelihood hot path behind the reference's Python API.

Same class names as `pybnesian` for the path: KDE, CKDE, CKDEType, BandwidthSelector,
NormalReferenceRule, ScottsBandwidth, SingularCovarianceData, ...
"""
from ._lib import SingularCovarianceData, Context, default_context, set_default_context, LIB_PATH
from .dataset import DataFrame, CrossValidation, HoldOut
from .kde import BandwidthSelector, NormalReferenceRule, ScottsBandwidth, UCV, UCVScorer, KDE, ProductKDE
from .factors import (Factor, FactorType, CKDE, CKDEType, LinearGaussianCPD, LinearGaussianCPDType,
                      UnknownFactorType, MLE, MLELinearGaussianCPD, LinearGaussianParams)
from .hybrid import (Assignment, DiscreteFactor, DiscreteFactorType, HCKDE, CLinearGaussianCPD, MLEDiscreteFactor,
                     DiscreteFactorParams)
from .models import (Dag, BayesianNetwork, BayesianNetworkType, GaussianNetwork, GaussianNetworkType, KDENetwork,
                     KDENetworkType, SemiparametricBN, SemiparametricBNType, HeterogeneousBN, HeterogeneousBNType, load)
from .scores import (Args, Kwargs, Arguments, Score, ValidatedScore, BIC, CVLikelihood, HoldoutLikelihood,
                     ValidatedLikelihood)
from .operators import (Operator, ArcOperator, AddArc, RemoveArc, FlipArc, ChangeNodeType, OperatorTabuSet,
                        LocalScoreCache, OperatorSet, ArcOperatorSet, ChangeNodeTypeSet, OperatorPool)
from .hillclimbing import GreedyHillClimbing, Callback, SaveModel, hc
from . import parallel

__version__ = "0.1.0"
