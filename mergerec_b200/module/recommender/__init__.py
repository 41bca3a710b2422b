from .loss_fn import *  # noqa: F401,F403
from .module import RecEvaluation, RecJointEvaluation, WeightedMeanLog, catalog_cross_entropy  # noqa: F401,E402
