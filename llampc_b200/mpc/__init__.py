from .lookback import LookBack  # noqa: F401
from .lookahead import LookAhead  # noqa: F401
from .evaluate_models_vectorized import evaluate_models_vectorized  # noqa: F401
from .mu_estimator import MuEstimator  # noqa: F401
