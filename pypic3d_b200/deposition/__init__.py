from .Esirkepov import Esirkepov_current  # noqa: F401
from .J_from_rhov import J_from_rhov  # noqa: F401
from .rho import compute_rho  # noqa: F401
