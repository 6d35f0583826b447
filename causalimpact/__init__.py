"""Drop-in alias: `from causalimpact import CausalImpact` resolves to the B200-native package."""
from tfcausalimpact_b200.main import CausalImpact  # noqa: F401
from tfcausalimpact_b200 import __version__  # noqa: F401
