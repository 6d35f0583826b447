"""B200-native implementation of tfcausalimpact's fit / forecast / reduction hot path behind the
unchanged `CausalImpact(data, pre_period, post_period, model_args={...})` API."""
__version__ = '0.1.0'


def __getattr__(name):
    if name == 'CausalImpact':
        from .main import CausalImpact
        return CausalImpact
    raise AttributeError(name)
