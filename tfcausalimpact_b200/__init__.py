"""B200-native implementation of tfcausalimpact's fit / forecast / reduction hot path."""
__version__ = '0.1.0'
