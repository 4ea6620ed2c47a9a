"""ccvpe_b200 -- B200-native cross-view matching + pose decoder of CCVPE."""
__version__ = "0.1.0"
