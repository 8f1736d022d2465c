"""ripor_b200: B200-native constrained-beam-search retrieval engine behind RIPOR's generation API."""
__version__ = "0.1.0"
