"""padel_analytics_b200 — B200-native (sm_100a) inference engine for the padel_analytics tracker hot path."""
__version__ = "0.1.0"
