"""nanowakeword_b200 — B200-native engine for the nanowakeword per-window hot path."""
__version__ = "0.1.0"
