"""B200-native flow-LoRA hot path (ConditionalCFM training step + Euler-ODE inference over the
ConditionalDecoder estimator) behind the reference's Python API; see DESIGN.md."""
__version__ = "0.1.0"
