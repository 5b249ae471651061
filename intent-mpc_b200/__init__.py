"""intent-mpc_b200: B200-native batched QP engine for Intent-MPC's mpcPlanner hot path."""
__version__ = "0.1.0"
