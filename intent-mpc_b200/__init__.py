"""intent-mpc_b200: B200-native batched QP engine for Intent-MPC's mpcPlanner hot path
(trajectory_planner/include/trajectory_planner/mpcPlanner.cpp:375-541)."""
__version__ = "0.1.0"
