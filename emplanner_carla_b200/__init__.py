"""B200-native Frenet EM-planner hot path (S-L DP + QP, S-T DP, Cartesian<->Frenet)."""
__version__ = "0.1.0"
