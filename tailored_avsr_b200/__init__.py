"""tailored_avsr_b200 — B200 (sm_100a) implementation of tailored-avsr's Branchformer encoder + CTC."""
__version__ = "0.1.0"
