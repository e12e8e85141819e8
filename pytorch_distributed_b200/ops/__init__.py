"""Hand-written sm_100a ops with PyTorch reference fall-backs (CPU / oracle)."""
