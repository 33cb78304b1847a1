"""Callers of the hot path (reference experiments/run_spectral_analysis.py:297-412,505-513)."""
